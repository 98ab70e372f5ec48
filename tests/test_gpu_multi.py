"""Two ranks on two GPUs of one box (needs `gpurun --gpus 2`; skipped on a single GPU): the
owner-computes partition with the in-kernel NVLink exchange and with the compact NCCL
all-reduce, fp64 and fp32, against the single-process oracle (tests/multi_gpu_worker.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_the_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "MULTI_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
