"""Which call invalidates the CUDA-graph capture of an iteration?  Runs the drop-in classes over
every golden like tests/test_gpu_parity.py::test_trajectory_fp64, with the stream's capture status
queried after every C-ABI call made during capture (GPU box only; prints one line per golden)."""
import contextlib
import ctypes
import io
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)

import torch

import conftest
import helpers
import specs
from test_dropin_host import ENGINE_CLASS, make_injector

cudart = ctypes.CDLL("libcudart.so.12")
STATUS = {0: "none", 1: "active", 2: "INVALIDATED"}


def capture_status():
    st = ctypes.c_int(-1)
    rc = cudart.cudaStreamIsCapturing(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(st))
    return f"rc={rc} {STATUS.get(st.value, st.value)}"


def instrument(eng, log):
    lib = eng.lib
    for name in ("lhvi_factor_expect_grad", "lhvi_step_tick", "lhvi_finish_step", "lhvi_finish", "lhvi_param_step"):
        orig = getattr(lib, name)

        def wrapped(*a, _orig=orig, _name=name):
            rc = _orig(*a)
            log.append((_name, rc, capture_status()))
            return rc
        setattr(lib, name, wrapped)


def main():
    ns = conftest.repo_namespace()
    only = sys.argv[1:] or None
    for path in helpers.golden_files():
        name, engine, gold = helpers.load_golden(path)
        gid = helpers.golden_id(path)
        if only and gid not in only:
            continue
        builder, K, T, _ = specs.CASES[name]
        g, rvs = builder(ns)
        vi = ENGINE_CLASS[engine]()(g, K, T)
        vi.init_param = make_injector(vi, rvs, engine, K, int(gold["seed"]))
        log = []
        orig_make = vi._make_engine

        def make(model, _orig=orig_make, _log=log):
            eng = _orig(model)
            instrument(eng, _log)
            return eng
        vi._make_engine = make
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            with contextlib.redirect_stdout(io.StringIO()):
                vi.run(int(gold["steps"]), lr=float(gold["lr"]), is_log=False)
        bad = [str(w.message)[:80] for w in caught if "capture" in str(w.message)]
        print(gid, "CAPTURE FAILED" if bad else "ok", flush=True)
        if bad:
            for entry in log:
                print("   ", entry)


if __name__ == "__main__":
    main()
