"""World-size-2 run of the sharding protocol on CPU (gloo): ``ShardPlan`` partitions the
records owner-computes style, each rank evaluates its records, the ranks all-reduce only the
compact ``[G_w | energy | shared-variable gradients]`` vector, every rank steps the variables it
owns plus the shared ones, and ``ShardPlan.merge`` assembles the state -- which must match the
single-process result.  The per-rank pass is the numpy oracle standing in for the CUDA kernels;
the plan / exchange / merge objects are the ones the GPU engine uses."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import lhvi_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import lhvi_b200 as pkg
    from lhvi_b200.dist import ShardPlan
    from oracle.vi_numpy import NumpyVI, grad_pass, tau_gradients
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        syn = pkg.synthetic
        model = syn.relational_hybrid(60, 4, 2, 3, seed=3, weighted=True)
        eta, tau, w_tau = syn.random_state(model, 7)
        plan = ShardPlan()
        assert plan.world == world and plan.rank == rank
        mine = plan.shard(model)
        assert "shared" in plan.describe()
        shared = plan.shared_idx
        vi = NumpyVI(model)
        vi.eta[:], vi.tau[:], vi.w_tau = eta, tau, w_tau
        vi.refresh()
        K = model.K
        for _ in range(3):
            g, gw, e = grad_pass(mine, vi.eta, vi.w)
            x = torch.from_numpy(np.concatenate([gw, [e], g[shared]]))
            plan.all_reduce(x)                       # the only per-iteration exchange
            x = x.numpy()
            g[shared] = x[K + 1:]
            vi.gradients = lambda g=g, x=x: (*tau_gradients(model, g, x[:K], vi.eta, vi.w), x[K])
            vi.adam_step(0.1)
        merged = plan.merge(torch.from_numpy(vi.eta.copy())).numpy()
        np.save(os.path.join(out_dir, f"eta_{rank}.npy"), merged)
        np.save(os.path.join(out_dir, f"shared_{rank}.npy"), vi.eta[shared])
        np.save(os.path.join(out_dir, f"wtau_{rank}.npy"), vi.w_tau)
    finally:
        dist.destroy_process_group()


def test_two_rank_protocol_matches_single_process(tmp_path):
    from oracle.vi_numpy import NumpyVI
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(60, 4, 2, 3, seed=3, weighted=True)
    eta, tau, w_tau = syn.random_state(model, 7)
    ref = NumpyVI(model)
    ref.eta[:], ref.tau[:], ref.w_tau = eta, tau, w_tau
    ref.refresh()
    for _ in range(3):
        ref.adam_step(0.1)
    for r in range(world):
        np.testing.assert_allclose(np.load(tmp_path / f"eta_{r}.npy"), ref.eta, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(np.load(tmp_path / f"wtau_{r}.npy"), ref.w_tau, rtol=1e-10, atol=1e-12)
    # the merged state and the replicated shared variables are bit-identical across ranks
    assert np.array_equal(np.load(tmp_path / "eta_0.npy"), np.load(tmp_path / "eta_1.npy"))
    assert np.array_equal(np.load(tmp_path / "shared_0.npy"), np.load(tmp_path / "shared_1.npy"))
