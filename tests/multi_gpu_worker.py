"""Worker of tests/test_gpu_multi.py (launched under torchrun, one rank per GPU): the sharded
engine -- owner-computes partition, in-kernel peer exchange or the compact collective -- must
reproduce the single-process numpy oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import lhvi_b200
    from lhvi_b200.engine import DeviceEngine
    from oracle.vi_numpy import NumpyVI, grad_pass

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    syn = lhvi_b200.synthetic
    models = {
        "relational": syn.relational_hybrid(3000, 5, 3, 3, seed=1, order="hub", weighted=True),
        "grid": syn.gaussian_grid(40, 2, 3),
    }
    steps = 5
    for name, model in models.items():
        eta, tau, w_tau = syn.random_state(model, 2)
        ref = NumpyVI(model)
        ref.eta[:], ref.tau[:], ref.w_tau = eta, tau, w_tau
        ref.refresh()
        g0, gw0, e0 = grad_pass(model, ref.eta, ref.w)
        for _ in range(steps):
            fe_last = ref.adam_step(0.1)
        for mode in ("p2p", "collective"):
            os.environ["LHVI_EXCHANGE"] = mode
            for dtype, tol in (("float64", 1e-9), ("float32", 2e-4)):
                eng = DeviceEngine(model, dtype=dtype, device=f"cuda:{local}")
                assert eng.exchange == mode, (eng.exchange, mode)
                eng.set_state(eta, tau, w_tau)
                eng.reset_moments()
                g, gw, e = eng.gradients()                      # dense sum over ranks
                np.testing.assert_allclose(e, e0, rtol=tol)
                np.testing.assert_allclose(gw, gw0, rtol=tol, atol=tol * np.abs(gw0).max())
                np.testing.assert_allclose(g, g0, rtol=tol, atol=tol * np.abs(g0).max())
                eng.iterate(steps, 0.1)
                e1, _, wt1, _ = eng.get_state()
                np.testing.assert_allclose(eng.last_free_energy(), fe_last, rtol=tol)
                np.testing.assert_allclose(e1, ref.eta, rtol=tol * 10, atol=tol * 10)
                np.testing.assert_allclose(wt1, ref.w_tau, rtol=tol * 10, atol=tol * 10)
                # belief / MAP queries read the merged state (every rank answers for every variable)
                m = model
                cont = np.flatnonzero(m.var_kind == 0)[:64]
                xq = ref.eta[m.var_off[cont]] + 0.3
                got = eng.mixture_belief(m.var_off[cont], m.var_dim[cont], m.var_kind[cont], xq).double().cpu().numpy()
                want = np.zeros(cont.size)
                for k in range(m.K):
                    mu, var = ref.eta[m.var_off[cont] + 2 * k], ref.eta[m.var_off[cont] + 2 * k + 1]
                    want += ref.w[k] * np.exp(-(xq - mu) ** 2 / (2 * var)) / (2.506628274631 * var)
                np.testing.assert_allclose(got, want, rtol=tol * 100, atol=tol * 100)
                # replicas of the shared state agree bit for bit
                wt = torch.as_tensor(wt1, device=f"cuda:{local}")
                lo, hi = wt.clone(), wt.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                assert torch.equal(lo, hi)
                if rank == 0:
                    print(f"[multi] {name} {mode} {dtype}: ok ({eng.plan.describe()})", flush=True)
                del eng
    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("MULTI_OK", flush=True)
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
