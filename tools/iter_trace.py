"""Where does an iteration of the persistent kernel go?  Runs the bench workload (or a smaller one) with
LHVI_ITER_TRACE=1 and prints, per record group, the time its blocks spent in it (mean / max over the
blocks), then the waits at the two grid barriers and the optimiser step -- from the %globaltimer stamps
lhvi_iterate writes into lhvi_optim::trace.  GPU box only.

    python tools/iter_trace.py [--entities N] [--groups G] [--iters n] [--dtype float32]
"""
import argparse
import os
import sys

os.environ["LHVI_ITER_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import lhvi_b200
from lhvi_b200.engine import DeviceEngine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entities", type=int, default=1_000_000)
    ap.add_argument("--groups", type=int, default=10)
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--dtype", default="float32")
    a = ap.parse_args()
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(a.entities, a.groups, 3, 3, seed=0, order="hub", weighted=True)
    eta, tau, w_tau = syn.random_state(model, 0)
    eng = DeviceEngine(model, dtype=a.dtype)
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    assert eng.persistent()
    eng.iterate(3, 0.1)
    torch.cuda.synchronize()
    eng.iterate(a.iters, 0.1)
    torch.cuda.synchronize()
    tr = eng.iter_trace.cpu().numpy().astype(np.int64)
    blocks = int(np.count_nonzero(tr.reshape(-1, 16)[:, 0])) // a.iters
    tr = tr[:a.iters * blocks * 16].reshape(a.iters, blocks, 16)
    print(f"plan={os.environ.get('LHVI_ITER_PLAN', 'split')} blocks={blocks} iter_plan={eng.iter_plan}")
    names = [f"{i}:{'node' if g.node else ('pure' if g.pure else 'full')} nc={g.nc} ne={g.ne} n={g.n}"
             for i, (_, _, g) in enumerate(eng.groups)]
    for it in range(1, a.iters):
        t = tr[it]
        t0 = t[:, 0].min()
        print(f"-- iteration {it}: {(t[:, 14].max() - t0) / 1e3:.1f} us (start skew {(t[:, 0].max() - t0) / 1e3:.1f})")
        for p, name in enumerate(names):
            did = t[:, 1 + p] > 0
            if not did.any():
                continue
            # duration of the phase in a block: from the previous stamp of that block
            prev = np.where(did[:, None], t[:, :13], 0)
            dur = []
            for b in np.flatnonzero(did):
                stamps = sorted(v for v in t[b, :12] if v > 0 and v < t[b, 1 + p])
                dur.append(t[b, 1 + p] - stamps[-1])
            dur = np.array(dur) / 1e3
            end = (t[did, 1 + p] - t0) / 1e3
            print(f"   {name:44s} blocks {int(did.sum()):3d}  in-phase mean {dur.mean():6.1f} max {dur.max():6.1f} us; "
                  f"ends at mean {end.mean():6.1f} max {end.max():6.1f}")
        print(f"   barrier 1 passed at {(t[:, 12].min() - t0) / 1e3:.1f}..{(t[:, 12].max() - t0) / 1e3:.1f}; "
              f"step done mean {(t[:, 13].mean() - t0) / 1e3:.1f} max {(t[:, 13].max() - t0) / 1e3:.1f} "
              f"(block 0: {(t[0, 13] - t0) / 1e3:.1f}); barrier 2 passed at {(t[:, 14].max() - t0) / 1e3:.1f}")


if __name__ == "__main__":
    main()
