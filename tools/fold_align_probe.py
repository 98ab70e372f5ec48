"""Does the streamed unary kernel lose time at hub boundaries that fall inside a tile?  Times the
group as generated and with every hub's run truncated to a multiple of 1024 records."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dataclasses
import numpy as np, torch
import lhvi_b200
from lhvi_b200.engine import DeviceEngine

P = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
syn = lhvi_b200.synthetic
base = syn.relational_hybrid(P, 10, 3, 3, seed=0, order="hub", weighted=True)
eta, tau, w_tau = syn.random_state(base, 0)
gi = [i for i, g in enumerate(base.groups) if g.pure and g.nc == 1 and g.ne == 1][0]
for label in ("as generated", "hub runs truncated to multiples of 1024"):
    model = base
    if label != "as generated":
        g = base.groups[gi]
        key = g.poff[0]
        starts = np.concatenate([[0], np.flatnonzero(key[1:] != key[:-1]) + 1, [g.n]])
        sel = np.concatenate([np.arange(a, a + (b - a) // 1024 * 1024) for a, b in zip(starts[:-1], starts[1:])])
        groups = list(base.groups); groups[gi] = g.take(sel)
        model = dataclasses.replace(base, groups=groups)
    eng = DeviceEngine(model, dtype="float32", device="cuda:0")
    eng.set_state(eta, tau, w_tau); eng.reset_moments()
    eng.iterate(2, 0.1); torch.cuda.synchronize()
    s = torch.cuda.current_stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng._launch_group(gi, torch.cuda.current_stream())
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{label:45s} n={model.groups[gi].n:9d}  {e0.elapsed_time(e1) / 30 * 1e3:7.1f} us", flush=True)
