"""Colour passing on the device (lifting_torch) against the host library (lhvi_lift.cpp, all host threads) on the
bench generator's ground graph: same class ids, time per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lhvi_b200
from lhvi_b200 import lifting, lifting_torch as lt, synthetic as syn

for P in [int(a) for a in sys.argv[1:]] or [100_000]:
    ga = syn.relational_hybrid_arrays(P, 10, seed=0)
    t0 = time.perf_counter()
    tg = lt.TorchGraph(ga, "cuda")
    torch.cuda.synchronize()
    t_prep = time.perf_counter() - t0
    for split in (False, True):
        start = lifting.initial_colouring(ga, split_cont_evidence=split)
        lifting.colour_passing(ga, start=start, use_native=True)            # prepares the host graph
        t0 = time.perf_counter()
        v0, f0, s0 = lifting.colour_passing(ga, start=start, use_native=True)
        t_host = time.perf_counter() - t0
        dstart = torch.as_tensor(start).cuda()
        lt.colour_passing(tg, dstart)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v1, f1, s1 = lt.colour_passing(tg, dstart)
        torch.cuda.synchronize()
        t_dev = time.perf_counter() - t0
        same = np.array_equal(v0, v1.cpu().numpy()) and all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(f0, f1))
        print(f"P={P} ground factors {ga.n_factors} variables {ga.n_vars} start split={split}: classes {int(v0.max()) + 1}, "
              f"sweeps {s0}/{s1}, identical ids {same}; host {t_host * 1e3:.1f} ms, device {t_dev * 1e3:.1f} ms "
              f"(graph upload + incidence sort once: {t_prep * 1e3:.0f} ms)", flush=True)
