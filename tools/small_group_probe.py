"""What each small record group of the bench model costs the iteration: the replayed iteration with all groups,
and with the node-entropy / unary-prior / constant groups taken out of the model one at a time (results are then
wrong; only the time matters).  With the engine's fusions on (default) those groups are empty and the numbers
coincide; LHVI_FUSE_RUN_EXTRAS=0 LHVI_FUSE_CONSTANTS=0 shows what they cost as launches of their own."""
import dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lhvi_b200
from lhvi_b200.engine import DeviceEngine
syn = lhvi_b200.synthetic
model = syn.relational_hybrid(1_000_000, 10, 3, 3, seed=0, order="hub", weighted=True)
state = syn.random_state(model, 0)
def timed(m, label):
    eng = DeviceEngine(m, dtype="float32")
    eng.set_state(*state); eng.reset_moments()
    eng.iterate(5, 0.1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        e0.record(); eng.iterate(50, 0.1); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 50 * 1e3)
    print(f"{label:60s} {min(ts):7.1f} us (min of 5 x 50 iterations)", flush=True)
    eng.close()
timed(model, "all groups")
keep = lambda pred: dataclasses.replace(model, groups=[g for g in model.groups if pred(g)])
timed(keep(lambda g: not g.node), "without the node group")
timed(keep(lambda g: not (g.pure and g.nc == 1 and g.ne == 0)), "without the prior group (pure nc=1 ne=0)")
timed(keep(lambda g: not g.node and not (g.pure and g.nc == 1 and g.ne == 0)), "without node and prior groups")
timed(keep(lambda g: not (g.pure and g.nc == 0)), "without the constants group (pure nc=0)")
timed(keep(lambda g: (g.nc == 2) or (g.pure and g.nc == 1 and g.ne == 1)), "run-major + streamed only")
