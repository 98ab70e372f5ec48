"""fp32 against fp64 at the bench size: error of the gradients of the ten group-level variables (sums of ~10^6
cancelling terms) and of everything else, for the specialised and the generic kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lhvi_b200
from lhvi_b200.engine import DeviceEngine

syn = lhvi_b200.synthetic
model = syn.relational_hybrid(1_000_000, 10, 3, 3, seed=0, order="hub", weighted=True)
state = syn.random_state(model, 0)

def run(dtype, **kw):
    eng = DeviceEngine(model, dtype=dtype, **kw)
    eng.set_state(*state)
    g, gw, e = eng.gradients()
    eng.close()
    return np.array(g, dtype=np.float64), np.array(gw, dtype=np.float64), float(e)

count = np.zeros(model.n_param, dtype=np.int64)
for g in model.groups:
    for a in range(g.nh):
        count += np.bincount(g.poff[a], minlength=model.n_param)
hub = (np.flatnonzero(count > 100_000)[:, None] + np.arange(6)[None, :]).reshape(-1)
exact = run("float64")
for name, kw in (("specialised fp32", {}), ("generic fp32", {"force_generic": True})):
    got = run("float32", **kw)
    eh = np.abs(got[0][hub] - exact[0][hub])
    rest = np.ones(model.n_param, dtype=bool); rest[hub] = False
    er = np.abs(got[0][rest] - exact[0][rest])
    print(f"{name}: hub gradients |value| {np.abs(exact[0][hub]).min():.3g}..{np.abs(exact[0][hub]).max():.3g}, "
          f"max abs error {eh.max():.3g}, max relative error {(eh / np.abs(exact[0][hub])).max():.3g}; "
          f"other gradients max |value| {np.abs(exact[0][rest]).max():.3g}, max abs error {er.max():.3g}; "
          f"G_w relative error {np.abs(got[1] / exact[1] - 1).max():.2g}, free energy relative error {abs(got[2] / exact[2] - 1):.2g}", flush=True)
print("hub (mu, var) gradients, fp64:", np.round(exact[0][hub][:12], 1))
