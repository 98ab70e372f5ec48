"""Where an iteration's time goes when replayed from a CUDA graph: times graphs of sub-sequences of
the iteration on the bench workload (one B200).  usage: python tools/timeline_probe.py [entities]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lhvi_b200
from lhvi_b200.engine import DeviceEngine

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 10
syn = lhvi_b200.synthetic
model = syn.relational_hybrid(P, G, 3, 3, seed=0, order="hub", weighted=True)
eta, tau, w_tau = syn.random_state(model, 0)
eng = DeviceEngine(model, dtype="float32", device="cuda:0")
eng.set_state(eta, tau, w_tau); eng.reset_moments()
eng.iterate(3, 0.1)
torch.cuda.synchronize()

def timed(fn, name, reps=30):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:55s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us", flush=True)

main = lambda: torch.cuda.current_stream()
n = len(eng.groups)
for i in range(n):
    d, _, g = eng.groups[i]
    timed(lambda i=i: eng._launch_group(i, main()), f"group {i} alone ({'node' if g.node else 'pure' if g.pure else 'full'} nc={g.nc} ne={g.ne} n={g.n})")
timed(lambda: eng._launch_groups(), "all groups (parallel branches)")
def serial():
    for i in range(n): eng._launch_group(i, main())
timed(serial, "all groups (one stream)")
timed(lambda: eng._finish(tick=False, exchange=False), "finish alone")
timed(lambda: eng.param_step(0.0, sgd=True, zero_grad=False), "param_step alone (lr=0)")
def gf():
    eng._launch_groups(); eng._finish(tick=False, exchange=False)
timed(gf, "groups + finish")
def full():
    eng._launch_groups(); eng._finish(tick=False, exchange=False); eng.param_step(0.0, sgd=True, zero_grad=True)
timed(full, "groups + finish + param_step (lr=0)")
