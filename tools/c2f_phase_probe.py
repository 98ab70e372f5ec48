"""Where a coarse-to-fine round's device passes spend their time (full config-5 size by default): wraps the
functions of lifting_torch with synchronised timers and prints totals over a run."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lhvi_b200
from lhvi_b200 import lifting, lifting_torch as lt, synthetic as syn

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
its = int(sys.argv[2]) if len(sys.argv) > 2 else 100
tot, cnt = collections.Counter(), collections.Counter()
calls = {}
def wrap(name):
    fn = getattr(lt, name)
    def timed(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); tot[name] += time.perf_counter() - t0; cnt[name] += 1
        calls.setdefault(name, []).append(round((time.perf_counter() - t0) * 1e3, 1))
        return out
    setattr(lt, name, timed)
for n in ("colour_passing", "refine_bookkeeping", "layout", "inherit", "split_evidence", "lower_partition", "class_stats",
          "initial_colouring", "rank_first", "TorchGraph"):
    wrap(n)
lifting.C2FArrayVI(syn.relational_hybrid_arrays(2000, 10, seed=1), 3, 3, dtype="float32").run(20, 0.05)
tot.clear(); cnt.clear(); calls.clear()
ga = syn.relational_hybrid_arrays(P, 10, seed=0)
vi = lifting.C2FArrayVI(ga, 3, 3, dtype="float32")
t0 = time.perf_counter(); vi.run(its, 0.05); total = time.perf_counter() - t0
print(f"run {total:.3f} s; phases {dict((k, round(v, 3)) for k, v in vi.timing.items())}")
for k, v in tot.most_common():
    print(f"  {k:20s} {v * 1e3:8.1f} ms in {cnt[k]:4d} calls ({v / cnt[k] * 1e3:6.2f} ms each)   [nested calls are counted in their callers too]")
for k in ("TorchGraph", "initial_colouring", "colour_passing", "lower_partition", "layout", "split_evidence"):
    print(f"  {k}: per call ms {calls.get(k)}")
