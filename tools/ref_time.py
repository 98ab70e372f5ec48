"""Time the UNMODIFIED reference (/root/reference, build container only) on the bench generator's
object-graph twin: C2FVarInference / VarInference, K=3, T=3, is_log=False, a few iterations of
ADAM_update on a sub-sample.  Writes profiles/r2_reference_python_timing.json, which bench.py quotes
in cpu_baseline.sample (the GPU box has no /root/reference)."""
import collections
import collections.abc
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
collections.MutableSet = collections.abc.MutableSet          # the two import shims of SURVEY section 8 c
np.Inf = np.inf
REF = "/root/reference"

import lhvi_b200

sys.path.insert(0, REF)
import importlib.util


def load_ref(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    P, G, K, T = int(sys.argv[1]) if len(sys.argv) > 1 else 60, 10, 3, 3
    out = {"generator": f"synthetic.relational_hybrid_graph({P}, {G}) (object-graph twin of the bench workload)",
           "K": K, "T": T, "host": "build container, 1 core (the reference is single-threaded)"}
    for name in ("VarInference", "C2FVarInference"):
        g = lhvi_b200.synthetic.relational_hybrid_graph(P, G, seed=0)[0]
        mod = load_ref(name)
        vi = mod.VarInference(g, K, T)
        its = 2 if name == "VarInference" else 10           # C2F runs whole rounds of 10
        np.random.seed(0)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            vi.run(its, lr=0.1, is_log=False)
        dt = (time.perf_counter() - t0) / its
        n_f = len(g.factors)
        out[name] = {"ground_factors": n_f, "iterations": its, "s_per_iter": dt, "ground_factors_per_s": n_f / dt}
        print(name, out[name], flush=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_reference_python_timing.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
