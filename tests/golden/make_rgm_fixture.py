"""Fixture for the reference's RGM demo (Demo/RGM/RGMTimeLog.py: 100 categories x 10 banks, 1111
variables, 2100 Gaussian factors; evidence sets Demo/Data/RGM/time_log_{5,20}percent with 30 and
121 observed values).  Build container only: imports the UNMODIFIED reference from /root/reference.

Written to rgm_demo.json, per evidence set:
  evidence            the observed (ground atom key, value) pairs of the reference's data file
  recorded_final      the last free energies of Demo/Data/RGM/time_log_*_result.  They were logged
                      with an evidence set that is no longer in the repository (the unmodified
                      reference does not reach them on these files either), so they are kept for
                      information only.
  lvi_final           free energy of the reference's LiftedVarInference after 200 Adam iterations,
                      K=1, T=3, lr=0.2 (the demo's call), numpy seed 0.  With K=1 and Gaussian
                      factors the objective is convex, so the value does not depend on the draw.
  c2f_final / c2f_log the same for C2FVarInference, started from mu=0.5, var=1 in the single initial
                      hidden cluster (its 20 refinement rounds end 10 iterations after the last
                      split, so the value depends on the start); c2f_log = free energy every 10th
                      iteration; c2f_mu = final means of a few ground atoms.  The reference seeds
                      its evidence k-means with the first values of a Python set of RV objects
                      (CompressedGraphWithObs.py:94-98): the later rounds depend on object addresses
                      and differ from one run of the reference to the next; the first rounds, where
                      the two-centroid k-means converges to the same split from any seed, do not.

Usage: python make_rgm_fixture.py            (four reference runs in parallel, ~4 min)
"""
import collections
import collections.abc
import contextlib
import io
import json
import os
import sys
import types
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LHVI_REFERENCE", "/root/reference")
PROBES = [("recession", "all"), ("market", "c0"), ("market", "c50"), ("revenue", "b3"), ("loss", "c7", "b2")]


def reference_run(job):
    tag, engine = job
    import numpy as np
    collections.MutableSet = collections.abc.MutableSet           # OrderedSet.py:5 on python >= 3.10
    np.Inf = np.inf                                               # GaBP.py:2 on numpy 2
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    os.chdir(REF)
    from Demo.Data.RGM.Generator import generate_rel_graph, load_data
    rel_g = generate_rel_graph()
    data = load_data(f"Demo/Data/RGM/time_log_{tag}percent")
    rel_g.ground_graph()
    g, rvs_table = rel_g.add_evidence(data)
    np.random.seed(0)
    vi = __import__(engine).VarInference(g, num_mixtures=1, num_quadrature_points=3)
    if engine == "C2FVarInference":
        def init_param():
            vi.w_tau = np.zeros(1)
            vi.w = np.ones(1)
            vi.eta, vi.eta_tau = {}, {}
            for rv in vi.g.rvs:
                if rv.value is None:
                    vi.eta[rv] = np.array([[0.5, 1.0]])
        vi.init_param = init_param
    with contextlib.redirect_stdout(io.StringIO()):
        vi.run(200, lr=0.2)
    out = {"final": float(vi.free_energy()), "log": [float(fe) for _, fe in vi.time_log][9::10]}
    if engine == "C2FVarInference":
        out["mu"] = [[list(k), float(vi.eta[rvs_table[k].cluster][0, 0])] for k in PROBES
                     if rvs_table[k].value is None]
    return tag, engine, out


def main():
    out = {}
    for tag in ("5", "20"):
        data = json.load(open(os.path.join(REF, "Demo", "Data", "RGM", f"time_log_{tag}percent")))
        result = json.load(open(os.path.join(REF, "Demo", "Data", "RGM", f"time_log_{tag}_result")))
        out[tag] = {"evidence": [[list(eval(k)), v] for k, v in data.items()],
                    "recorded_final": {name: log[-1][1] for name, log in result.items()}}
    jobs = [(tag, engine) for tag in ("5", "20") for engine in ("LiftedVarInference", "C2FVarInference")]
    with ProcessPoolExecutor(4) as pool:
        for tag, engine, res in pool.map(reference_run, jobs):
            short = "lvi" if engine == "LiftedVarInference" else "c2f"
            out[tag][f"{short}_final"] = res["final"]
            out[tag][f"{short}_log"] = res["log"]
            if "mu" in res:
                out[tag]["c2f_mu"] = res["mu"]
            print(tag, engine, res["final"], flush=True)
    json.dump(out, open(os.path.join(HERE, "rgm_demo.json"), "w"))


if __name__ == "__main__":
    main()
