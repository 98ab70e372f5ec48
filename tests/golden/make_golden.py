"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case in ``specs.CASES`` and every engine listed for it, the reference class is
instantiated on the graph, parameters are injected deterministically (``specs.inject_values``)
and the following are stored in ``<case>__<engine>.npz`` (arrays indexed by the creation
order of the ground variables):

* ``fe0, gw0``                free_energy() and gradient_w_tau() at the injected parameters
* ``grad0[i]``                gradient_mu_var / gradient_category_tau of variable i (NaN-padded)
* ``cluster[i]``              class id of variable i (lifted engines; -1 for ground)
* ``w_tau1, eta1[i], fe1``    state after ``steps`` Adam iterations at ``lr``
* ``*_fixed``                 the same from a copy of the class whose ``gradient_category_tau``
                              builds the other arguments' axes from ``rv_.domain`` (the
                              intended maths) -- patched in memory, never on disk.

Nothing from /root/reference is copied into the repository: only numbers.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LHVI_REFERENCE", "/root/reference")
SEED = 2024
LR = 0.2
STEPS = 5
C2F_ITERS = 20

sys.path.insert(0, REF)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import collections  # noqa: E402
import collections.abc  # noqa: E402

collections.MutableSet = collections.abc.MutableSet      # OrderedSet.py:5 on py>=3.10
if not hasattr(np, "Inf"):
    np.Inf = np.inf                                      # utils.py:11 on numpy 2

import Graph as ref_graph  # noqa: E402
import MLNPotential as ref_mln  # noqa: E402
import Potential as ref_pot  # noqa: E402
import specs  # noqa: E402


def reference_namespace():
    ns = types.SimpleNamespace()
    for mod in (ref_graph, ref_pot, ref_mln):
        for name in dir(mod):
            if not name.startswith("_"):
                setattr(ns, name, getattr(mod, name))
    return ns


def load_engine(module_name, fixed):
    """The reference class, optionally with the H2 lines patched in memory."""
    path = os.path.join(REF, module_name + ".py")
    src = open(path).read()
    if fixed:
        head, rest = src.split("    def gradient_category_tau", 1)
        body, tail = rest.split("    def free_energy", 1)
        assert "elif rv.domain.continuous:" in body and "(rv.domain.values, self.eta[rv_][k])" in body
        body = body.replace("elif rv.domain.continuous:", "elif rv_.domain.continuous:")
        body = body.replace("(rv.domain.values, self.eta[rv_][k])", "(rv_.domain.values, self.eta[rv_][k])")
        src = head + "    def gradient_category_tau" + body + "    def free_energy" + tail
    mod = types.ModuleType(module_name + ("_fixed" if fixed else ""))
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod.VarInference


ENGINE_MODULE = {"ground": "VarInference", "lifted": "LiftedVarInference", "c2f": "C2FVarInference"}


def min_index(handle, order):
    members = getattr(handle, "rvs", None)
    if members is None:
        return order[handle]
    return min(order[rv] for rv in members)


def make_injector(vi, order, K):
    def inject():
        vi.w_tau = np.linspace(-0.3, 0.4, K) if K > 1 else np.zeros(1)
        vi.eta, vi.eta_tau = {}, {}
        for rv in vi.g.rvs:
            if rv.value is not None:
                continue
            idx = min_index(rv, order)
            if rv.domain.continuous:
                vi.eta[rv] = specs.inject_values(idx, True, 0, K, SEED)
            else:
                vi.eta_tau[rv] = specs.inject_values(idx, False, len(rv.domain.values), K, SEED)
        vi.w = vi.softmax(vi.w_tau)
        for rv, table in vi.eta_tau.items():
            vi.eta[rv] = vi.softmax(table, 1)
    return inject


def zero_moments(vi, K):
    vi.alpha, vi.b1, vi.b2, vi.eps = LR, 0.9, 0.999, 1e-8
    vi.w_tau_g = [np.zeros(K), np.zeros(K)]
    vi.eta_g = [dict(), dict()]
    vi.eta_tau_g = [dict(), dict()]
    for rv in vi.g.rvs:
        if rv.value is not None:
            continue
        if rv.domain.continuous:
            vi.eta_g[0][rv] = np.zeros((K, 2))
            vi.eta_g[1][rv] = np.zeros((K, 2))
        else:
            D = len(rv.domain.values)
            vi.eta_tau_g[0][rv] = np.zeros((K, D))
            vi.eta_tau_g[1][rv] = np.zeros((K, D))
    vi.t = 0


def pack(rows, width):
    out = np.full((len(rows), width), np.nan)
    for i, r in enumerate(rows):
        if r is not None:
            r = np.asarray(r, dtype=float).reshape(-1)
            out[i, :r.size] = r
    return out


def snapshot(vi, rvs, engine, K, width):
    """fe, gw, per-ground-variable gradient rows, cluster ids."""
    handle = (lambda rv: rv) if engine == "ground" else (lambda rv: rv.cluster)
    fe = float(vi.free_energy())
    gw = np.asarray(vi.gradient_w_tau(), dtype=float)
    cache, rows = {}, []
    for rv in rvs:
        h = handle(rv)
        if h.value is not None:
            rows.append(None)
            continue
        if h not in cache:
            cache[h] = (vi.gradient_mu_var(h) if h.domain.continuous else vi.gradient_category_tau(h))
        rows.append(cache[h])
    ids = {}
    cluster = np.array([-1 if engine == "ground" else ids.setdefault(handle(rv), len(ids)) for rv in rvs])
    return fe, gw, pack(rows, width), cluster


def params(vi, rvs, engine, width):
    handle = (lambda rv: rv) if engine == "ground" else (lambda rv: rv.cluster)
    rows = [None if handle(rv).value is not None else vi.eta[handle(rv)] for rv in rvs]
    ids = {}
    cluster = np.array([-1 if engine == "ground" else ids.setdefault(handle(rv), len(ids)) for rv in rvs])
    ev = np.array([np.nan if handle(rv).value is None else float(handle(rv).value) for rv in rvs])
    return pack(rows, width), cluster, ev


def run_case(name, engine, fixed):
    builder, K, T, _ = specs.CASES[name]
    ns = reference_namespace()
    g, rvs = builder(ns)
    order = {rv: i for i, rv in enumerate(rvs)}
    width = K * max([2] + [len(rv.domain.values) for rv in rvs if not rv.domain.continuous])
    cls = load_engine(ENGINE_MODULE[engine], fixed)
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        vi = cls(g, K, T)
        if engine == "c2f":
            vi.g.init_cluster(is_split_cont_evidence=False)
            make_injector(vi, order, K)()
            zero_moments(vi, K)
            vi.cp_run()
        else:
            make_injector(vi, order, K)()
        out["fe0"], out["gw0"], out["grad0"], out["cluster0"] = snapshot(vi, rvs, engine, K, width)

        # optimiser trajectory from the same injected start
        g2, rvs2 = builder(reference_namespace())
        order2 = {rv: i for i, rv in enumerate(rvs2)}
        vi2 = cls(g2, K, T)
        vi2.init_param = make_injector(vi2, order2, K)
        if engine == "c2f":
            vi2.run(C2F_ITERS, lr=LR, is_log=False)
        else:
            vi2.run(STEPS, lr=LR, is_log=False)
        out["w_tau1"] = np.asarray(vi2.w_tau, dtype=float)
        out["eta1"], out["cluster1"], out["evidence1"] = params(vi2, rvs2, engine, width)
        out["fe1"] = float(vi2.free_energy())
    return out


def main():
    only = set(sys.argv[1:])
    for name, (_, K, T, engines) in specs.CASES.items():
        if only and name not in only:
            continue
        for engine in engines:
            data = {"K": K, "T": T, "lr": LR,
                    "steps": C2F_ITERS if engine == "c2f" else STEPS, "seed": SEED}
            for fixed in (False, True):
                res = run_case(name, engine, fixed)
                sfx = "_fixed" if fixed else ""
                for k, v in res.items():
                    data[k + sfx] = v
            path = os.path.join(HERE, f"{name}__{engine}.npz")
            np.savez_compressed(path, **data)
            dfe = abs(data["fe1"] - data["fe1_fixed"])
            print(f"{name:16s} {engine:7s} fe0={data['fe0']:+.10e} fe1={data['fe1']:+.10e} "
                  f"|fe1-fe1_fixed|={dfe:.2e}  h2_fires={bool(np.nanmax(np.abs(data['grad0'] - data['grad0_fixed'])) > 1e-12)}")


if __name__ == "__main__":
    main()
