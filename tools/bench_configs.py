"""Auxiliary figures of bench.py, one per BASELINE.json config besides the headline (config 5), plus
an fp64 line of the headline workload and the drop-in classes at demo size.  Each returns a small
dict; bench.py puts them under ``configs`` / ``fp64`` / ``dropin`` of its JSON line (single-GPU run
only) and records an exception as ``{"error": ...}`` instead of failing the line.

Config 1  Demo/HMLN paper-popularity hybrid MLN at its own size (3 390 factors, 2 173 hidden booleans,
          the reference's evidence file), ground VarInference settings K=2, Gauss-Hermite degree 10,
          1000 iterations (SURVEY section 8 d).
Config 2  relational Kalman filter, 1000 state dimensions x 100 steps (100 k hidden state variables +
          observation leaves), lifted: colour passing, compression ratio, iterations on the classes.
Config 3  paper-popularity scaled to 1 M ground link factors, K=3 (relation atoms observed).
Config 4  1000 x 1000 pairwise Gaussian grid with one noisy observation per node (1 M variables,
          3 M factors), K=1: VI means against the exact solve of J mu = h (conjugate gradients).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _timed_iterations(eng, n, lr, repeats=3):
    """Milliseconds per iteration of ``eng.iterate(n)`` (CUDA events, best of ``repeats``)."""
    import torch
    eng.iterate(min(n, 8), lr)                       # warm-up (and the persistent kernel's tuning launches)
    best = None
    for _ in range(repeats):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        eng.iterate(n, lr)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        best = ms if best is None else min(best, ms)
    return best


def _namespace():
    import types

    import lhvi_b200
    ns = types.SimpleNamespace()
    for mod in (lhvi_b200.Graph, lhvi_b200.Potential, lhvi_b200.MLNPotential, lhvi_b200.RelationalGraph):
        for name in dir(mod):
            if not name.startswith("_"):
                setattr(ns, name, getattr(mod, name))
    return ns


def config1(cpu=True):
    import ctypes as C

    import lhvi_b200
    import specs
    from lhvi_b200.engine import DeviceEngine
    builder, _, _, _ = specs.CASES["hmln_demo"]
    g, _ = builder(_namespace())
    model = lhvi_b200.lowering.lower_ground(g, 2, 10)
    eta, tau, w_tau = lhvi_b200.synthetic.random_state(model, 3)
    out = {"workload": "Demo/HMLN paper-popularity (DemoPaperPopularity.py, evidence file Demo/Data/HMLN/0)",
           "factor_records": int(model.n_records), "hidden_variables": int(model.n_vars), "K": 2, "T": 10,
           "iterations": 1000,
           "reference_published_s_per_iter": {"value": 2.93, "note": "BASELINE.md section 1: VI, K=2, T=3 (degree 10 costs "
                                              "about 11x more grid points), unknown CPU, one core"}}
    for dtype in ("float32", "float64"):
        eng = DeviceEngine(model, dtype=dtype)
        spec = all(eng.lib.lhvi_has_specialisation(C.byref(eng.desc), C.byref(d)) == 1 for d, _, _ in eng.groups)
        eng.set_state(eta, tau, w_tau)
        eng.reset_moments()
        ms = _timed_iterations(eng, 1000, 0.2, repeats=2)
        out[f"ms_per_iter_{dtype}"] = ms
        out[f"free_energy_{dtype}"] = eng.last_free_energy()
        out["specialised_kernels_only"] = bool(spec)
        out["launch_mode"] = "persistent" if eng.persistent() else "CUDA graph of per-group launches"
        eng.close()
    if cpu:
        from oracle import cpu_port
        runner = cpu_port.make_runner(model, eta, tau, w_tau)
        runner.step(0.2)
        t0 = time.perf_counter()
        for _ in range(5):
            runner.step(0.2)
        out["cpu_port_ms_per_iter"] = (time.perf_counter() - t0) / 5 * 1e3
        out["cpu_port"] = f"{runner.describe}, {runner.cores} thread(s)"
    return out


def config2(dtype="float32", n=1000, t_steps=100, period=10):
    import lhvi_b200
    lifting, syn = lhvi_b200.lifting, lhvi_b200.synthetic
    t0 = time.perf_counter()
    ga, _ = syn.kalman_arrays(n, t_steps, levels=2, seed=0, period=period)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    vi = lifting.ArrayVI(ga, 1, 3, lifted=True, dtype=dtype)
    t_lift = time.perf_counter() - t0
    ms = _timed_iterations(vi.engine, 200, 0.1)
    out = {"workload": f"relational Kalman filter, {n} state dimensions x {t_steps} steps, sparse transition, observations "
                       f"quantised to 2 levels and repeating every {period} state dimensions (synthetic.kalman_arrays), "
                       "LiftedVarInference over the colour-passing partition (lifting.ArrayVI)",
           "ground_variables": int(ga.n_vars), "ground_factors": int(ga.n_factors),
           "variable_classes": int(vi.quotient.n_var_classes), "compressed_records": int(vi.model.n_records),
           "compression_ratio_factors": float(ga.n_factors / max(1, vi.model.n_records)),
           "K": 1, "T": 3, "ground_build_s": t_build, "colour_passing_lowering_upload_s": t_lift,
           "ms_per_iter": ms, "ground_factors_per_s": ga.n_factors / (ms * 1e-3),
           "launch_mode": "persistent" if vi.engine.persistent() else "CUDA graph of per-group launches",
           "free_energy": float(vi.free_energy())}
    vi.engine.close()
    return out


def config3(dtype="float32", entities=100_000, groups=10):
    import lhvi_b200
    from lhvi_b200.engine import DeviceEngine
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(entities, groups, 3, 3, seed=0, order="hub", weighted=False)
    eta, tau, w_tau = syn.random_state(model, 0)
    eng = DeviceEngine(model, dtype=dtype)
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    ms = _timed_iterations(eng, 100, 0.1)
    out = {"workload": f"paper-popularity hybrid MLN scaled to {entities} papers x {groups} topics = "
                       f"{entities * groups} ground link factors (+ priors, sessions), relation atoms observed, "
                       "70 % of the popularities observed, ground record format",
           "factor_records": int(model.n_records), "K": 3, "T": 3, "ms_per_iter": ms,
           "it_per_s": 1e3 / ms, "factor_records_per_s": model.n_records / (ms * 1e-3),
           "launch_mode": "persistent" if eng.persistent() else "CUDA graph of per-group launches",
           "free_energy_last": eng.last_free_energy()}
    eng.close()
    return out


def grid_with_observations(n, K, T, *, obs_var=0.5, seed=0):
    """n x n pairwise Gaussian grid (synthetic.gaussian_grid's attractive edges) with one noisy
    observation y_i per node, tied to x_i by LinearGaussianPotential(1, obs_var): the lowered model
    and the sparse system J mu = h of its exact posterior means."""
    import scipy.sparse as sp

    import lhvi_b200
    from lhvi_b200.lowering import EC, HC, LoweredModel, PotentialTable, slot_size
    from lhvi_b200.Potential import GaussianPotential, LinearGaussianPotential
    syn = lhvi_b200.synthetic
    rng = np.random.default_rng(seed)
    V = n * n
    y = rng.uniform(-3.0, 3.0, size=V)
    table = PotentialTable()
    slot = slot_size(2 * K)
    var_off = (np.arange(V, dtype=np.int64) * slot).astype(np.int32)
    idx = np.arange(V).reshape(n, n)
    right = np.stack([idx[:, :-1].reshape(-1), idx[:, 1:].reshape(-1)])
    down = np.stack([idx[:-1, :].reshape(-1), idx[1:, :].reshape(-1)])
    edges = np.concatenate([right, down], axis=1)
    edges = edges[:, np.argsort(edges[0], kind="stable")]
    deg = np.bincount(edges.reshape(-1), minlength=V) + 1
    blk_u = table.block(LinearGaussianPotential(1.0, obs_var), (HC, EC), (None, 0.0))
    blk_e = table.block(GaussianPotential([0.0, 0.0], syn._GRID_SIG), (HC, HC), (None, None))
    E = edges.shape[1]
    scale = (deg - 1 - 1).astype(np.float64)          # minus the unary observation factor (unary split)
    groups = [
        syn._group(0, 1, 0, 0, True, np.zeros(V, np.int32), var_off[None, :], np.zeros((0, V)), nscale=scale,
                   wf=scale.copy()),
        syn._group(0, 1, 0, 1, False, np.full(V, blk_u), var_off[None, :], y[None, :], pure=True),
        syn._group(0, 2, 0, 0, False, np.full(E, blk_e), var_off[edges], np.zeros((0, E))),
    ]
    model = LoweredModel(K, T, int(V * slot), np.zeros(V, np.uint8), np.full(V, 2, np.int32), var_off,
                         table.array(), groups)
    prec = np.linalg.inv(np.array(syn._GRID_SIG))
    diag = np.full(V, 1.0 / obs_var)
    np.add.at(diag, edges[0], prec[0, 0])
    np.add.at(diag, edges[1], prec[1, 1])
    J = sp.coo_matrix((np.concatenate([diag, np.full(E, prec[0, 1]), np.full(E, prec[0, 1])]),
                       (np.concatenate([np.arange(V), edges[0], edges[1]]),
                        np.concatenate([np.arange(V), edges[1], edges[0]]))), shape=(V, V)).tocsr()
    return model, J, y / obs_var


def config4(dtype="float32", n=1000, iterations=3000, lr=0.05):
    import scipy.sparse.linalg as spla

    import lhvi_b200
    from lhvi_b200.engine import DeviceEngine
    model, J, h = grid_with_observations(n, 1, 3)
    t0 = time.perf_counter()
    exact, info = spla.cg(J, h, rtol=1e-10, maxiter=2000)
    t_cg = time.perf_counter() - t0
    eta, tau, w_tau = lhvi_b200.synthetic.random_state(model, 0)
    eng = DeviceEngine(model, dtype=dtype)
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    ms = _timed_iterations(eng, 100, lr)
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    eng.iterate(iterations, lr)
    e1 = eng.get_state()[0]
    mu = e1[model.var_off]
    err = np.abs(mu - exact)
    out = {"workload": f"{n} x {n} pairwise Gaussian grid, one noisy observation per node "
                       "(LinearGaussianPotential), attractive GaussianPotential edges",
           "variables": int(n * n), "factor_records": int(model.n_records), "K": 1, "T": 3,
           "ms_per_iter": ms, "it_per_s": 1e3 / ms, "iterations_to_compare": iterations, "lr": lr,
           "cross_check": "posterior means vs the exact solve of J mu = h (scipy conjugate gradients, rtol 1e-10; "
                          "K=1 mean-field stationary means are exact for a Gaussian model)",
           "max_abs_mean_error": float(err.max()), "mean_abs_mean_error": float(err.mean()),
           "cg_info": int(info), "cg_seconds": t_cg,
           "launch_mode": "persistent" if eng.persistent() else "CUDA graph of per-group launches"}
    eng.close()
    # the cross-check BASELINE names: Gaussian belief propagation on the device at full size (GaBP.py,
    # csrc/lhvi_gabp.cu), timed with CUDA events; bytes per sweep: per directed edge 3 indices + 5
    # coefficients + 2 messages read, 2 written, 2 gathered totals, 2 REDs; per variable 2 read, 2 written
    import torch
    from lhvi_b200 import GaBP as gabp
    arrays = gabp.gabp_from_model(model)
    bp = gabp.DeviceGaBP(arrays, dtype)
    sweeps = 200
    bp.sweeps(5)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    bp.sweeps(sweeps)
    ev1.record()
    torch.cuda.synchronize()
    gms = ev0.elapsed_time(ev1) / sweeps
    gmean, gvar = bp.marginals()
    gerr = np.abs(gmean.double().cpu().numpy() - exact)
    esz = 4 if dtype == "float32" else 8
    gbytes = int(arrays.src.size) * (3 * 4 + 13 * esz) + int(arrays.n_vars) * 4 * esz
    out["gabp"] = {"engine": "DeviceGaBP (lhvi_gabp_sweeps)", "directed_edges": int(arrays.src.size), "sweeps": sweeps + 5,
                   "ms_per_sweep": gms, "algorithmic_gbs": gbytes / (gms * 1e-3) / 1e9,
                   "max_abs_mean_error_vs_exact": float(gerr.max()),
                   "max_abs_vi_minus_gabp_mean": float(np.abs(mu - gmean.double().cpu().numpy()).max())}
    return out


def fp64_headline(a):
    """The headline workload in fp64 (the drop-in classes' default dtype; north_star's 1e-6 mode)."""
    import lhvi_b200
    from lhvi_b200.engine import DeviceEngine
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(a.entities, a.groups, a.K, a.T, seed=0, order=a.order, weighted=True)
    eta, tau, w_tau = syn.random_state(model, 0)
    eng = DeviceEngine(model, dtype="float64")
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    ms = _timed_iterations(eng, 10, 0.1)
    out = {"dtype": "f64", "factor_records": int(model.n_records), "ms_per_step": ms, "value": 1e3 / ms, "unit": "it/s",
           "launch_mode": "persistent" if eng.persistent() else "CUDA graph of per-group launches",
           "free_energy_last": eng.last_free_energy()}
    eng.close()
    return out


def dropin_rgm(iterations=200):
    """The public drop-in classes at demo size: Demo/RGM (1111 variables, 2100 factors, 20 % evidence),
    K=1, T=3, lr 0.2, 200 iterations -- the run BASELINE.md times at 3.33 s / iteration for the
    reference's VarInference (1.16 LVI, 1.11 C2FVI) in the build container."""
    import lhvi_b200
    import relational_specs
    fix = json.load(open(os.path.join(ROOT, "tests", "golden", "rgm_demo.json")))
    out = {"workload": "Demo/RGM relational Gaussian model, 20 % evidence (tests/golden/rgm_demo.json), K=1, T=3, lr 0.2",
           "iterations": iterations,
           "reference_s_per_iter": {"VarInference": 3.33, "LiftedVarInference": 1.16, "C2FVarInference": 1.11,
                                    "note": "unmodified reference, one core of the build container (BASELINE.md section 2)"}}
    for label, cls, kw in (("VarInference", lhvi_b200.VarInference.VarInference, dict(is_log=False)),
                           ("VarInference_is_log", lhvi_b200.VarInference.VarInference, dict(is_log=True)),
                           ("LiftedVarInference", lhvi_b200.LiftedVarInference.VarInference, dict(is_log=False)),
                           ("C2FVarInference", lhvi_b200.C2FVarInference.VarInference, dict(is_log=False))):
        rel, _ = relational_specs.rgm_relational(_namespace(), 100, 10)
        data = {tuple(k): v for k, v in fix["20"]["evidence"]}
        g, _ = rel.ground_graph()
        rel.add_evidence(data)
        np.random.seed(1)
        t0 = time.perf_counter()
        vi = cls(g, 1, 3, dtype="float64")
        t_ctor = time.perf_counter() - t0
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            vi.run(iterations, lr=0.2, **kw)
        wall = time.perf_counter() - t0
        out[label] = {"constructor_s": t_ctor, "run_wall_s": wall, "s_per_iter_wall": wall / iterations,
                      "free_energy": float(vi.free_energy())}
        if kw.get("is_log"):
            out[label]["update_s_per_iter"] = vi.total_time / iterations       # what the reference's time_log measures
    return out


ALL = {"1": config1, "2": config2, "3": config3, "4": config4}
