"""Parity of the CUDA path (through the C ABI) with the goldens of the reference and with
the numpy oracle.  Tolerances: fp64 mode 1e-6 relative per north_star (we assert tighter,
1e-9, since only the summation order differs); fp32 mode is checked against the fp64 oracle
with the tolerance stated at each assert."""
import contextlib
import io

import numpy as np
import pytest

import helpers
import lhvi_b200
import specs
from oracle.vi_numpy import NumpyVI, grad_pass
from test_dropin_host import ENGINE_CLASS, check_trajectory, run_trajectory

pytestmark = pytest.mark.gpu

FP64_RTOL = 1e-9


def _engine_for(model, dtype="float64", **kw):
    from lhvi_b200.engine import DeviceEngine
    return DeviceEngine(model, dtype=dtype, **kw)


def _golden_setup(path, ns):
    name, engine, gold = helpers.load_golden(path)
    builder, K, T, _ = specs.CASES[name]
    g, rvs = builder(ns)
    handles, _, cg = helpers.setup_mode(g, engine)
    model = helpers.lower_for(engine, g, cg, K, T)
    cont, disc = helpers.injected_params(handles, rvs, engine, K, int(gold["seed"]))
    tau = helpers.flat_params(model, cont, disc)
    ref = NumpyVI(model)
    ref.eta[:] = tau
    ref.tau[:] = tau
    ref.w_tau = helpers.injected_w_tau(K)
    ref.refresh()
    return name, engine, gold, model, rvs, ref


@pytest.mark.parametrize("path", helpers.golden_files(gpu=True), ids=helpers.golden_id)
@pytest.mark.parametrize("force_generic", [True, False], ids=["generic", "dispatch"])
def test_snapshot_fp64(path, force_generic, ns):
    name, engine, gold, model, rvs, ref = _golden_setup(path, ns)
    eng = _engine_for(model, "float64", force_generic=force_generic)
    eng.set_state(ref.eta, ref.tau, ref.w_tau)
    grad, g_w, energy = eng.gradients()
    # against the reference (patched for H2) through the golden file
    np.testing.assert_allclose(energy, gold["fe0_fixed"], rtol=FP64_RTOL)
    from oracle.vi_numpy import tau_gradients
    g_flat, g_wtau = tau_gradients(model, grad, g_w, ref.eta, ref.w)
    np.testing.assert_allclose(g_wtau, gold["gw0_fixed"], rtol=FP64_RTOL, atol=1e-11)
    want = gold["grad0_fixed"]
    got = helpers.rows_from_flat(model, g_flat, rvs, engine, want.shape[1])
    np.testing.assert_allclose(got, want, rtol=FP64_RTOL, atol=1e-10)
    # and against the oracle on the raw accumulators
    og, ogw, oe = grad_pass(model, ref.eta, ref.w)
    np.testing.assert_allclose(grad, og, rtol=FP64_RTOL, atol=1e-10)
    np.testing.assert_allclose(g_w, ogw, rtol=FP64_RTOL, atol=1e-10)


@pytest.mark.parametrize("path", helpers.golden_files(gpu=True), ids=helpers.golden_id)
def test_snapshot_fp32(path, ns):
    name, engine, gold, model, rvs, ref = _golden_setup(path, ns)
    eng = _engine_for(model, "float32")
    eng.set_state(ref.eta, ref.tau, ref.w_tau)
    grad, g_w, energy = eng.gradients()
    og, ogw, oe = grad_pass(model, ref.eta, ref.w)
    scale = max(1.0, np.abs(og).max())
    # fp32 arithmetic: 2e-5 of the largest gradient entry
    np.testing.assert_allclose(energy, oe, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(g_w, ogw, rtol=2e-5, atol=2e-5 * max(1.0, np.abs(ogw).max()))
    np.testing.assert_allclose(grad, og, rtol=2e-5, atol=2e-5 * scale)


@pytest.mark.parametrize("path", helpers.golden_files(gpu=True), ids=helpers.golden_id)
def test_trajectory_fp64(path, ns):
    """Drop-in classes on the real engine: same Adam trajectory as the reference, including
    the C2F split rounds."""
    name, engine, gold = helpers.load_golden(path)
    vi, rvs = run_trajectory(name, engine, ns, gold, lambda v: v)
    check_trajectory(vi, rvs, engine, gold, rtol=1e-7, atol=1e-9)


def test_param_step_matches_oracle_sgd_and_adam(ns):
    g, rvs = specs.robot_like(ns)
    model = lhvi_b200.lowering.lower_ground(g, 3, 3)
    eta, tau, w_tau = lhvi_b200.synthetic.random_state(model, 4)
    for sgd in (False, True):
        ref = NumpyVI(model)
        ref.eta[:], ref.tau[:], ref.w_tau = eta, tau, w_tau + np.array([0.1, -0.2, 0.05])
        ref.refresh()
        eng = _engine_for(model)
        eng.set_state(ref.eta, ref.tau, ref.w_tau)
        eng.reset_moments()
        for _ in range(4):
            ref.sgd_step(0.05) if sgd else ref.adam_step(0.1)
        eng.iterate(4, 0.05 if sgd else 0.1, sgd=sgd)
        e2, t2, wt2, w2 = eng.get_state()
        np.testing.assert_allclose(e2, ref.eta, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(wt2, ref.w_tau, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(w2, ref.w, rtol=1e-9)


def test_queries_between_iterations_keep_the_gradient_buffer_consistent():
    """iterate() clears the gradient slots inside the parameter step instead of a memset;
    gradients() / free_energy() calls in between must not leak into the next iteration."""
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(150, 4, 2, 3, seed=5, weighted=True)
    eta, tau, w_tau = syn.random_state(model, 6)
    ref = NumpyVI(model)
    ref.eta[:], ref.tau[:], ref.w_tau = eta, tau, w_tau
    ref.refresh()
    eng = _engine_for(model)
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    for n in (1, 2, 1):
        g, gw, e = eng.gradients()
        og, ogw, oe = grad_pass(model, ref.eta, ref.w)
        np.testing.assert_allclose(e, oe, rtol=FP64_RTOL)
        np.testing.assert_allclose(g, og, rtol=FP64_RTOL, atol=1e-10)
        for _ in range(n):
            fe = ref.adam_step(0.1)
        eng.iterate(n, 0.1)
        np.testing.assert_allclose(eng.last_free_energy(), fe, rtol=FP64_RTOL)
        e2, _, wt2, _ = eng.get_state()
        np.testing.assert_allclose(e2, ref.eta, rtol=1e-9, atol=1e-11)
    # eager launches (no CUDA graph, sequential groups) give the same trajectory
    eng2 = _engine_for(model)
    eng2.use_graph = False
    eng2.parallel_groups = False
    eng2.set_state(eta, tau, w_tau)
    eng2.reset_moments()
    eng2.iterate(4, 0.1)
    np.testing.assert_allclose(eng2.get_state()[0], ref.eta, rtol=1e-9, atol=1e-11)


def test_belief_queries(ns):
    g, rvs = specs.hmln_hidden(ns)
    vi = lhvi_b200.VarInference.VarInference(g, 2, 3)
    np.random.seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        vi.run(3, lr=0.2, is_log=False)
    for rv in rvs:
        if rv.value is None:
            xs = (0.37, -1.2) if rv.domain.continuous else rv.domain.values
            for x in xs:
                assert np.isclose(vi.belief(x, rv), vi.rvs_belief((x,), (rv,)), rtol=1e-12)


def test_batched_map_finds_the_mixture_modes(ns):
    """lhvi_mixture_map against a dense scan + scipy refinement of every variable's mixture
    (the reference calls scipy.optimize.minimize per variable, VarInference.py:355-376)."""
    from scipy.optimize import minimize_scalar
    g, rvs = specs.hmln_hidden(ns)
    vi = lhvi_b200.VarInference.VarInference(g, 3, 3)
    np.random.seed(8)
    with contextlib.redirect_stdout(io.StringIO()):
        vi.run(4, lr=0.3, is_log=False)
    table = vi.map_all()
    assert set(table) == {rv for rv in rvs if rv.value is None}
    for rv in rvs:
        if rv.value is not None:
            assert vi.map(rv) == rv.value
            continue
        if not rv.domain.continuous:
            marg = [vi.belief(x, rv) for x in rv.domain.values]
            assert vi.map(rv) == rv.domain.values[int(np.argmax(marg))]
            continue
        eta = vi.eta[rv]
        f = lambda x: -float(vi.rvs_belief((x,), (rv,)))
        # the reference's answer: local optimum reached from the best component mean
        x0 = eta[int(np.argmax([-f(m) for m in eta[:, 0]])), 0]
        res = minimize_scalar(f, bracket=(x0 - 1e-3, x0 + 1e-3), tol=1e-12)
        got = vi.map(rv)
        assert -f(got) >= -f(x0) - 1e-15
        assert abs(got - res.x) < 1e-5 or abs(f(got) - res.fun) < 1e-12


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 5e-5)])
def test_generated_models_against_oracle(dtype, tol):
    syn = lhvi_b200.synthetic
    for model in (syn.relational_hybrid(300, 5, 3, 3, seed=2, weighted=True),
                  syn.relational_hybrid(300, 5, 2, 3, seed=2, order="entity"),
                  syn.gaussian_grid(12, 1, 3), syn.gaussian_grid(7, 2, 5)):
        eta, tau, w_tau = syn.random_state(model, 1)
        K = model.K
        w = np.full(K, 1.0 / K)
        og, ogw, oe = grad_pass(model, eta, w)
        eng = _engine_for(model, dtype)
        eng.set_state(eta, tau, w_tau)
        grad, g_w, energy = eng.gradients()
        np.testing.assert_allclose(energy, oe, rtol=tol)
        np.testing.assert_allclose(g_w, ogw, rtol=tol, atol=tol * np.abs(ogw).max())
        np.testing.assert_allclose(grad, og, rtol=tol, atol=tol * np.abs(og).max())


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 5e-5)])
@pytest.mark.parametrize("K", [1, 2, 3])
def test_run_major_kernel_matches_record_major_and_oracle(dtype, tol, K):
    """Full records with a few-valued (hub) argument run through the run-major kernel
    (lhvi_group::run_*): same sums as the record-major kernel and the oracle, for both positions
    of the hub argument, ragged runs, runs longer than the split length, weighted and not."""
    import dataclasses
    syn = lhvi_b200.synthetic
    rng = np.random.default_rng(5)
    for weighted, swap in ((True, False), (False, True)):
        model = syn.relational_hybrid(400, 7, K, 3, seed=3, weighted=weighted, observed_frac=0.5)
        groups = []
        for g in model.groups:
            if not g.node and not g.pure and g.nc == 2:
                keep = np.flatnonzero(rng.random(g.n) < 0.6)            # ragged runs
                g = g.take(keep)
                if weighted:
                    g = dataclasses.replace(g, wf=rng.uniform(0.5, 3.0, g.n),
                                            gam=rng.integers(0, 4, size=(2, g.n)).astype(np.float64))
                if swap:                                               # hub argument first
                    blocks = {}
                    pot = g.pot.copy()
                    for b in np.unique(g.pot):
                        blocks[b] = b
                    ptab = model.ptab.copy()
                    for b in np.unique(g.pot):                         # c, l0, l1, A00, A01, A11 -> swap roles
                        c, l0, l1, a00, a01, a11 = ptab[b:b + 6]
                        ptab = np.concatenate([ptab, [c, l1, l0, a11, a01, a00]])
                        pot[g.pot == b] = ptab.size - 6
                    model = dataclasses.replace(model, ptab=ptab)
                    g = dataclasses.replace(g, pot=pot.astype(np.int32), poff=g.poff[::-1].copy(), gam=g.gam[::-1].copy())
            groups.append(g)
        model = dataclasses.replace(model, groups=groups)
        eta, tau, w_tau = syn.random_state(model, 1)
        w = np.full(K, 1.0 / K)
        og, ogw, oe = grad_pass(model, eta, w)
        outs = []
        for run_major in (True, False):
            eng = _engine_for(model, dtype, run_major=run_major)
            used = [bool(d.run_start) for d, _, _ in eng.groups]
            assert any(used) == run_major
            eng.set_state(eta, tau, w_tau)
            outs.append(eng.gradients())
        for grad, g_w, energy in outs:
            np.testing.assert_allclose(energy, oe, rtol=tol)
            np.testing.assert_allclose(g_w, ogw, rtol=tol, atol=tol * np.abs(ogw).max())
            np.testing.assert_allclose(grad, og, rtol=tol, atol=tol * np.abs(og).max())


def test_run_major_kernel_floors():
    """Deep-tail potentials and far-apart components on a hub model: the run-major kernel must
    take the literal path (log(psi + 1e-100), log(b + 1e-100) in double) like the reference."""
    syn = lhvi_b200.synthetic
    K = 2
    model = syn.relational_hybrid(200, 4, K, 3, seed=1, observed_frac=0.4)
    eta, tau, w_tau = syn.random_state(model, 0)
    cont = model.var_off.astype(np.int64)
    eta[cont] = -40.0
    eta[cont + 2] = 45.0
    eta[cont + 1] = 0.1
    eta[cont + 3] = 0.1
    w_tau = np.array([-60.0, 0.0])
    e = np.e ** w_tau
    og, ogw, oe = grad_pass(model, eta, e / e.sum())
    for dtype, tol in (("float64", 1e-9), ("float32", 1e-4)):
        eng = _engine_for(model, dtype)
        assert any(bool(d.run_start) for d, _, _ in eng.groups)
        eng.set_state(eta, tau, w_tau)
        grad, g_w, energy = eng.gradients()
        assert np.isfinite(grad).all() and np.isfinite(g_w).all()
        np.testing.assert_allclose(energy, oe, rtol=tol)
        np.testing.assert_allclose(g_w, ogw, rtol=tol, atol=tol * np.abs(ogw).max())
        np.testing.assert_allclose(grad, og, rtol=tol * 10, atol=tol * np.abs(og).max())


def test_fp32_underflow_fallback():
    """Far-apart mixture components: float pdfs underflow to zero, the kernel must redo those
    points in double like the reference's log(b + 1e-100)."""
    syn = lhvi_b200.synthetic
    model = syn.gaussian_grid(4, 2, 3)
    eta, tau, w_tau = syn.random_state(model, 0)
    K = 2
    cont = model.var_off.astype(np.int64)
    eta[cont] = -40.0          # component 0 mean
    eta[cont + 2] = 45.0       # component 1 mean
    eta[cont + 1] = 0.1
    eta[cont + 3] = 0.1
    w_tau = np.array([-60.0, 0.0])      # w_0 ~ 1e-26
    e = np.e ** w_tau
    og, ogw, oe = grad_pass(model, eta, e / e.sum())
    eng = _engine_for(model, "float32")
    eng.set_state(eta, tau, w_tau)
    grad, g_w, energy = eng.gradients()
    assert np.isfinite(grad).all() and np.isfinite(g_w).all()
    np.testing.assert_allclose(energy, oe, rtol=1e-4)
    np.testing.assert_allclose(g_w, ogw, rtol=1e-4, atol=1e-4 * np.abs(ogw).max())


def test_large_scale_properties():
    """1 M-record model (fp32): the pass is a sum over records, so (a) it is invariant to the
    record order and (b) shards add up; (c) a bounded random sample of records agrees with
    the oracle evaluated on the same sample."""
    syn = lhvi_b200.synthetic
    P, G, K, T = 100_000, 10, 3, 3
    hub = syn.relational_hybrid(P, G, K, T, seed=0, order="hub")
    ent = syn.relational_hybrid(P, G, K, T, seed=0, order="entity")
    eta, tau, w_tau = syn.random_state(hub, 0)
    outs = []
    for m in (hub, ent):
        eng = _engine_for(m, "float32")
        eng.set_state(eta, tau, w_tau)
        outs.append(eng.gradients())
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-5)
    np.testing.assert_allclose(outs[0][1], outs[1][1], rtol=1e-5)
    scale = np.abs(outs[0][0]).max()
    # hub gradients are fp32 sums of 1e5 terms accumulated in a different order
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-3, atol=1e-3 * scale)
    # shards add up (what the NCCL all-reduce computes)
    tot = None
    for r in range(4):
        eng = _engine_for(hub.shard(r, 4), "float32")
        eng.set_state(eta, tau, w_tau)
        part = eng.gradients()
        tot = part if tot is None else tuple(a + b for a, b in zip(tot, part))
    np.testing.assert_allclose(tot[2], outs[0][2], rtol=1e-5)
    np.testing.assert_allclose(tot[0], outs[0][0], rtol=1e-3, atol=1e-3 * scale)
    # bounded sample against the oracle
    rng = np.random.default_rng(0)
    sample_groups = []
    for g in hub.groups:
        sel = np.sort(rng.choice(g.n, size=min(g.n, 2000), replace=False))
        sample_groups.append(g.take(sel))
    import dataclasses
    sample = dataclasses.replace(hub, groups=sample_groups)
    w = np.full(K, 1.0 / K)
    og, ogw, oe = grad_pass(sample, eta, w)
    eng = _engine_for(sample, "float32")
    eng.set_state(eta, tau, w_tau)
    grad, g_w, energy = eng.gradients()
    np.testing.assert_allclose(energy, oe, rtol=2e-5)
    np.testing.assert_allclose(grad, og, rtol=1e-4, atol=2e-5 * np.abs(og).max())


@pytest.mark.parametrize("weighted", [True, False])
def test_tile_aligned_hub_runs_through_every_kernel(weighted, monkeypatch):
    """Hub runs of the streamed group padded to whole tiles with null records (engine.align_runs):
    the streaming kernel, the per-record kernels and the generic kernel all sum the same values."""
    from lhvi_b200 import engine as eng_mod
    monkeypatch.setattr(eng_mod, "FOLD_ALIGN_MIN_TILES", 0)          # pad every run, however short
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(500, 4, 3, 3, seed=6, weighted=weighted)
    eta, tau, w_tau = syn.random_state(model, 2)
    w = np.full(3, 1.0 / 3)
    og, ogw, oe = grad_pass(model, eta, w)
    for dtype, tol in (("float64", 1e-9), ("float32", 5e-5)):
        for force_generic in (False, True):
            eng = _engine_for(model, dtype, force_generic=force_generic)
            padded = [d for d, _, g in eng.groups if d.fold and (d.hub_mask >> g.nd) & 1]
            assert padded and all(int(d.n) % 1024 == 0 and int(d.n) >= 4 * 1024 for d in padded)
            eng.set_state(eta, tau, w_tau)
            grad, g_w, energy = eng.gradients()
            np.testing.assert_allclose(energy, oe, rtol=tol)
            np.testing.assert_allclose(g_w, ogw, rtol=tol, atol=tol * np.abs(ogw).max())
            np.testing.assert_allclose(grad, og, rtol=tol, atol=tol * np.abs(og).max())


def test_state_pack_unpack_round_trip(ns):
    """lhvi_state_pack / lhvi_state_unpack: compact per-variable arrays <-> padded device slots."""
    import torch
    syn = lhvi_b200.synthetic
    for model in (syn.relational_hybrid(200, 4, 3, 3, seed=0), syn.gaussian_grid(6, 2, 3)):
        eta, tau, w_tau = syn.random_state(model, 3)
        for dtype in ("float64", "float32"):
            eng = _engine_for(model, dtype)
            eng.set_state(eta, tau, w_tau)
            mp = eng.packed_map()
            idx = eng.packed_index
            assert idx.size == int((model.K * model.var_dim).sum()) and (np.diff(idx) > 0).all()
            packed = torch.empty(idx.size, dtype=eng.tdtype, device=eng.device)
            eng.pack_state(packed, "eta")
            np.testing.assert_array_equal(packed.cpu().numpy(), eta[idx].astype(packed.cpu().numpy().dtype))
            eng.eta.zero_()
            eng.unpack_state(packed, "eta")
            got = eng.eta.double().cpu().numpy()
            np.testing.assert_array_equal(got[idx], eta[idx].astype(packed.cpu().numpy().dtype).astype(np.float64))
            mask = np.ones(got.size, bool); mask[idx] = False
            assert (got[mask] == 0).all()


def test_cabi_rejects_bad_arguments():
    import ctypes as C
    from lhvi_b200 import _cabi
    syn = lhvi_b200.synthetic
    eng = _engine_for(syn.gaussian_grid(3, 1, 3))
    d, _, _ = eng.groups[-1]
    saved = d.pot
    d.pot = None
    rc = eng.lib.lhvi_factor_expect_grad(C.byref(eng.desc), C.byref(d), 0, 0, None)
    d.pot = saved
    assert rc == -1
    with pytest.raises(ValueError):
        _cabi.check(rc, eng.lib)
    with pytest.raises(ValueError):
        lhvi_b200.lowering.lower_ground(syn.gaussian_grid_graph(2)[0], 9, 3)
