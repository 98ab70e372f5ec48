"""The persistent iteration kernel (``lhvi_iterate``: n iterations, every record group, reduction and
optimiser step in one cooperative launch) against the numpy oracle and against the per-group
launches it replaces (reference loop: ``VarInference.py:249-300``)."""
import numpy as np
import pytest

import lhvi_b200
from oracle.vi_numpy import NumpyVI

pytestmark = pytest.mark.gpu


def _models():
    syn = lhvi_b200.synthetic
    return {
        "relational_weighted": syn.relational_hybrid(3000, 5, 3, 3, seed=1, order="hub", weighted=True),
        "relational_entity_order": syn.relational_hybrid(2000, 7, 3, 3, seed=2, order="entity", weighted=True),
        "relational_k2": syn.relational_hybrid(1500, 4, 2, 3, seed=3, weighted=True),
        "relational_unweighted": syn.relational_hybrid(1200, 3, 3, 3, seed=4, weighted=False),
        "grid_k1": syn.gaussian_grid(24, 1, 3),
        "grid_k3": syn.gaussian_grid(12, 3, 3),
    }


def _run(model, dtype, persistent, steps, sgd=False, chunks=(None,)):
    from lhvi_b200.engine import DeviceEngine
    syn = lhvi_b200.synthetic
    eta, tau, w_tau = syn.random_state(model, 7)
    eng = DeviceEngine(model, dtype=dtype)
    eng.use_persistent = persistent
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    lr = 0.02 if sgd else 0.1
    for c in chunks:
        eng.iterate(steps if c is None else c, lr, sgd=sgd)
    out = eng.get_state() + (eng.last_free_energy(), eng.get_moments())
    return eng, out


@pytest.mark.parametrize("name", sorted(_models()))
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 2e-4)])
def test_persistent_matches_oracle_and_per_group_launches(name, dtype, tol):
    model = _models()[name]
    steps = 6
    syn = lhvi_b200.synthetic
    eta, tau, w_tau = syn.random_state(model, 7)
    ref = NumpyVI(model)
    ref.eta[:], ref.tau[:], ref.w_tau = eta, tau, w_tau
    ref.refresh()
    for _ in range(steps):
        fe_last = ref.adam_step(0.1)
    eng, (e1, t1, wt1, w1, fe1, mom) = _run(model, dtype, True, steps)
    assert eng.persistent(), "the model must run in the persistent kernel (no silent fall-back)"
    assert eng.launch_count <= 3            # the first two launches measure (two iterations each) and re-split
    np.testing.assert_allclose(fe1, fe_last, rtol=tol)
    np.testing.assert_allclose(e1, ref.eta, rtol=tol * 10, atol=tol * 10)
    np.testing.assert_allclose(wt1, ref.w_tau, rtol=tol * 10, atol=tol * 10)
    np.testing.assert_allclose(w1, ref.w, rtol=tol * 10, atol=tol * 10)
    assert mom[4] == steps                                   # the step counter ran inside the launch
    eng2, (e2, t2, wt2, w2, fe2, mom2) = _run(model, dtype, False, steps)
    assert not eng2.persistent() and eng2.launch_count > steps
    np.testing.assert_allclose(e1, e2, rtol=tol * 10, atol=tol * 10)
    np.testing.assert_allclose(fe1, fe2, rtol=tol)
    np.testing.assert_allclose(mom[0], mom2[0], rtol=tol * 100, atol=tol * 10)


def test_one_launch_of_n_equals_n_launches_of_one():
    """iterate(6) == iterate(1) x 6 == iterate(2); iterate(4): the state carried between the passes of
    one launch (parameters, moments, step counter, clean gradient slots) is the state between launches."""
    model = _models()["relational_weighted"]
    _, a = _run(model, "float64", True, 6)
    _, b = _run(model, "float64", True, 6, chunks=(1,) * 6)       # single iterations: the captured per-group launches
    _, c = _run(model, "float64", True, 6, chunks=(2, 4))
    for x in (b, c):
        np.testing.assert_allclose(x[0], a[0], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(x[2], a[2], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(x[4], a[4], rtol=1e-11)


def test_persistent_sgd_and_queries_in_between():
    model = _models()["grid_k3"]
    syn = lhvi_b200.synthetic
    eta, tau, w_tau = syn.random_state(model, 7)
    ref = NumpyVI(model)
    ref.eta[:], ref.tau[:], ref.w_tau = eta, tau, w_tau
    ref.refresh()
    from lhvi_b200.engine import DeviceEngine
    eng = DeviceEngine(model, dtype="float64")
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    assert eng.persistent()
    for n in (2, 1):
        fe = eng.free_energy()                 # per-group launches in between leave the slots dirty
        for _ in range(n):
            fe_ref = ref.sgd_step(0.02)
        eng.iterate(n, 0.02, sgd=True)
        np.testing.assert_allclose(eng.get_state()[0], ref.eta, rtol=1e-9, atol=1e-11)
    assert np.isfinite(fe) and np.isfinite(fe_ref)
    assert eng.get_moments()[4] == 0           # SGD does not advance the Adam step counter


def test_models_without_a_body_fall_back_loudly_not_wrongly():
    """A model with hidden discrete arguments has no body in the iteration kernel: the engine says so
    (``persistent()`` is False) and runs the captured per-group launches."""
    import conftest
    import specs
    ns = conftest.repo_namespace()
    g, _ = specs.robot_like(ns)
    model = lhvi_b200.lowering.lower_ground(g, 3, 3)
    from lhvi_b200.engine import DeviceEngine
    eng = DeviceEngine(model, dtype="float64")
    assert not eng.persistent()
    eta, tau, w_tau = lhvi_b200.synthetic.random_state(model, 4)
    ref = NumpyVI(model)
    ref.eta[:], ref.tau[:], ref.w_tau = eta, tau, w_tau
    ref.refresh()
    ref.adam_step(0.1)
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    eng.iterate(1, 0.1)
    np.testing.assert_allclose(eng.get_state()[0], ref.eta, rtol=1e-9, atol=1e-11)
