"""Gaussian belief propagation: the information-form lowering and the oracle
against the unmodified reference's ``GaBP`` (``tests/golden/gabp_grid.json``) and against the exact
posterior; the device sweeps against both (``-m gpu``)."""
import json
import os

import numpy as np
import pytest

import lhvi_b200
import specs
from lhvi_b200 import GaBP as gabp
from oracle import gabp_numpy

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "gabp_grid.json")))


def _arrays(ns):
    g, rvs = specs.gabp_grid(ns)
    model = lhvi_b200.lowering.lower_ground(g, 1, 3)
    hidden = [rv for rv in rvs if rv.value is None]
    assert [model.index[rv] for rv in hidden] == list(range(len(hidden)))      # golden order = slot order
    return g, hidden, gabp.gabp_from_model(model)


def _dense(a):
    """Joint precision and potential of the information form (every factor summed)."""
    J = np.diag(a.jd.copy())
    h = a.hd.copy()
    for e in range(a.src.size):
        s, d = a.src[e], a.dst[e]
        jss, jdd, jc, hs, hd = a.coef[:, e]
        if e < a.rev[e]:                       # each factor once
            J[s, s] += jss
            J[d, d] += jdd
            J[s, d] += jc
            J[d, s] += jc
            h[s] += hs
            h[d] += hd
    return J, h


@pytest.mark.parametrize("n", GOLD["iterations"])
def test_oracle_matches_the_reference_after_n_iterations(ns, n):
    _, _, a = _arrays(ns)
    mean, var = gabp_numpy.run(a, n)
    np.testing.assert_allclose(mean, GOLD["mean"][str(n)], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(var, GOLD["var"][str(n)], rtol=1e-10)


def test_converged_means_are_the_exact_posterior_means(ns):
    _, _, a = _arrays(ns)
    J, h = _dense(a)
    exact = np.linalg.solve(J, h)
    mean, _ = gabp_numpy.run(a, 200)
    np.testing.assert_allclose(mean, exact, rtol=1e-9, atol=1e-11)


def test_information_form_from_the_array_generators():
    """The config-4 grid of the bench (array-native model): GaBP's fixed point solves J mu = h."""
    from tools.bench_configs import grid_with_observations
    model, J, h = grid_with_observations(12, 1, 3)
    a = gabp.gabp_from_model(model)
    Jd, hd = _dense(a)
    np.testing.assert_allclose(Jd, J.toarray(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(hd, h, rtol=1e-12, atol=1e-12)
    mean, _ = gabp_numpy.run(a, 400)
    np.testing.assert_allclose(mean, np.linalg.solve(Jd, hd), rtol=1e-8, atol=1e-10)


def _random_tree(ns, n, seed):
    """Random tree of hidden reals with LinearGaussian / Gaussian / XY edges, an X2 prior on every node and
    a few observed leaves: Gaussian belief propagation is exact on it (means and variances)."""
    rng = np.random.default_rng(seed)
    dom = ns.Domain((-10, 10), continuous=True)
    x = [ns.RV(dom) for _ in range(n)]
    fs = [ns.F(ns.X2Potential(1.0, float(rng.uniform(1.0, 3.0))), [v]) for v in x]
    for i in range(1, n):
        j = int(rng.integers(0, i))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            pot = ns.LinearGaussianPotential(float(rng.uniform(0.3, 1.2)), float(rng.uniform(0.5, 2.0)))
        elif kind == 1:
            c = float(rng.uniform(-0.6, 0.6))
            pot = ns.GaussianPotential([0.0, 0.0], [[1.5, c], [c, 1.2]])
        else:
            pot = ns.XYPotential(float(rng.uniform(-0.4, 0.4)), 2.0)
        fs.append(ns.F(pot, [x[j], x[i]]))
    obs = []
    for i in rng.choice(n, size=max(1, n // 4), replace=False):
        y = ns.RV(dom, float(rng.uniform(-2, 2)))
        obs.append(y)
        fs.append(ns.F(ns.LinearGaussianPotential(0.9, 0.7), [x[int(i)], y]))
    g = ns.Graph()
    g.rvs, g.factors = set(x + obs), set(fs)
    g.init_nb()
    return g, x


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_exact_marginals_on_trees(ns, seed):
    g, x = _random_tree(ns, 25, seed)
    model = lhvi_b200.lowering.lower_ground(g, 1, 3)
    a = gabp.gabp_from_model(model)
    J, h = _dense(a)
    cov = np.linalg.inv(J)
    mean, var = gabp_numpy.run(a, 60)
    np.testing.assert_allclose(mean, cov @ h, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, np.diag(cov), rtol=1e-9)


@pytest.mark.gpu
def test_device_gabp_is_exact_on_trees(ns):
    for seed in range(3):
        g, x = _random_tree(ns, 40, seed)
        bp = gabp.GaBP(g).run(80)
        J, h = _dense(bp.arrays)
        cov = np.linalg.inv(J)
        got = np.array([bp.get_belief_params(rv) for rv in x])
        order = [bp.model.index[rv] for rv in x]
        np.testing.assert_allclose(got[:, 0], (cov @ h)[order], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(got[:, 1], np.diag(cov)[order], rtol=1e-9)


def test_unsupported_models_are_refused(ns):
    g, _ = specs.chain_table(ns)
    with pytest.raises(NotImplementedError):
        gabp.gabp_from_model(lhvi_b200.lowering.lower_ground(g, 1, 3))
    g, _ = specs.denoise(ns)
    with pytest.raises(NotImplementedError):
        gabp.gabp_from_model(lhvi_b200.lowering.lower_ground(g, 1, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,rtol", [("float64", 1e-10), ("float32", 2e-5)])
def test_device_sweeps_match_the_reference(ns, dtype, rtol):
    g, hidden, a = _arrays(ns)
    for n in GOLD["iterations"]:
        bp = gabp.GaBP(g, dtype=dtype).run(n)
        got = np.array([bp.get_belief_params(rv) for rv in hidden])
        np.testing.assert_allclose(got[:, 0], GOLD["mean"][str(n)], rtol=rtol, atol=rtol)
        np.testing.assert_allclose(got[:, 1], GOLD["var"][str(n)], rtol=rtol)
    assert bp.map(hidden[0]) == got[0, 0]
    assert bp.belief(0.3, hidden[1]) == pytest.approx(gabp.GaBP.norm_pdf(0.3, got[1, 0], got[1, 1]))
    seen = next(rv for rv in g.rvs if rv.value is not None)
    assert bp.map(seen) == seen.value and bp.belief(seen.value, seen) == 1 and bp.belief(seen.value + 1, seen) == 0


@pytest.mark.gpu
def test_device_gabp_and_vi_agree_on_the_config4_grid():
    """BASELINE config 4 at 200 x 200: GaBP means on the device = exact solve = the K=1 VI means."""
    import scipy.sparse.linalg as spla
    from lhvi_b200.engine import DeviceEngine
    from tools.bench_configs import grid_with_observations
    model, J, h = grid_with_observations(200, 1, 3)
    exact, info = spla.cg(J, h, rtol=1e-12, maxiter=5000)
    assert info == 0
    bp = gabp.DeviceGaBP(gabp.gabp_from_model(model), "float64")
    bp.sweeps(300)
    mean, _ = bp.marginals()
    np.testing.assert_allclose(mean.cpu().numpy(), exact, rtol=1e-8, atol=1e-9)
    eta, tau, w_tau = lhvi_b200.synthetic.random_state(model, 0)
    eng = DeviceEngine(model, dtype="float64")
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    eng.iterate(3000, 0.05)
    mu = eng.get_state()[0][model.var_off]
    eng.close()
    assert np.abs(mu - exact).max() < 2e-2          # Adam at lr 0.05 after 3000 iterations (bench: 8e-3 at 1000 x 1000)
