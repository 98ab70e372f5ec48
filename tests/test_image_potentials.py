"""the image-denoising potentials (reference ``Potential.py:400-424``).

``ImageNodePotential`` is a Gaussian bump in ``x0 - x1`` -- exp-quadratic, so the lowering takes it like
any other potential (quadratic fit of ``log get`` verified on probes) and the kernels' table form
reproduces the loop oracle, which calls ``potential.get`` at every grid point.  ``ImageEdgePotential``
(truncated-Laplacian smoothness prior) and ``MLNHardPotential`` over continuous arguments are not
log-quadratic: their groups carry a potential kind and the generic kernel evaluates them point by point
(``tests/test_gpu_parity.py`` runs the ``denoise`` and ``hard_mln`` goldens of the unmodified reference on
the device); a formula that is not quadratic in its continuous arguments is refused loudly."""
import numpy as np
import pytest

import lhvi_b200
from oracle.vi_loops import LoopOracle
from oracle.vi_numpy import NumpyVI, tau_gradients


def _denoising_chain(ns, n=5):
    rng = np.random.default_rng(4)
    d = ns.Domain((-5, 5), continuous=True)
    x = [ns.RV(d) for _ in range(n)]
    y = [ns.RV(d, float(v)) for v in rng.uniform(-2, 2, size=n)]
    fs = [ns.F(ns.ImageNodePotential(0.2, 0.7), [x[i], y[i]]) for i in range(n)]
    smooth = ns.GaussianPotential([0.0, 0.0], [[1.0, 0.7], [0.7, 1.0]])
    fs += [ns.F(smooth, [x[i], x[i + 1]]) for i in range(n - 1)]
    g = ns.Graph()
    g.rvs, g.factors = set(x + y), set(fs)
    g.init_nb()
    return g, x


def test_image_node_potential_lowers_to_the_table_form(ns):
    g, x = _denoising_chain(ns)
    K, T = 2, 3
    model = lhvi_b200.lowering.lower_ground(g, K, T)
    eta, tau, w_tau = lhvi_b200.synthetic.random_state(model, 2)
    vi = NumpyVI(model)
    vi.eta[:], vi.tau[:], vi.w_tau = eta, tau, w_tau
    vi.refresh()
    g_flat, g_wtau, energy = vi.gradients()
    loop = LoopOracle(sorted(g.rvs, key=lambda r: r.id), sorted(g.factors, key=lambda f: id(f)), K, T)
    cont = {rv: eta[model.var_off[i]:model.var_off[i] + 2 * K].reshape(K, 2) for rv, i in model.index.items()}
    loop.set_params(w_tau, cont, {})
    np.testing.assert_allclose(energy, loop.free_energy(), rtol=1e-9)
    np.testing.assert_allclose(g_wtau, loop.gradient_w_tau(), rtol=1e-9, atol=1e-12)
    for rv, i in model.index.items():
        off = int(model.var_off[i])
        np.testing.assert_allclose(g_flat[off:off + 2 * K].reshape(K, 2), loop.gradient_mu_var(rv), rtol=1e-9, atol=1e-11)


def _check_against_loops(g, K, T, seed):
    model = lhvi_b200.lowering.lower_ground(g, K, T)
    eta, tau, w_tau = lhvi_b200.synthetic.random_state(model, seed)
    vi = NumpyVI(model)
    vi.eta[:], vi.tau[:], vi.w_tau = eta, tau, w_tau
    vi.refresh()
    g_flat, g_wtau, energy = vi.gradients()
    loop = LoopOracle(sorted(g.rvs, key=lambda r: r.id), sorted(g.factors, key=lambda f: id(f)), K, T)
    cont = {rv: eta[model.var_off[i]:model.var_off[i] + 2 * K].reshape(K, 2)
            for rv, i in model.index.items() if rv.domain.continuous}
    loop.set_params(w_tau, cont, {})
    np.testing.assert_allclose(energy, loop.free_energy(), rtol=1e-9)
    np.testing.assert_allclose(g_wtau, loop.gradient_w_tau(), rtol=1e-9, atol=1e-12)
    for rv, i in model.index.items():
        off = int(model.var_off[i])
        np.testing.assert_allclose(g_flat[off:off + 2 * K].reshape(K, 2), loop.gradient_mu_var(rv), rtol=1e-9, atol=1e-10)
    return model


def test_point_by_point_potentials_lower_with_their_kind(ns):
    """``ImageEdgePotential`` and ``MLNHardPotential`` over continuous arguments are not log-quadratic:
    their groups carry ``kind`` (LHVI_POT_IMAGE_EDGE / LHVI_POT_HARD) and the table-form oracle, reading the
    same coefficient blocks the device reads, agrees with the loop oracle that calls ``potential.get``."""
    d = ns.Domain((-5, 5), continuous=True)
    a, b, c = ns.RV(d), ns.RV(d), ns.RV(d, 0.8)
    edge = ns.ImageEdgePotential(0.1, 1.0, 2.0)
    hard = ns.MLNHardPotential(lambda v: v[0] - v[1] + 0.2)
    x2 = ns.X2Potential(1.0, 2.0)
    g = ns.Graph()
    g.rvs = {a, b, c}
    g.factors = {ns.F(edge, [a, b]), ns.F(edge, [b, c]), ns.F(hard, [a, b]), ns.F(hard, [c, a]), ns.F(x2, [a]), ns.F(x2, [b])}
    g.init_nb()
    model = _check_against_loops(g, 2, 3, 5)
    kinds = sorted({(grp.kind, grp.pure) for grp in model.groups if not grp.node})
    L = lhvi_b200.lowering
    assert (L.POT_IMAGE_EDGE, False) in kinds and (L.POT_IMAGE_EDGE, True) in kinds
    assert (L.POT_HARD, False) in kinds and (L.POT_HARD, True) in kinds
    for grp in model.groups:                      # the streaming / run-major forms are for quadratics only
        if grp.kind != L.POT_QUADRATIC:
            assert L.fold_unary(grp, model.ptab) is None


def test_formulas_that_are_not_quadratic_are_still_refused(ns):
    d = ns.Domain((-5, 5), continuous=True)
    a, b = ns.RV(d), ns.RV(d)
    for pot in (ns.MLNHardPotential(lambda v: v[0] ** 3 - v[1]), ns.MLNPotential(lambda v: abs(v[0] - v[1]))):
        g = ns.Graph()
        g.rvs, g.factors = {a, b}, {ns.F(pot, [a, b])}
        g.init_nb()
        with pytest.raises(NotImplementedError):
            lhvi_b200.lowering.lower_ground(g, 2, 3)
