"""the image-denoising potentials (reference ``Potential.py:400-424``).

``ImageNodePotential`` is a Gaussian bump in ``x0 - x1`` -- exp-quadratic, so the lowering takes it like
any other potential (quadratic fit of ``log get`` verified on probes) and the kernels' table form
reproduces the loop oracle, which calls ``potential.get`` at every grid point.  ``ImageEdgePotential``
(truncated-Laplacian smoothness prior) and ``MLNHardPotential`` over continuous arguments are not
log-quadratic: the lowering refuses them loudly instead of approximating (the reference uses them with
its belief-propagation baselines and MaxWalkSAT only)."""
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


def test_non_quadratic_potentials_are_refused(ns):
    d = ns.Domain((-5, 5), continuous=True)
    a, b = ns.RV(d), ns.RV(d)
    for pot in (ns.ImageEdgePotential(0.1, 1.0, 2.0), ns.MLNHardPotential(lambda v: v[0] - v[1])):
        g = ns.Graph()
        g.rvs, g.factors = {a, b}, {ns.F(pot, [a, b])}
        g.init_nb()
        with pytest.raises(NotImplementedError):
            lhvi_b200.lowering.lower_ground(g, 2, 3)
