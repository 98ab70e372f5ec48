"""Long-run and cross-check parity on the GPU (north_star acceptance criteria):

* fp64 mode follows the numpy oracle (itself pinned to the reference goldens) over 1000 Adam
  iterations, and fp32 mode ends with beliefs within 1e-4 of it;
* config-1 settings (K=2, Gauss-Hermite degree 10) on a hybrid MLN with observed relation atoms;
* config-4 cross-check: on a pairwise Gaussian grid the converged VI means equal the exact
  solution of J mu = h (the reference compares against GaBP / the exact solve in
  Demo/RGM/RGMKLDivergence.py:40-46 and Demo/RKF/LRKFDemoCycle.py:85-102)."""
import contextlib
import io

import numpy as np
import pytest

import lhvi_b200
import specs
from oracle.vi_numpy import NumpyVI, grad_pass, norm_pdf_ref

pytestmark = pytest.mark.gpu


def _beliefs(model, eta, w, xs):
    """Mixture density of every continuous variable at the probe points ``xs``."""
    K = model.K
    out = []
    for off in model.var_off[model.var_kind == 0]:
        mu, var = eta[off:off + 2 * K:2], eta[off + 1:off + 2 * K:2]
        out.append([(w * norm_pdf_ref(x, mu, var)).sum() for x in xs])
    return np.array(out)


def test_1000_iterations_fp64_trajectory_and_fp32_beliefs():
    from lhvi_b200.engine import DeviceEngine
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(60, 4, 3, 3, seed=7, weighted=True)
    eta, tau, w_tau = syn.random_state(model, 5)
    ref = NumpyVI(model)
    ref.eta[:], ref.tau[:], ref.w_tau = eta, tau, w_tau
    ref.refresh()
    for _ in range(1000):
        ref.adam_step(0.05)
    xs = np.linspace(-2.0, 9.0, 7)
    want = _beliefs(model, ref.eta, ref.w, xs)
    out = {}
    for dtype in ("float64", "float32"):
        eng = DeviceEngine(model, dtype=dtype)
        eng.set_state(eta, tau, w_tau)
        eng.reset_moments()
        eng.iterate(1000, 0.05)
        e, _, wt, w = eng.get_state()
        out[dtype] = (e, wt, _beliefs(model, e, w, xs))
    # fp64: the whole trajectory stays on the oracle's (north_star: 1e-6 relative per iteration)
    np.testing.assert_allclose(out["float64"][0], ref.eta, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(out["float64"][1], ref.w_tau, rtol=1e-6, atol=1e-8)
    # fp32: final beliefs within 1e-4 (north_star)
    np.testing.assert_allclose(out["float32"][2], want, rtol=0, atol=1e-4)


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-9), ("float32", 5e-5)])
def test_config1_settings_k2_t10(dtype, tol, ns):
    """K=2 mixtures, quadrature degree 10 (BASELINE config 1) on the hybrid MLN with observed
    relation atoms: free energy, G_w and every gradient against the oracle."""
    from lhvi_b200.engine import DeviceEngine
    g, rvs = specs.hmln_evidence(ns)
    model = lhvi_b200.lowering.lower_ground(g, 2, 10)
    eta, tau, w_tau = lhvi_b200.synthetic.random_state(model, 3)
    w = np.full(2, 0.5)
    og, ogw, oe = grad_pass(model, eta, w)
    eng = DeviceEngine(model, dtype=dtype)
    eng.set_state(eta, tau, w_tau)
    grad, g_w, energy = eng.gradients()
    np.testing.assert_allclose(energy, oe, rtol=tol)
    np.testing.assert_allclose(g_w, ogw, rtol=tol, atol=tol * np.abs(ogw).max())
    np.testing.assert_allclose(grad, og, rtol=tol, atol=tol * np.abs(og).max())


def test_gaussian_grid_means_match_exact_solve(ns):
    """With K=1 the objective is the mean-field ELBO, whose stationary means are the exact means of
    a Gaussian model: run the drop-in engine to convergence and compare with the solve."""
    n = 10
    rng = np.random.default_rng(11)
    dom = ns.Domain((-20, 20), continuous=True)
    X = [ns.RV(dom) for _ in range(n * n)]
    m_i = rng.uniform(-3.0, 3.0, size=n * n)
    s_i = rng.uniform(0.5, 2.0, size=n * n)
    sig = np.array([[1.5, 0.9], [0.9, 1.5]])
    prec = np.linalg.inv(sig)
    pe = ns.GaussianPotential([0.0, 0.0], sig.tolist())
    fs = [ns.F(ns.GaussianPotential([float(m)], [[float(s)]]), [x]) for x, m, s in zip(X, m_i, s_i)]
    J = np.diag(1.0 / s_i)
    h = m_i / s_i
    for r in range(n):
        for c in range(n):
            for (r2, c2) in ((r, c + 1), (r + 1, c)):
                if r2 < n and c2 < n:
                    i, j = r * n + c, r2 * n + c2
                    fs.append(ns.F(pe, [X[i], X[j]]))
                    J[np.ix_([i, j], [i, j])] += prec
    g = ns.Graph()
    g.rvs, g.factors = set(X), set(fs)
    g.init_nb()
    exact = np.linalg.solve(J, h)

    vi = lhvi_b200.VarInference.VarInference(g, 1, 3)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        vi.run(1500, lr=0.05, is_log=False)
    got = np.array([vi.map(x) for x in X])
    np.testing.assert_allclose(got, exact, atol=2e-3)
    # the lifted engine on the same graph agrees (nothing to lift here, but the path differs)
    lvi = lhvi_b200.LiftedVarInference.VarInference(g, 1, 3)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        lvi.run(1500, lr=0.05, is_log=False)
    np.testing.assert_allclose(np.array([lvi.map(x) for x in X]), exact, atol=2e-3)
