"""BASELINE config 2 on the GPU: relational Kalman filter, ``LiftedVarInference`` over the
array-native colour passing (lifting.py).  The lifted run must follow the ground run of the same
model step by step, and with K=1 the converged means are the exact Kalman posterior means
(the reference checks its RKF demos the same way, Demo/RKF/LRKFDemoCycle.py:85-102)."""
import numpy as np
import pytest

import lhvi_b200

pytestmark = pytest.mark.gpu

lifting = lhvi_b200.lifting
syn = lhvi_b200.synthetic


def _gaussian_system(ga):
    """Precision matrix and potential vector of the hidden variables of an all-Gaussian
    ``GroundArrays`` model, from the potentials' own quadratic parameters."""
    hidden = np.flatnonzero(np.isnan(ga.var_value))
    pos = {int(v): i for i, v in enumerate(hidden)}
    J = np.zeros((hidden.size, hidden.size))
    h = np.zeros(hidden.size)
    for b in ga.blocks:
        A, lin, _ = b.potential.get_quadratic_params()          # log psi = x^T A x + lin^T x + c
        A = np.asarray(A, dtype=float)
        lin = np.asarray(lin, dtype=float).reshape(-1)
        S = A + A.T                                              # d/dx (x^T A x) = S x
        for row in b.args:
            hid = [a for a, v in enumerate(row) if int(v) in pos]
            for a in hid:
                i = pos[int(row[a])]
                h[i] += lin[a]
                for c, v in enumerate(row):
                    if int(v) in pos:
                        J[i, pos[int(v)]] -= S[a, c]
                    else:
                        h[i] += S[a, c] * ga.var_value[int(v)]
    return hidden, J, h


def test_config2_lifted_follows_ground_and_reaches_the_exact_means():
    ga, state = syn.kalman_arrays(12, 6, levels=2, seed=3, period=3)
    lifted = lifting.ArrayVI(ga, 1, 3, lifted=True, dtype="float64")
    ground = lifting.ArrayVI(ga, 1, 3, lifted=False, dtype="float64")
    assert lifted.quotient.compression > 2.5
    ground.tie_to(lifted)
    lifted.init_param(0)
    np.testing.assert_allclose(lifted.free_energy(), ground.free_energy(), rtol=1e-10)
    for _ in range(3):
        lifted.run(50, 0.05)
        ground.run(50, 0.05)
        pl, _ = lifted.ground_params()
        pg, _ = ground.ground_params()
        for v in pg:
            np.testing.assert_allclose(pl[v], pg[v], rtol=1e-7, atol=1e-9)
    lifted.run(2500, 0.02)
    hidden, J, h = _gaussian_system(ga)
    exact = np.linalg.solve(J, h)
    pl, _ = lifted.ground_params()
    got = np.array([pl[int(v)][0, 0] for v in hidden])
    np.testing.assert_allclose(got, exact, atol=3e-3)


def test_config2_shape_runs_lifted_in_fp32():
    """1000 state dimensions x 100 steps (199 000 variables, 592 000 ground factors) compress 100x;
    the lifted fp32 engine lowers the free energy monotonically at a small step."""
    ga, state = syn.kalman_arrays(1000, 100, levels=2, seed=0, period=10)
    vi = lifting.ArrayVI(ga, 1, 3, lifted=True, dtype="float32")
    assert vi.quotient.compression > 50
    fe0 = vi.free_energy()
    vi.run(200, 0.02)
    fe1 = vi.free_energy()
    assert np.isfinite(fe1) and fe1 < fe0


def test_config5_c2f_on_arrays_matches_the_oracle_engine_and_scales():
    """C2FArrayVI on the device engine: (a) small model, fp64: same result as the same host logic
    over the numpy oracle; (b) 20 000 entities x 5 groups (100 000 link factors), fp32: runs its
    refinement rounds, the number of classes grows, the free energy stays finite."""
    from oracle_engine import OracleEngine
    ga = syn.relational_hybrid_arrays(60, 3, seed=2)
    dev = lifting.C2FArrayVI(ga, 2, 3, dtype="float64")
    dev.run(30, 0.05)
    ref = lifting.C2FArrayVI(ga, 2, 3, engine_factory=lambda m: OracleEngine(m, var_threshold=0.1))
    ref.run(30, 0.05)
    np.testing.assert_array_equal(dev.vcol, ref.vcol)
    pd, wd = dev.ground_params()
    pr, wr = ref.ground_params()
    for v in pr:
        np.testing.assert_allclose(pd[v], pr[v], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(dev.free_energy(), ref.free_energy(), rtol=1e-9)

    big = syn.relational_hybrid_arrays(20_000, 5, seed=0)
    vi = lifting.C2FArrayVI(big, 3, 3, dtype="float32")
    vi.run(40, 0.05)
    classes = [n for n, _ in vi.history]
    assert classes == sorted(classes) and classes[-1] > classes[0]
    assert np.isfinite(vi.free_energy())
