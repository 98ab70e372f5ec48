"""Array-native benchmark generators against their object-graph twins lowered through the
plugin API (same potentials, same evidence): free energy and every gradient must agree."""
import numpy as np
import pytest

import lhvi_b200
from oracle.vi_numpy import grad_pass


def _compare(model_a, model_b, seed=3):
    syn = lhvi_b200.synthetic
    assert model_a.n_param == model_b.n_param
    np.testing.assert_array_equal(model_a.var_off, model_b.var_off)
    eta, _, _ = syn.random_state(model_a, seed)
    K = model_a.K
    w = np.full(K, 1.0 / K)
    ga, gwa, ea = grad_pass(model_a, eta, w)
    gb, gwb, eb = grad_pass(model_b, eta, w)
    np.testing.assert_allclose(ea, eb, rtol=1e-11)
    np.testing.assert_allclose(gwa, gwb, rtol=1e-11)
    np.testing.assert_allclose(ga, gb, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("order", ["hub", "entity"])
@pytest.mark.parametrize("weighted", [False, True])
def test_relational_hybrid_twin(order, weighted):
    syn, low = lhvi_b200.synthetic, lhvi_b200.lowering
    P, G, K, T = 23, 4, 3, 3
    fast = syn.relational_hybrid(P, G, K, T, seed=5, order=order, weighted=weighted)
    g, topics, entities = syn.relational_hybrid_graph(P, G, seed=5)
    slow = low.lower_ground(g, K, T)
    assert fast.n_records == slow.n_records == P * G + P + G * (G - 1)
    _compare(fast, slow)


def test_gaussian_grid_twin():
    syn, low = lhvi_b200.synthetic, lhvi_b200.lowering
    n, K, T = 5, 2, 3
    fast = syn.gaussian_grid(n, K, T)
    g, X = syn.gaussian_grid_graph(n)
    slow = low.lower_ground(g, K, T)
    assert fast.n_records == slow.n_records
    _compare(fast, slow)


def test_shard_partitions_records():
    syn = lhvi_b200.synthetic
    m = syn.relational_hybrid(50, 5, 2, 3, seed=1)
    eta, _, _ = syn.random_state(m, 0)
    w = np.array([0.4, 0.6])
    full = grad_pass(m, eta, w)
    parts = [grad_pass(m.shard(r, 3), eta, w) for r in range(3)]
    assert sum(p.n_records for p in [m.shard(r, 3) for r in range(3)]) == m.n_records
    np.testing.assert_allclose(sum(p[0] for p in parts), full[0], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(sum(p[1] for p in parts), full[1], rtol=1e-11)
    np.testing.assert_allclose(sum(p[2] for p in parts), full[2], rtol=1e-11)


def test_fold_unary_matches_the_coefficient_blocks():
    """lowering.fold_unary: c0 + l0 x + a0 x^2 equals the ptab quadratic with the record's point
    evidence substituted, for every pure one-argument group of the relational model."""
    syn, low = lhvi_b200.synthetic, lhvi_b200.lowering
    model = syn.relational_hybrid(40, 3, 2, 3, seed=9, weighted=True)
    rng = np.random.default_rng(0)
    seen = 0
    for g in model.groups:
        fold = low.fold_unary(g, model.ptab)
        if g.node or not g.pure or g.nc != 1:
            assert fold is None
            continue
        seen += 1
        assert fold.shape == (3, g.n)
        ncoef = low.ncoef_for(1 + g.ne)
        for r in rng.choice(g.n, size=min(g.n, 25), replace=False):
            coef = model.ptab[g.pot[r]:g.pot[r] + ncoef]
            for x in (-2.3, 0.0, 1.7):
                want = low.eval_quadratic(coef, [x] + [g.ecval[e, r] for e in range(g.ne)])
                got = fold[0, r] + fold[1, r] * x + fold[2, r] * x * x
                assert abs(want - got) <= 1e-12 * max(1.0, abs(want))
    assert seen >= 2
