"""Owner-computes partition (``dist.partition``): structural invariants, and a simulated
multi-rank Adam run (numpy oracle standing in for the kernels) that exchanges only
``[G_w | energy | shared-variable gradients]`` and must reproduce the single-process run."""
import numpy as np
import pytest

import lhvi_b200
from oracle.vi_numpy import NumpyVI, grad_pass, tau_gradients


def _models():
    syn = lhvi_b200.synthetic
    return {
        "relational_hub": syn.relational_hybrid(120, 4, 2, 3, seed=3, order="hub", weighted=True),
        "relational_entity": syn.relational_hybrid(90, 3, 3, 3, seed=4, order="entity"),
        "grid": syn.gaussian_grid(9, 2, 3),
    }


@pytest.mark.parametrize("name", ["relational_hub", "relational_entity", "grid"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_partition_invariants(name, world):
    from lhvi_b200 import dist as D
    model = _models()[name]
    part = D.partition(model, world)
    lut = D._var_of_slot(model)
    # every record on exactly one rank
    for g, rr in zip(model.groups, part.rec_rank):
        assert rr.shape == (g.n,) and rr.min() >= 0 and rr.max() < world
    locals_ = [D.local_model(model, part, r) for r in range(world)]
    assert sum(m.n_records for m in locals_) == model.n_records
    assert sum(m.n_node_records for m in locals_) == model.n_node_records
    # a rank's records touch only variables it owns or shared ones, and it steps exactly those
    for r, m in enumerate(locals_):
        allowed = (part.var_owner == r) | (part.var_owner < 0)
        for g in m.groups:
            for a in range(g.nh):
                assert allowed[lut[g.poff[a]]].all()
        assert set(m.var_off.tolist()) == set(model.var_off[allowed].tolist())
    # owned variables are owned once; together with the shared ones they cover everything
    assert ((part.var_owner >= -1) & (part.var_owner < world)).all()


def test_relational_model_shares_only_the_group_variables():
    from lhvi_b200 import dist as D
    syn = lhvi_b200.synthetic
    G = 5
    model = syn.relational_hybrid(4000, G, 3, 3, seed=0, order="hub", weighted=True)
    part = D.partition(model, 8)
    assert part.shared.tolist() == list(range(G))            # the G hub variables come first
    counts = np.zeros(8)
    for g, rr in zip(model.groups, part.rec_rank):
        if not g.node:
            counts += np.bincount(rr, minlength=8)
    assert counts.max() / counts.mean() < 1.1               # balanced


def test_grid_shares_only_block_boundaries():
    from lhvi_b200 import dist as D
    n = 24
    model = lhvi_b200.synthetic.gaussian_grid(n, 1, 3)
    part = D.partition(model, 4)
    assert 0 < part.shared.size <= 3 * (n + 2)


@pytest.mark.parametrize("name", ["relational_hub", "grid"])
@pytest.mark.parametrize("world", [2, 5])
def test_simulated_ranks_match_single_process(name, world):
    from lhvi_b200 import dist as D
    syn = lhvi_b200.synthetic
    model = _models()[name]
    K = model.K
    eta, tau, w_tau = syn.random_state(model, 11)

    ref = NumpyVI(model)
    ref.eta[:], ref.tau[:], ref.w_tau = eta, tau, w_tau
    ref.refresh()

    part = D.partition(model, world)
    shared = D.slot_elements(model, part.shared)
    ranks = []
    for r in range(world):
        vi = NumpyVI(model)
        vi.eta[:], vi.tau[:], vi.w_tau = eta, tau, w_tau
        vi.refresh()
        keep = D.slot_elements(model, np.flatnonzero(part.var_owner == r))
        ranks.append((vi, D.local_model(model, part, r), keep))

    energies = []
    for _ in range(3):
        energies.append(ref.adam_step(0.1))
        outs = [grad_pass(m, vi.eta, vi.w) for vi, m, _ in ranks]
        x = sum(np.concatenate([gw, [e], g[shared]]) for g, gw, e in outs)      # the exchange
        for (vi, m, _), (g, gw, e) in zip(ranks, outs):
            g = g.copy()
            g[shared] = x[K + 1:]
            vi.gradients = lambda g=g, vi=vi: (*tau_gradients(model, g, x[:K], vi.eta, vi.w), x[K])
            assert np.isclose(vi.adam_step(0.1), energies[-1], rtol=1e-12)

    merged = np.zeros_like(ref.eta)
    for vi, _, keep in ranks:
        merged[keep] = vi.eta[keep]
    merged[shared] = ranks[0][0].eta[shared]
    np.testing.assert_allclose(merged, ref.eta, rtol=1e-10, atol=1e-12)
    for vi, _, _ in ranks:
        np.testing.assert_allclose(vi.w_tau, ref.w_tau, rtol=1e-10, atol=1e-12)
        np.testing.assert_array_equal(vi.eta[shared], ranks[0][0].eta[shared])
