"""Array-native colour passing (lifting.py) against the object-level implementation
(CompressedGraphWithObs.py, itself pinned to the reference's partitions by the goldens) and
against the ground model: same partition, same lowered model, same free energy."""
import numpy as np
import pytest

import lhvi_b200
import specs
from oracle.vi_numpy import grad_pass

lifting = lhvi_b200.lifting
syn = lhvi_b200.synthetic


def _partition(colour):
    groups = {}
    for i, c in enumerate(colour):
        groups.setdefault(int(c), []).append(i)
    return {frozenset(v) for v in groups.values()}


def _object_partition(g, rvs, split=True):
    cg = lhvi_b200.CompressedGraphWithObs.CompressedGraph(g)
    if split:
        cg.run()
    else:
        cg.init_cluster(is_split_cont_evidence=False)
        n = -1
        while n != len(cg.rvs):
            n = len(cg.rvs)
            cg.split_factors()
            cg.split_rvs()
    index = {id(rv): i for i, rv in enumerate(rvs)}
    return {frozenset(index[id(rv)] for rv in c.rvs) for c in cg.rvs}, cg


@pytest.mark.parametrize("name", sorted(specs.CASES))
@pytest.mark.parametrize("split", [True, False], ids=["exact-evidence", "lumped-evidence"])
def test_partition_matches_object_colour_passing(name, split, ns):
    builder = specs.CASES[name][0]
    g, _ = builder(ns)
    ga, rvs = lifting.arrays_from_graph(g)
    vcol, fcols, sweeps = lifting.colour_passing(ga, split_cont_evidence=split)
    want, cg = _object_partition(g, rvs, split)
    assert _partition(vcol) == want
    assert sum(len(np.unique(c)) for c in fcols) >= 1
    n_fclasses = len(np.unique(np.concatenate(fcols)))
    assert n_fclasses == len(cg.factors)


@pytest.mark.parametrize("n,t,period", [(4, 4, None), (6, 5, 2), (8, 4, 4)])
def test_kalman_arrays_twin_and_lifted_lowering(n, t, period):
    ga, state = syn.kalman_arrays(n, t, levels=2, seed=1, period=period)
    g, rvs = syn.kalman_graph(n, t, levels=2, seed=1, period=period)
    ga2, rvs2 = lifting.arrays_from_graph(g)
    assert ga.n_vars == ga2.n_vars and ga.n_factors == ga2.n_factors
    np.testing.assert_array_equal(np.isnan(ga.var_value), np.isnan(ga2.var_value))
    np.testing.assert_array_equal(np.nan_to_num(ga.var_value), np.nan_to_num(ga2.var_value))
    vcol, fcols, _ = lifting.colour_passing(ga)
    want, cg = _object_partition(g, rvs)
    assert _partition(vcol) == want
    # lowered models of the two routes: same free energy and G_w under class-tied parameters
    K, T = 2, 3
    m_arr, q = lifting.lower_lifted(ga, K, T)
    m_obj = lhvi_b200.lowering.lower_compressed(cg, K, T)
    assert m_arr.n_records == m_obj.n_records and m_arr.n_vars == m_obj.n_vars
    rng = np.random.default_rng(0)
    w = np.array([0.3, 0.7])

    def tied_state(model, rep_of_handle):
        eta = np.zeros(model.n_param)
        for h, off in zip(model.handles, model.var_off):
            r = np.random.default_rng(rep_of_handle(h))
            for k in range(K):
                eta[off + 2 * k] = r.uniform(-1.5, 1.5)
                eta[off + 2 * k + 1] = r.uniform(0.5, 2.0)
        return eta
    index = {id(rv): i for i, rv in enumerate(rvs)}
    e_arr = grad_pass(m_arr, tied_state(m_arr, lambda h: h.rep), w)
    e_obj = grad_pass(m_obj, tied_state(m_obj, lambda h: min(index[id(rv)] for rv in h.rvs)), w)
    np.testing.assert_allclose(e_arr[2], e_obj[2], rtol=1e-12)
    np.testing.assert_allclose(e_arr[1], e_obj[1], rtol=1e-12)
    np.testing.assert_allclose(np.sort(e_arr[0]), np.sort(e_obj[0]), rtol=1e-9, atol=1e-12)
    # and the lifted free energy equals the ground one when the ground parameters are tied by class
    m_gr, qg = lifting.lower_ground_arrays(ga, K, T)
    eta_g = tied_state(m_gr, lambda h: int(q.rvs[int(vcol[h.rep])].rep))
    e_gr = grad_pass(m_gr, eta_g, w)
    np.testing.assert_allclose(e_gr[2], e_arr[2], rtol=1e-10)
    np.testing.assert_allclose(e_gr[1], e_arr[1], rtol=1e-10)


def test_kalman_compression_at_scale():
    """Config 2 shape (reduced so the CPU suite stays fast): 200 state dimensions x 50 steps,
    binary observations; the partition must be equitable and much smaller than the ground graph."""
    ga, state = syn.kalman_arrays(200, 50, levels=2, seed=0, period=4)
    vcol, fcols, sweeps = lifting.colour_passing(ga)
    q = lifting.quotient(ga, vcol, fcols)
    assert ga.n_vars == 200 * 50 + 200 * 49
    # equitable: members of a class see the same multiset of factor classes
    fc_all = np.concatenate(fcols)
    off = np.cumsum([0] + [b.n for b in ga.blocks])
    sig = {}
    for bi, b in enumerate(ga.blocks):
        for a in range(b.arity):
            for v, f in zip(b.args[:, a], fc_all[off[bi]:off[bi + 1]]):
                sig.setdefault(int(v), []).append(int(f))
    by_class = {}
    for v, lst in sig.items():
        by_class.setdefault(int(vcol[v]), set()).add(tuple(sorted(lst)))
    assert all(len(s) == 1 for s in by_class.values())
    # 4 classes of states per time step; observation leaves merge by value as well
    assert len(np.unique(vcol[state.reshape(-1)])) == 4 * 50
    assert q.compression > 20
