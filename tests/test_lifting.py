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
@pytest.mark.parametrize("native", [True, False], ids=["c++", "numpy"])
def test_partition_matches_object_colour_passing(name, split, native, ns):
    builder = specs.CASES[name][0]
    g, _ = builder(ns)
    ga, rvs = lifting.arrays_from_graph(g)
    vcol, fcols, sweeps = lifting.colour_passing(ga, split_cont_evidence=split, use_native=native)
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


# ---- coarse-to-fine on arrays against the object-level drop-in class --------------------------

def _c2f_pair(builder_name, ns, K, T, iterations, lr):
    """Run C2FVarInference (object route, oracle engine as device double) and C2FArrayVI (array
    route, same double) from corresponding initial points; return both."""
    import contextlib
    import io
    from oracle_engine import OracleEngine, use_oracle_engine
    builder = specs.CASES[builder_name][0]
    g, _ = builder(ns)
    ga, rvs = lifting.arrays_from_graph(g)
    index = {id(rv): i for i, rv in enumerate(rvs)}

    def table_for(rep, cont, dim):
        r = np.random.default_rng([7, rep])
        if cont:
            t = np.ones((K, 2))
            t[:, 0] = r.random(K) * 3 - 1.5
            return t
        return r.random((K, dim)) * 10

    arr = lifting.C2FArrayVI(ga, K, T, engine_factory=lambda m: OracleEngine(m, var_threshold=0.1), init_fn=table_for)
    arr.run(iterations, lr)

    vi = use_oracle_engine(lhvi_b200.C2FVarInference.VarInference(g, K, T))

    def init_param():
        vi.w_tau = np.zeros(K)
        vi.eta, vi.eta_tau = {}, {}
        for h in sorted(vi.g.rvs):
            if h.value is not None:
                continue
            rep = min(index[id(rv)] for rv in h.rvs)
            if h.domain.continuous:
                vi.eta[h] = table_for(rep, True, 2)
            else:
                vi.eta_tau[h] = table_for(rep, False, len(h.domain.values))
        e = np.e ** vi.w_tau
        vi.w = e / e.sum()
        for h, t in vi.eta_tau.items():
            ex = np.e ** t
            vi.eta[h] = ex / ex.sum(axis=1, keepdims=True)
    vi.init_param = init_param
    with contextlib.redirect_stdout(io.StringIO()):
        vi.run(iterations, lr=lr, is_log=False)
    return arr, vi, rvs, index


@pytest.mark.parametrize("name", ["rgm_split", "hmln_evidence", "chain_table"])
def test_c2f_arrays_follow_the_object_route(name, ns):
    K, T = 2, 3
    arr, vi, rvs, index = _c2f_pair(name, ns, K, T, 30, 0.05)
    # same final partition
    want = {frozenset(index[id(rv)] for rv in c.rvs) for c in vi.g.rvs}
    assert _partition(arr.vcol) == want
    # same parameters of every hidden ground variable, same mixture weights
    got, w = arr.ground_params()
    vi._pull()
    for rv in rvs:
        if rv.value is None:
            np.testing.assert_allclose(got[index[id(rv)]], vi.eta[rv.cluster], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(arr.w_tau, vi.w_tau, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(arr.free_energy(), vi.free_energy(), rtol=1e-9)


def test_relational_arrays_match_the_object_twin():
    """``relational_hybrid_arrays`` describes the same ground graph as ``relational_hybrid_graph``:
    same partition under colour passing, same lifted free energy as the ground array model
    ``relational_hybrid`` under class-tied parameters."""
    P, G, K, T = 40, 3, 2, 3
    ga = syn.relational_hybrid_arrays(P, G, seed=4)
    g, topics, entities = syn.relational_hybrid_graph(P, G, seed=4)
    ga2, rvs = lifting.arrays_from_graph(g)
    assert ga.n_vars == ga2.n_vars and ga.n_factors == ga2.n_factors
    v1, _, _ = lifting.colour_passing(ga)
    v2, _, _ = lifting.colour_passing(ga2)
    assert len(np.unique(v1)) == len(np.unique(v2))
    # lifted vs ground free energy (ground: the record-level generator)
    m_l, q = lifting.lower_lifted(ga, K, T)
    ground = syn.relational_hybrid(P, G, K, T, seed=4)
    eta_g, _, _ = syn.random_state(ground, 0)
    # tie the ground parameters by class: ground slots are topics then hidden entities
    hidden_idx = np.flatnonzero(np.isnan(ga.var_value))
    assert hidden_idx.size == ground.n_vars
    rep = np.array([q.rvs[int(v1[v])].rep for v in hidden_idx])
    pos = {int(v): i for i, v in enumerate(hidden_idx)}
    for i, r in enumerate(rep):
        o, orr = ground.var_off[i], ground.var_off[pos[int(r)]]
        eta_g[o:o + 2 * K] = eta_g[orr:orr + 2 * K]
    eta_l = np.zeros(m_l.n_param)
    for h, off in zip(m_l.handles, m_l.var_off):
        o = ground.var_off[pos[int(h.rep)]]
        eta_l[off:off + 2 * K] = eta_g[o:o + 2 * K]
    w = np.array([0.4, 0.6])
    e_g = grad_pass(ground, eta_g, w)
    e_l = grad_pass(m_l, eta_l, w)
    np.testing.assert_allclose(e_l[2], e_g[2], rtol=1e-10)
    np.testing.assert_allclose(e_l[1], e_g[1], rtol=1e-10)


# ---- array-native lowering (lower_partition) against the object-level lowering -----------------

def _assert_same_model(got, want, tag):
    assert got.n_param == want.n_param, tag
    for f in ("var_kind", "var_dim", "var_off", "ptab"):
        x, y = getattr(got, f), getattr(want, f)
        assert x.dtype == y.dtype, (tag, f)
        np.testing.assert_array_equal(x, y, err_msg=f"{tag} {f}")
    assert [g.signature for g in got.groups] == [g.signature for g in want.groups], tag
    for a, b in zip(got.groups, want.groups):
        for f in ("pot", "poff", "egval", "egvar", "ecval", "wf", "gam", "nscale"):
            x, y = getattr(a, f), getattr(b, f)
            assert x.shape == y.shape and x.dtype == y.dtype, (tag, a.signature, f)
            np.testing.assert_array_equal(x, y, err_msg=f"{tag} {a.signature} {f}")
    np.testing.assert_array_equal(got.slot_class, [h.uid for h in want.handles])


def _check_array_lowering(ga, tag):
    for split in (True, False):
        vcol, fcols, _ = lifting.colour_passing(ga, split_cont_evidence=split)
        ev = np.flatnonzero(~np.isnan(ga.var_value))
        centroids = {int(c): 0.25 + 0.5 * i for i, c in enumerate(np.unique(vcol[ev])[:2])}
        for gobs, ev_value, K in ((False, None, 1), (True, None, 3), (True, centroids, 2)):
            q = lifting.quotient(ga, vcol, fcols, ev_value)
            want = lhvi_b200.lowering.lower_compressed(q, K, 3, gaussian_obs=gobs, min_obs_var=0.0)
            got = lifting.lower_partition(ga, vcol, fcols, K, 3, ev_value=ev_value, gaussian_obs=gobs)
            _assert_same_model(got, want, f"{tag} split={split} gaussian_obs={gobs} K={K}")
    want, _ = lifting.lower_ground_arrays(ga, 2, 3)
    off = np.cumsum([0] + [b.n for b in ga.blocks])
    trivial = [np.arange(b.n, dtype=np.int64) + o for b, o in zip(ga.blocks, off)]
    _assert_same_model(lifting.lower_partition(ga, np.arange(ga.n_vars), trivial, 2, 3), want, f"{tag} ground")


@pytest.mark.parametrize("name", sorted(specs.CASES))
def test_array_lowering_equals_object_lowering(name, ns):
    """Every column of every record group, the coefficient table and the slot layout: the
    partition lowered on arrays is the model ``lower_compressed`` builds from class objects --
    exact and lumped evidence, Gaussian-evidence mode, k-means centroids as values, and the
    trivial (ground) partition."""
    g, _ = specs.CASES[name][0](ns)
    ga, _ = lifting.arrays_from_graph(g)
    if name in ("hmln_hidden", "robot_like", "smokers", "tri_table3", "chain_table"):
        # discrete evidence: centroids only make sense for continuous observations
        vcol, fcols, _ = lifting.colour_passing(ga)
        for gobs, K in ((False, 2), (True, 3)):
            want = lhvi_b200.lowering.lower_compressed(lifting.quotient(ga, vcol, fcols), K, 3, gaussian_obs=gobs)
            _assert_same_model(lifting.lower_partition(ga, vcol, fcols, K, 3, gaussian_obs=gobs), want, name)
    else:
        _check_array_lowering(ga, name)


def test_array_lowering_on_generated_models():
    _check_array_lowering(syn.kalman_arrays(6, 5, levels=2, seed=1, period=2)[0], "kalman")
    _check_array_lowering(syn.relational_hybrid_arrays(40, 3, seed=1), "relational")


def test_array_vi_lifted_follows_ground_on_the_oracle_engine():
    """``ArrayVI`` host logic without a GPU (the numpy oracle as the device double): a lifted and
    a ground run of the relational Kalman filter from corresponding starts stay together."""
    from oracle_engine import OracleEngine
    factory = lambda m: OracleEngine(m, var_threshold=0.1)
    ga, _ = syn.kalman_arrays(8, 4, levels=2, seed=3, period=2)
    lifted = lifting.ArrayVI(ga, 2, 3, lifted=True, engine_factory=factory)
    ground = lifting.ArrayVI(ga, 2, 3, lifted=False, engine_factory=factory)
    assert lifted.quotient.compression > 1.5 and ground.quotient.compression == 1.0
    ground.tie_to(lifted)
    lifted.init_param(0)
    np.testing.assert_allclose(lifted.free_energy(), ground.free_energy(), rtol=1e-10)
    lifted.run(25, 0.05)
    ground.run(25, 0.05)
    pl, wl = lifted.ground_params()
    pg, wg = ground.ground_params()
    assert set(pl) == set(pg) == set(np.flatnonzero(np.isnan(ga.var_value)).tolist())
    for v in pg:
        np.testing.assert_allclose(pl[v], pg[v], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(wl, wg, rtol=1e-9)


# ---- host-side C++ passes (liblhvi_lift.so) against the numpy statement of the same passes -------

def _same_partition(a, b):
    """Two labelings describe the same partition: the label pairs are a bijection."""
    pairs = np.unique(np.stack([np.asarray(a), np.asarray(b)]), axis=1)
    return len(np.unique(pairs[0])) == pairs.shape[1] == len(np.unique(pairs[1]))


def test_native_lifting_library_is_built_and_exports_its_header():
    import ctypes
    import os
    import re
    from lhvi_b200 import _lift_native, build
    build.build_lift()
    lib = _lift_native.load()
    assert lib is not None and lib.lhvi_lift_abi_version() == _lift_native.ABI_VERSION
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "lhvi_lift.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(lhvi_lift_[a-z0-9_]+)\s*\(", text)))
    assert declared == sorted(_lift_native.SYMBOLS)
    raw = ctypes.CDLL(_lift_native.LIB_PATH)
    assert all(hasattr(raw, name) for name in declared)
    # argument validation
    blk = (_lift_native.LiftBlock * 1)()
    col = np.zeros(3, dtype=np.int64)
    assert lib.lhvi_lift_colour_passing(3, None, blk, 0, 10, None) == -1
    args = np.array([[0, 5]], dtype=np.int64)
    fc = np.zeros(1, dtype=np.int64)
    blk[0].args, blk[0].n, blk[0].arity, blk[0].colour = args.ctypes.data, 1, 2, fc.ctypes.data
    assert lib.lhvi_lift_colour_passing(3, col.ctypes.data, blk, 1, 10, None) == -3
    blk[0].arity = 0
    assert lib.lhvi_lift_colour_passing(3, col.ctypes.data, blk, 1, 10, None) == -2
    ids, n = _lift_native.rank64(lib, np.array([7, 3, 7, 2 ** 63, 3], dtype=np.uint64))
    assert n == 3 and ids.tolist() == [0, 1, 0, 2, 1]
    assert lib.lhvi_lift_colour_passing(0, None, None, 0, 10, None) == 0


@pytest.mark.parametrize("model", ["kalman", "relational", "relational-lumped", "grid"])
def test_native_colour_passing_equals_numpy(model):
    if model == "kalman":
        ga, split = syn.kalman_arrays(40, 12, levels=2, seed=1, period=4)[0], True
    elif model == "grid":
        g, _ = syn.gaussian_grid_graph(7)
        ga, split = lifting.arrays_from_graph(g)[0], True
    else:
        ga, split = syn.relational_hybrid_arrays(300, 4, seed=3), model == "relational"
    v_np, f_np, s_np = lifting.colour_passing(ga, split_cont_evidence=split, use_native=False)
    v_cc, f_cc, s_cc = lifting.colour_passing(ga, split_cont_evidence=split, use_native=True)
    assert s_np == s_cc
    assert _same_partition(v_np, v_cc)
    assert _same_partition(np.concatenate(f_np), np.concatenate(f_cc))
    assert v_cc.min() == 0 and v_cc.max() + 1 == len(np.unique(v_cc))
    # refinement of a coarser start (C2F's cp_run): same result from both
    start = lifting.initial_colouring(ga, split_cont_evidence=False)
    r_np = lifting.colour_passing(ga, start=start, use_native=False)[0]
    r_cc = lifting.colour_passing(ga, start=start, use_native=True)[0]
    assert _same_partition(r_np, r_cc)
    # and the lowered models agree in every sum (labels, hence record order, differ)
    m_np = lifting.lower_partition(ga, v_np, f_np, 2, 3)
    m_cc = lifting.lower_partition(ga, v_cc, f_cc, 2, 3)
    assert m_np.n_records == m_cc.n_records and m_np.n_param == m_cc.n_param
    w = np.array([0.4, 0.6])

    def tied(m, vcol):
        reps = lifting.class_stats(ga, vcol)["rep"][m.slot_class]
        eta = np.zeros(m.n_param)
        for r, off, kind, dim in zip(reps, m.var_off, m.var_kind, m.var_dim):
            rg = np.random.default_rng(int(r))
            if kind == 0:
                eta[off:off + 4:2] = rg.uniform(-1.5, 1.5, 2)
                eta[off + 1:off + 4:2] = rg.uniform(0.5, 2.0, 2)
            else:
                p = rg.uniform(0.1, 1.0, (2, dim))
                eta[off:off + 2 * dim] = (p / p.sum(axis=1, keepdims=True)).reshape(-1)
        return eta
    e_np = grad_pass(m_np, tied(m_np, v_np), w)
    e_cc = grad_pass(m_cc, tied(m_cc, v_cc), w)
    np.testing.assert_allclose(e_np[2], e_cc[2], rtol=1e-12)
    np.testing.assert_allclose(e_np[1], e_cc[1], rtol=1e-12)


@pytest.mark.parametrize("k,its,seed", [(2, 10, 0), (3, 4, 1), (2, 1, 2), (5, 10, 3)])
def test_native_evidence_split_equals_numpy_bit_for_bit(k, its, seed):
    """``lhvi_lift_split_evidence`` against ``_split_evidence_numpy``: same classes with the same
    ids, same flags, same centroids to the last bit -- classes of 1 to ~3000 members (numpy's
    pairwise sums recurse above 128), repeated values, several thresholds."""
    from oracle_engine import OracleEngine
    rng = np.random.default_rng(seed)
    ga = syn.relational_hybrid_arrays(1500, 4, observed_frac=0.8, seed=seed)
    # make the observed values lumpy: many repeats, a few wide clusters
    ev = np.flatnonzero(~np.isnan(ga.var_value))
    ga.var_value[ev] = np.round(rng.normal(0, 3, ev.size) + rng.choice([-8.0, 0.0, 5.0], ev.size), rng.integers(0, 3))
    runs = []
    for native in (True, False):
        vi = lifting.C2FArrayVI(ga, 2, 3, engine_factory=lambda m: OracleEngine(m), use_native=native)
        vi.k_mean_k, vi.k_mean_its = k, its
        vi.degrees = ga.degrees()
        vi.vcol = lifting.initial_colouring(ga, split_cont_evidence=False)
        n0 = int(vi.vcol.max()) + 1
        vi.may_split = np.zeros(n0, dtype=bool)
        vi.may_split[vi.vcol[~vi.hidden & vi.cont_dom]] = True
        vi.ev_has, vi.ev_val = np.zeros(n0, dtype=bool), np.zeros(n0)
        trace = []
        for epsilon in (4.0, 2.5, 1.0, 0.3, 0.0):
            vi._split_evidence(epsilon)
            trace.append((vi.vcol.copy(), vi.may_split.copy(), vi.ev_has.copy(), vi.ev_val.copy()))
        runs.append(trace)
    assert len(np.unique(runs[0][-1][0])) >= 20
    for a, b in zip(*runs):
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and x.shape == y.shape
            np.testing.assert_array_equal(x, y)


def test_c2f_native_and_numpy_passes_give_the_same_run():
    from oracle_engine import OracleEngine
    ga = syn.relational_hybrid_arrays(120, 3, seed=5)
    out = []
    for native in (True, False):
        vi = lifting.C2FArrayVI(ga, 2, 3, engine_factory=lambda m: OracleEngine(m, var_threshold=0.1), use_native=native)
        vi.run(40, 0.05)
        out.append(vi)
    a, b = out
    assert [n for n, _ in a.history] == [n for n, _ in b.history]
    assert _same_partition(a.vcol, b.vcol)
    np.testing.assert_allclose(a.free_energy(), b.free_energy(), rtol=1e-10)
    pa, wa = a.ground_params()
    pb, wb = b.ground_params()
    for v in pa:
        np.testing.assert_allclose(pa[v], pb[v], rtol=1e-8, atol=1e-10)


def test_prepared_graph_is_reused_and_survives_copies():
    """The native incidence lists are built once per ``GroundArrays`` and reused by later calls;
    a deep copy of the arrays drops them and rebuilds on first use; replacing a block's
    argument array invalidates them."""
    import copy
    ga = syn.relational_hybrid_arrays(200, 3, seed=4)
    v1 = lifting.colour_passing(ga)[0]
    g1 = ga._lift_graph
    v2 = lifting.colour_passing(ga, start=lifting.initial_colouring(ga, split_cont_evidence=False))[0]
    assert ga._lift_graph is g1
    clone = copy.deepcopy(ga)
    assert clone._lift_graph is None
    np.testing.assert_array_equal(lifting.colour_passing(clone)[0], v1)
    assert clone._lift_graph is not None and clone._lift_graph is not g1
    ga.blocks[0].args = ga.blocks[0].args.copy()
    np.testing.assert_array_equal(lifting.colour_passing(ga)[0], v1)
    assert ga._lift_graph is not g1
    assert _same_partition(v2, lifting.colour_passing(ga, split_cont_evidence=False)[0])


def test_native_passes_do_not_depend_on_the_thread_count(tmp_path):
    """Same class ids (not just the same partition) with 1, 3 and all OpenMP threads, on a model
    large enough for the parallel paths (> 65 536 items)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {root!r})\n"
        "import lhvi_b200\n"
        "ga = lhvi_b200.synthetic.relational_hybrid_arrays(30000, 4, seed=2)\n"
        "v, f, s = lhvi_b200.lifting.colour_passing(ga, use_native=True)\n"
        "r = lhvi_b200.lifting.colour_passing(ga, start=lhvi_b200.lifting.initial_colouring(ga, False), use_native=True)[0]\n"
        "np.savez(sys.argv[1], v=v, f=np.concatenate(f), r=r, s=s)\n")
    outs = []
    for threads in ("1", "3", None):
        env = dict(os.environ)
        env.pop("OMP_NUM_THREADS", None)
        if threads:
            env["OMP_NUM_THREADS"] = threads
        path = str(tmp_path / f"cp_{threads}.npz")
        subprocess.run([sys.executable, "-c", script, path], check=True, env=env, timeout=300)
        outs.append(dict(np.load(path)))
    assert outs[0]["v"].max() > 1000
    for other in outs[1:]:
        for k in ("v", "f", "r", "s"):
            np.testing.assert_array_equal(outs[0][k], other[k])


def test_map_values_of_the_array_engines_match_the_object_route(ns):
    """``ArrayVI.map_values`` / ``C2FArrayVI.map_values`` against ``LiftedVarInference.map`` on the
    hybrid paper-popularity golden graph (hidden reals and hidden booleans)."""
    import contextlib
    import io
    from oracle_engine import OracleEngine, use_oracle_engine
    factory = lambda m: OracleEngine(m, var_threshold=0.1)
    g, rvs = specs.CASES["hmln_hidden"][0](ns)
    ga, order = lifting.arrays_from_graph(g)
    arr = lifting.ArrayVI(ga, 2, 3, lifted=True, engine_factory=factory)
    arr.run(30, 0.1)
    got = arr.map_values()
    assert got.shape == (ga.n_vars,) and not np.isnan(got).any()
    ev = ~np.isnan(ga.var_value)
    np.testing.assert_array_equal(got[ev], ga.var_value[ev])
    # the same state pushed through the object-level class: same MAP values
    vi = use_oracle_engine(lhvi_b200.LiftedVarInference.VarInference(g, 2, 3))
    params, w = arr.ground_params()
    index = {id(rv): i for i, rv in enumerate(order)}

    def init_param():
        vi.w_tau = np.log(w)
        vi.w = w.copy()
        vi.eta, vi.eta_tau = {}, {}
        for h in vi.g.rvs:
            if h.value is not None:
                continue
            table = params[min(index[id(rv)] for rv in h.rvs)]
            vi.eta[h] = table.copy()
            if not h.domain.continuous:
                vi.eta_tau[h] = np.log(table)
    vi.init_param = init_param
    with contextlib.redirect_stdout(io.StringIO()):
        vi.run(0, lr=0.1)
    for rv in order:
        if rv.value is None:
            want = vi.map(rv)
            assert abs(got[index[id(rv)]] - want) < 1e-6, (rv, got[index[id(rv)]], want)
    hidden_disc = [index[id(rv)] for rv in order if rv.value is None and not rv.domain.continuous]
    assert hidden_disc and set(got[hidden_disc]) <= {0.0, 1.0}
    c2f = lifting.C2FArrayVI(ga, 2, 3, engine_factory=factory).run(20, 0.1)
    m = c2f.map_values()
    assert m.shape == (ga.n_vars,) and not np.isnan(m).any() and set(m[hidden_disc]) <= {0.0, 1.0}


def test_native_and_numpy_colour_passing_agree_on_random_graphs():
    """Random small factor graphs: several blocks of arity 1-4, symmetric and ordered potentials,
    repeated rows, blocks sharing a potential object, mixed evidence -- same partitions of variables
    and factors and the same number of sweeps from both implementations."""
    class Stub:                     # only .symmetric and identity matter to colour passing
        def __init__(self, symmetric):
            self.symmetric = symmetric
    doms = [lhvi_b200.Graph.Domain((-5, 5), continuous=True), lhvi_b200.Graph.Domain((0, 1))]
    for seed in range(80):
        rng = np.random.default_rng(seed)
        nv = int(rng.integers(3, 60))
        var_dom = rng.integers(0, 2, nv).astype(np.int32)
        val = np.full(nv, np.nan)
        ev = rng.random(nv) < 0.4
        val[ev] = np.where(var_dom[ev] == 0, rng.integers(0, 3, ev.sum()) * 0.5, rng.integers(0, 2, ev.sum()))
        blocks = []
        for _ in range(int(rng.integers(1, 5))):
            arity, n = int(rng.integers(1, 5)), int(rng.integers(1, 40))
            base = rng.integers(0, nv, (max(1, n // 3), arity))
            blocks.append(lifting.FactorBlock(Stub(bool(rng.integers(0, 2))), base[rng.integers(0, base.shape[0], n)].astype(np.int64)))
        if rng.random() < 0.3 and len(blocks) > 1:
            blocks[1] = lifting.FactorBlock(blocks[0].potential, rng.integers(0, nv, (5, blocks[0].arity)).astype(np.int64))
        ga = lifting.GroundArrays(doms, var_dom, val, blocks)
        for split in (True, False):
            a = lifting.colour_passing(ga, split_cont_evidence=split, use_native=False)
            b = lifting.colour_passing(ga, split_cont_evidence=split, use_native=True)
            assert a[2] == b[2], (seed, split)
            assert _same_partition(a[0], b[0]), (seed, split)
            assert _same_partition(np.concatenate(a[1]), np.concatenate(b[1])), (seed, split)
