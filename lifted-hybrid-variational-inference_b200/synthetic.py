"""Array-native synthetic models at benchmark scale (SURVEY section 8 d, configs 3-5).

Building 10 M ``F`` objects in Python takes minutes and gigabytes, so the large configurations
are emitted directly as ``LoweredModel`` column arrays.  Each generator has an object-graph twin
(``*_graph``) that builds the *same* model with ``Graph`` / ``RV`` / ``F`` at small sizes; the
tests lower the twin with ``lowering.lower_ground`` and check both routes agree, which ties the
array-native path to the plugin API.

Models

* ``relational_hybrid``: the paper-popularity hybrid MLN of the reference
  (``Demo/Data/HMLN/GeneratorPaperPopularity.py``) scaled up: P entities x G groups,
  link factors ``exp(w * In(p,t) * -(Pop(p) - Pop(t))^2)`` with the relation atoms observed,
  unary priors ``exp(w0 * -(Pop(p) - 1)^2)``, session factors between groups, a fraction of the
  entity popularities observed with unique values.
* ``gaussian_grid``: pairwise Gaussian MRF on an n x n grid with unary ``X2`` terms (config 4).
* ``kalman_arrays``: the relational Kalman filter of ``KalmanFilter.grounded_graph`` with a sparse
  transition matrix and quantised observations, as ``lifting.GroundArrays`` (config 2).
"""
from __future__ import annotations

import numpy as np

from . import lowering
from .Graph import F, RV, Domain, Graph
from .lowering import EC, ED, HC, LoweredModel, PotentialTable, RecordGroup, slot_size
from .MLNPotential import MLNPotential, eq_op
from .Potential import GaussianPotential, LinearGaussianPotential, X2Potential, XYPotential

_LINK_W, _PRIOR_W, _SESSION_W = 1.0, 0.3, 0.5


def _potentials():
    prior = MLNPotential(lambda x: eq_op(x[0], 1), w=_PRIOR_W)
    link = MLNPotential(lambda x: x[0] * eq_op(x[1], x[2]), w=_LINK_W)
    sess = MLNPotential(lambda x: x[0] * eq_op(x[1], x[2]), w=_SESSION_W)
    return prior, link, sess


def _relational_draws(P, G, observed_frac, seed):
    rng = np.random.default_rng(seed)
    observed = rng.random(P) < observed_frac
    value = rng.uniform(0.0, 10.0, size=P)
    member = rng.integers(0, 2, size=(P, G)).astype(np.int8)       # In(p,t), all observed
    sess = rng.integers(0, 2, size=(G, G)).astype(np.int8)
    return observed, value, member, sess


def _group(nd, nc, ng, ne, node, pot, poff, ecval, nscale=None, weighted=False, wf=None, gam=None,
           pure=False):
    n = len(pot)
    nh = nd + nc
    f64 = np.float64
    return RecordGroup(
        nd, nc, ng, ne, (), node,
        np.ascontiguousarray(pot, dtype=np.int32),
        np.ascontiguousarray(np.asarray(poff, dtype=np.int32).reshape(nh, n)),
        np.zeros((ng, n), f64), np.zeros((ng, n), f64),
        np.ascontiguousarray(np.asarray(ecval, dtype=f64).reshape(ne, n)),
        np.ones(n, f64) if wf is None else wf,
        np.ones((nh, n), f64) if gam is None else gam,
        np.zeros(n, f64) if nscale is None else np.asarray(nscale, dtype=f64),
        weighted or node, pure)


def relational_hybrid(P, G, K, T, *, observed_frac=0.7, seed=0, order="hub", weighted=False):
    """``LoweredModel`` of the relational hybrid model with ``P*G`` link factors.

    ``order='hub'`` sorts link records group-major (all records of one group variable are
    contiguous, entity offsets ascend inside), ``'entity'`` keeps entity-major order.
    ``weighted=True`` ships explicit unit lifted weights (the record format of a fully refined
    C2F graph)."""
    observed, value, member, sess = _relational_draws(P, G, observed_frac, seed)
    prior, link, sesspot = _potentials()
    table = PotentialTable()
    slot = slot_size(2 * K)

    hidden_p = np.flatnonzero(~observed)
    n_hidden = G + hidden_p.size
    var_off = (np.arange(n_hidden, dtype=np.int64) * slot).astype(np.int32)
    off_topic = var_off[:G]
    off_paper = np.full(P, -1, dtype=np.int64)
    off_paper[hidden_p] = var_off[G:]

    sess_pairs = [(a, b) for a in range(G) for b in range(G) if a != b]
    deg_topic = np.full(G, P, dtype=np.int64)
    for a, b in sess_pairs:
        deg_topic[a] += 1
        deg_topic[b] += 1
    deg_paper = G + 1

    groups = []
    # node-entropy records: F = log b_v scaled by (N_v - 1) minus the variable's unary factors
    # (lowering.py "unary split"): every observed-entity link is unary in its group variable,
    # every hidden entity has one unary prior
    n_obs = int(observed.sum())
    scale = np.concatenate([deg_topic - 1 - n_obs, np.full(hidden_p.size, deg_paper - 1 - 1)]).astype(np.float64)
    groups.append(_group(0, 1, 0, 0, True, np.zeros(n_hidden, np.int32), var_off[None, :], np.zeros((0, n_hidden)),
                         nscale=scale, wf=scale.copy()))

    # priors: hidden entity -> one continuous argument; observed entity -> constant record
    blk_prior_h = table.block(prior, (HC,), (None,))
    blk_prior_e = table.block(prior, (EC,), (0.0,))
    groups.append(_group(0, 1, 0, 0, False, np.full(hidden_p.size, blk_prior_h), off_paper[hidden_p][None, :],
                         np.zeros((0, hidden_p.size)), weighted=weighted, pure=True))
    obs_p = np.flatnonzero(observed)
    groups.append(_group(0, 0, 0, 1, False, np.full(obs_p.size, blk_prior_e), np.zeros((0, obs_p.size)),
                         value[obs_p][None, :], weighted=weighted, pure=True))

    # link factors In(p,t) * -(Pop(p) - Pop(t))^2
    if order == "hub":
        pp, tt = np.meshgrid(np.arange(P), np.arange(G), indexing="ij")
        pp, tt = pp.T.reshape(-1), tt.T.reshape(-1)          # group-major
    else:
        pp, tt = np.meshgrid(np.arange(P), np.arange(G), indexing="ij")
        pp, tt = pp.reshape(-1), tt.reshape(-1)
    mem = member[pp, tt]
    is_obs = observed[pp]
    # observed entity: canonical args [Pop(t) | value(p)]
    blk_e = np.array([table.block(link, (ED, EC, HC), (v, 0.0, None)) for v in (0, 1)])
    sel = np.flatnonzero(is_obs)
    groups.append(_group(0, 1, 0, 1, False, blk_e[mem[sel]], off_topic[tt[sel]][None, :],
                         value[pp[sel]][None, :], weighted=weighted, pure=True))
    # hidden entity: canonical args [Pop(p), Pop(t)]
    blk_h = np.array([table.block(link, (ED, HC, HC), (v, None, None)) for v in (0, 1)])
    sel = np.flatnonzero(~is_obs)
    pot_h = blk_h[mem[sel]]
    poff_h = np.stack([off_paper[pp[sel]], off_topic[tt[sel]]])
    # session factors between groups share the signature (two hidden continuous arguments)
    blk_s = np.array([table.block(sesspot, (ED, HC, HC), (v, None, None)) for v in (0, 1)])
    if sess_pairs:
        sa = np.array([a for a, _ in sess_pairs])
        sb = np.array([b for _, b in sess_pairs])
        pot_h = np.concatenate([pot_h, blk_s[sess[sa, sb]]])
        poff_h = np.concatenate([poff_h, np.stack([off_topic[sa], off_topic[sb]])], axis=1)
    groups.append(_group(0, 2, 0, 0, False, pot_h, poff_h, np.zeros((0, pot_h.size)), weighted=weighted))

    groups = [g for g in groups if g.n > 0]
    return LoweredModel(K, T, int(n_hidden * slot), np.zeros(n_hidden, np.uint8),
                        np.full(n_hidden, 2, np.int32), var_off, table.array(), groups)


def relational_hybrid_graph(P, G, *, observed_frac=0.7, seed=0):
    """Object-graph twin of ``relational_hybrid`` (small sizes only).  Returns
    ``(graph, topics, entities)``; hidden variables are created topics first, then entities,
    matching the slot order of the array-native model."""
    observed, value, member, sess = _relational_draws(P, G, observed_frac, seed)
    prior, link, sesspot = _potentials()
    d_bool = Domain((0, 1))
    d_real = Domain((-15, 15), continuous=True)
    topics = [RV(d_real) for _ in range(G)]
    entities = [RV(d_real, float(value[p]) if observed[p] else None) for p in range(P)]
    rvs = topics + entities
    fs = [F(prior, [e]) for e in entities]
    for p in range(P):
        for t in range(G):
            rel = RV(d_bool, int(member[p, t]))
            rvs.append(rel)
            fs.append(F(link, [rel, entities[p], topics[t]]))
    for a in range(G):
        for b in range(G):
            if a != b:
                rel = RV(d_bool, int(sess[a, b]))
                rvs.append(rel)
                fs.append(F(sesspot, [rel, topics[a], topics[b]]))
    g = Graph()
    g.rvs = set(rvs)
    g.factors = set(fs)
    g.init_nb()
    return g, topics, entities


def relational_hybrid_arrays(P, G, *, observed_frac=0.7, seed=0):
    """``relational_hybrid`` / ``relational_hybrid_graph`` as ``lifting.GroundArrays`` -- the ground
    graph itself (relation atoms included as observed discrete variables), the input of the
    array-native colour passing and of ``lifting.C2FArrayVI`` (config 5 as the reference runs it).
    Variable order: topics, entities, In(p, t) atoms (entity-major), session atoms."""
    from .lifting import FactorBlock, GroundArrays
    observed, value, member, sess = _relational_draws(P, G, observed_frac, seed)
    prior, link, sesspot = _potentials()
    d_bool = Domain((0, 1))
    d_real = Domain((-15, 15), continuous=True)
    topics = np.arange(G, dtype=np.int64)
    ents = G + np.arange(P, dtype=np.int64)
    rel = G + P + np.arange(P * G, dtype=np.int64).reshape(P, G)
    pairs = [(a, b) for a in range(G) for b in range(G) if a != b]
    srel = G + P + P * G + np.arange(len(pairs), dtype=np.int64)
    n_vars = G + P + P * G + len(pairs)
    var_dom = np.zeros(n_vars, dtype=np.int32)              # 0: real, 1: bool
    var_dom[G + P:] = 1
    var_value = np.full(n_vars, np.nan)
    var_value[ents[observed]] = value[observed]
    var_value[rel.reshape(-1)] = member.reshape(-1)
    if pairs:
        var_value[srel] = [sess[a, b] for a, b in pairs]
    pp, tt = np.meshgrid(np.arange(P), np.arange(G), indexing="ij")
    blocks = [
        FactorBlock(prior, ents.reshape(-1, 1)),
        FactorBlock(link, np.stack([rel.reshape(-1), ents[pp.reshape(-1)], topics[tt.reshape(-1)]], axis=1)),
    ]
    if pairs:
        blocks.append(FactorBlock(sesspot, np.stack([srel, topics[[a for a, _ in pairs]], topics[[b for _, b in pairs]]], axis=1)))
    return GroundArrays([d_real, d_bool], var_dom, var_value, blocks)


# ---- pairwise Gaussian grid (config 4) -------------------------------------------------------

_GRID_SIG = [[1.5, 0.6], [0.6, 1.5]]


def gaussian_grid(n, K, T, *, unary_coeff=1.0, unary_sig=2.0):
    """n x n grid: ``X2Potential`` on every node, attractive ``GaussianPotential`` on every
    horizontal / vertical edge.  Variables are row-major; edge records are sorted by their
    first endpoint."""
    table = PotentialTable()
    slot = slot_size(2 * K)
    V = n * n
    var_off = (np.arange(V, dtype=np.int64) * slot).astype(np.int32)
    idx = np.arange(V).reshape(n, n)
    right = np.stack([idx[:, :-1].reshape(-1), idx[:, 1:].reshape(-1)])
    down = np.stack([idx[:-1, :].reshape(-1), idx[1:, :].reshape(-1)])
    edges = np.concatenate([right, down], axis=1)
    edges = edges[:, np.argsort(edges[0], kind="stable")]
    deg = np.bincount(edges.reshape(-1), minlength=V) + 1
    blk_u = table.block(X2Potential(unary_coeff, unary_sig), (HC,), (None,))
    blk_e = table.block(GaussianPotential([0.0, 0.0], _GRID_SIG), (HC, HC), (None, None))
    E = edges.shape[1]
    scale = (deg - 1 - 1).astype(np.float64)       # minus the unary X2 factor (unary split)
    groups = [
        _group(0, 1, 0, 0, True, np.zeros(V, np.int32), var_off[None, :], np.zeros((0, V)), nscale=scale,
               wf=scale.copy()),
        _group(0, 1, 0, 0, False, np.full(V, blk_u), var_off[None, :], np.zeros((0, V)), pure=True),
        _group(0, 2, 0, 0, False, np.full(E, blk_e), var_off[edges], np.zeros((0, E))),
    ]
    return LoweredModel(K, T, int(V * slot), np.zeros(V, np.uint8), np.full(V, 2, np.int32),
                        var_off, table.array(), groups)


def gaussian_grid_graph(n, *, unary_coeff=1.0, unary_sig=2.0):
    d = Domain((-10, 10), continuous=True)
    X = [RV(d) for _ in range(n * n)]
    pu = X2Potential(unary_coeff, unary_sig)
    pe = GaussianPotential([0.0, 0.0], _GRID_SIG)
    fs = [F(pu, [x]) for x in X]
    for r in range(n):
        for c in range(n):
            if c + 1 < n:
                fs.append(F(pe, [X[r * n + c], X[r * n + c + 1]]))
            if r + 1 < n:
                fs.append(F(pe, [X[r * n + c], X[(r + 1) * n + c]]))
    g = Graph()
    g.rvs = set(X)
    g.factors = set(fs)
    g.init_nb()
    return g, X


def random_state(model, seed=0):
    """Deterministic ``(eta, tau, w_tau)`` with the reference's initial distributions
    (VarInference.py:197-208): means U(-1.5,1.5), variances 1, logits U(0,10), w_tau = 0."""
    rng = np.random.default_rng(seed)
    K = model.K
    eta = np.zeros(model.n_param)
    tau = np.zeros(model.n_param)
    cont = model.var_off[model.var_kind == 0].astype(np.int64)
    mu_idx = (cont[:, None] + 2 * np.arange(K)[None, :])
    eta[mu_idx] = rng.random(mu_idx.shape) * 3 - 1.5
    eta[mu_idx + 1] = 1.0
    for off, d in zip(model.var_off[model.var_kind == 1], model.var_dim[model.var_kind == 1]):
        logits = rng.random((K, d)) * 10
        tau[off:off + K * d] = logits.reshape(-1)
        e = np.e ** logits
        eta[off:off + K * d] = (e / e.sum(axis=1, keepdims=True)).reshape(-1)
    return eta, tau, np.zeros(K)


# ---- relational Kalman filter (config 2) -------------------------------------------------------

def _kalman_setup(n, t_steps, a, c, levels, seed, period=None):
    """Sparse transition ``A = a I + c shift`` (state x feeds x and x+1, cyclic), observations of
    every state at every step quantised to ``levels`` values (SURVEY section 8 d, config 2).
    With ``period`` (a divisor of n) the observation pattern repeats every ``period`` state
    dimensions, so the model is invariant under that cyclic shift and colour passing lifts it to
    ``period`` classes per time step; without, the observations are i.i.d. and nothing merges."""
    rng = np.random.default_rng(seed)
    A = np.zeros((n, n))
    A[np.arange(n), np.arange(n)] = a
    A[np.arange(n), (np.arange(n) + 1) % n] = c
    if period is None:
        data = rng.integers(0, levels, size=(n, t_steps)).astype(np.float64)
    else:
        if n % period:
            raise ValueError("period must divide n")
        data = np.tile(rng.integers(0, levels, size=(period, t_steps)), (n // period, 1)).astype(np.float64)
    return A, data


def kalman_arrays(n, t_steps, *, a=0.9, c=0.05, trans_var=1.0, obs_coeff=1.0, obs_var=0.5, levels=2, seed=0,
                  period=None):
    """``KalmanFilter.grounded_graph`` (reference ``KalmanFilter.py:15-104``) as arrays: state
    variables ``(t, x)`` (t = 0 observed with ``data[x, 0]``, later steps hidden), one observation
    leaf per hidden state, and the reference's decomposition of the transition into ``X2`` /
    ``XY`` potentials.  Returns ``(GroundArrays, state_index [t_steps, n])``."""
    from .lifting import FactorBlock, GroundArrays
    A, data = _kalman_setup(n, t_steps, a, c, levels, seed, period)
    dom = Domain((-30, 30), continuous=True)
    state = np.arange(t_steps * n, dtype=np.int64).reshape(t_steps, n)
    n_state = t_steps * n
    obs = n_state + np.arange((t_steps - 1) * n, dtype=np.int64).reshape(t_steps - 1, n)
    n_vars = n_state + (t_steps - 1) * n
    var_value = np.full(n_vars, np.nan)
    var_value[state[0]] = data[:, 0]
    var_value[obs.reshape(-1)] = data[:, 1:].T.reshape(-1)
    xy = A
    xx = xy @ xy.T                                   # sum_y outer(xy[:, y], xy[:, y])
    blocks = []
    # observation factors F(LinearGaussian(obs_coeff, obs_var), [state(t, x), observe(t, x)]), t >= 1
    blocks.append(FactorBlock(LinearGaussianPotential(obs_coeff, obs_var),
                              np.stack([state[1:].reshape(-1), obs.reshape(-1)], axis=1)))
    # node factors X2(xx[x, x]) for 0 < t < T - 1
    for val in np.unique(np.diag(xx)):
        if val != 0 and t_steps > 2:
            xs = np.flatnonzero(np.diag(xx) == val)
            blocks.append(FactorBlock(X2Potential(float(val), trans_var), state[1:t_steps - 1][:, xs].reshape(-1, 1)))
    # xy factors XY(-2 xy[x, y]) on (state(t, x), state(t + 1, y)), one block per distinct coefficient
    xs, ys = np.nonzero(xy)
    for val in np.unique(xy[xs, ys]):
        sel = xy[xs, ys] == val
        left = state[:-1][:, xs[sel]].reshape(-1)
        right = state[1:][:, ys[sel]].reshape(-1)
        blocks.append(FactorBlock(XYPotential(float(-2 * val), trans_var), np.stack([left, right], axis=1)))
    # xx factors XY(2 xx[x, y]) on (state(t, x), state(t, y)), x < y, 0 < t < T - 1
    if t_steps > 2:
        xs, ys = np.nonzero(np.triu(xx, 1))
        for val in np.unique(xx[xs, ys]):
            sel = xx[xs, ys] == val
            left = state[1:t_steps - 1][:, xs[sel]].reshape(-1)
            right = state[1:t_steps - 1][:, ys[sel]].reshape(-1)
            blocks.append(FactorBlock(XYPotential(float(2 * val), trans_var), np.stack([left, right], axis=1)))
    # X2(1) on every hidden state
    blocks.append(FactorBlock(X2Potential(1.0, trans_var), state[1:].reshape(-1, 1)))
    ga = GroundArrays([dom], np.zeros(n_vars, dtype=np.int32), var_value, blocks)
    return ga, state


def kalman_graph(n, t_steps, *, a=0.9, c=0.05, trans_var=1.0, obs_coeff=1.0, obs_var=0.5, levels=2, seed=0,
                 period=None):
    """Object-graph twin of ``kalman_arrays``: the loops of ``KalmanFilter.grounded_graph``
    (reference ``KalmanFilter.py:15-104``) over this repo's ``Graph`` classes.  Variables are
    created in the index order of the array model.  Returns ``(graph, rvs_in_index_order)``."""
    A, data = _kalman_setup(n, t_steps, a, c, levels, seed, period)
    dom = Domain((-30, 30), continuous=True)
    table = [[RV(dom, float(data[x, 0]) if t == 0 else None) for x in range(n)] for t in range(t_steps)]
    observe = [[RV(dom, float(data[x, t])) for x in range(n)] for t in range(1, t_steps)]
    fs = []
    for t in range(1, t_steps):
        for x in range(n):
            fs.append(F(LinearGaussianPotential(obs_coeff, obs_var), [table[t][x], observe[t - 1][x]]))
    xy = A
    xx = xy @ xy.T
    for t in range(t_steps - 1):
        for x in range(n):
            if t > 0 and xx[x, x] != 0:
                fs.append(F(X2Potential(float(xx[x, x]), trans_var), [table[t][x]]))
            for y in range(n):
                if xy[x, y] != 0:
                    fs.append(F(XYPotential(float(-2 * xy[x, y]), trans_var), [table[t][x], table[t + 1][y]]))
                if t > 0 and x < y and xx[x, y] != 0:
                    fs.append(F(XYPotential(float(2 * xx[x, y]), trans_var), [table[t][x], table[t][y]]))
    for t in range(1, t_steps):
        for x in range(n):
            fs.append(F(X2Potential(1.0, trans_var), [table[t][x]]))
    rvs = [rv for row in table for rv in row] + [rv for row in observe for rv in row]
    g = Graph()
    g.rvs = set(rvs)
    g.factors = set(fs)
    g.init_nb()
    return g, rvs
