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
"""
from __future__ import annotations

import numpy as np

from . import lowering
from .Graph import F, RV, Domain, Graph
from .lowering import EC, ED, HC, LoweredModel, PotentialTable, RecordGroup, slot_size
from .MLNPotential import MLNPotential, eq_op
from .Potential import GaussianPotential, X2Potential

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
