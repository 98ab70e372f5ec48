"""Small graphs used to pin the oracle (and through it the CUDA path) to the reference.

Every builder takes a namespace ``ns`` exposing the data-model and potential classes
(``Domain, RV, F, Graph, TablePotential, GaussianPotential, LinearGaussianPotential,
X2Potential, XYPotential, MLNPotential`` and the soft-logic operators).  ``make_golden.py``
passes the *reference's* modules (imported from ``/root/reference``), the tests pass this
repo's -- the same code therefore builds the same graph on both sides.  Builders return
``(graph, rvs)`` with ``rvs`` in creation order; that order indexes every golden array.
"""
from __future__ import annotations

import numpy as np


def _graph(ns, rvs, factors):
    g = ns.Graph()
    g.rvs = set(rvs)
    g.factors = set(factors)
    g.init_nb()
    return g, rvs


def chain_table(ns):
    """A-B-C-D-E boolean chain, B observed (reference Demo/old/VarInferenceTestDemo.py)."""
    dom = ns.Domain((0, 1))
    pot = ns.TablePotential({(0, 0): 1, (0, 1): 0.1, (1, 0): 0.1, (1, 1): 1})
    A, B, C, D, E = ns.RV(dom), ns.RV(dom, 1), ns.RV(dom), ns.RV(dom), ns.RV(dom)
    fs = [ns.F(pot, [A, B]), ns.F(pot, [E, D]), ns.F(pot, [B, C]), ns.F(pot, [D, C])]
    return _graph(ns, [A, B, C, D, E], fs)


def tri_table3(ns):
    """Three-state variables with a ternary table factor plus pairwise ones."""
    dom = ns.Domain((0, 1, 2))
    rng = np.random.default_rng(5)
    t3 = rng.uniform(0.2, 2.0, size=(3, 3, 3))
    t2 = rng.uniform(0.2, 2.0, size=(3, 3))
    p3 = ns.TablePotential(t3)
    p2 = ns.TablePotential(t2)
    X = [ns.RV(dom) for _ in range(4)] + [ns.RV(dom, 2)]
    fs = [ns.F(p3, [X[0], X[1], X[2]]), ns.F(p2, [X[2], X[3]]), ns.F(p2, [X[3], X[4]]),
          ns.F(p2, [X[0], X[3]])]
    return _graph(ns, X, fs)


def gauss_net(ns):
    """All-continuous net with every exp-quadratic potential class and point evidence."""
    dom = ns.Domain((-10, 10), continuous=True)
    X = [ns.RV(dom) for _ in range(5)] + [ns.RV(dom, 1.3), ns.RV(dom, -0.4)]
    pg = ns.GaussianPotential([0.0, 0.0], [[10.0, -7.0], [-7.0, 10.0]])
    pg2 = ns.GaussianPotential([1.0, -1.0], [[2.0, 0.5], [0.5, 1.0]])
    lg = ns.LinearGaussianPotential(0.8, 0.5)
    x2 = ns.X2Potential(1.0, 2.0)
    xy = ns.XYPotential(-0.6, 2.0)
    fs = [ns.F(pg, [X[0], X[1]]), ns.F(pg2, [X[1], X[2]]), ns.F(lg, [X[2], X[5]]),
          ns.F(lg, [X[3], X[6]]), ns.F(xy, [X[2], X[3]]), ns.F(xy, [X[3], X[4]]),
          ns.F(pg, [X[4], X[0]])]
    fs += [ns.F(x2, [X[i]]) for i in range(5)]
    return _graph(ns, X, fs)


def _paper_pop(ns, n_paper, n_topic, hidden_in, seed):
    """Paper-popularity HMLN (reference Demo/Data/HMLN/GeneratorPaperPopularity.py) built
    by hand: PaperIn(p,t) boolean, PaperPopularity(p) / TopicPopularity(t) real."""
    rng = np.random.default_rng(seed)
    d_bool = ns.Domain((0, 1))
    d_real = ns.Domain((-15, 15), continuous=True)
    prior = ns.MLNPotential(lambda x: ns.eq_op(x[0], 1), w=0.3)
    link = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], x[2]), w=1)
    sess = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], x[2]), w=0.5)
    topics = [ns.RV(d_real, None if t % 3 else float(rng.uniform(0, 10))) for t in range(n_topic)]
    papers = [ns.RV(d_real, None if rng.random() < 0.5 else float(rng.uniform(0, 10)))
              for _ in range(n_paper)]
    rvs = topics + papers
    fs = [ns.F(prior, [p]) for p in papers]
    for i, p in enumerate(papers):
        for j, t in enumerate(topics):
            is_hidden = hidden_in and (i + j) % 3 == 0
            pin = ns.RV(d_bool, None if is_hidden else int(rng.integers(0, 2)))
            rvs.append(pin)
            fs.append(ns.F(link, [pin, p, t]))
    for a in range(n_topic):
        for b in range(n_topic):
            if a != b and (a + b) % 2 == 1:
                same = ns.RV(d_bool, int(rng.integers(0, 2)))
                rvs.append(same)
                fs.append(ns.F(sess, [same, topics[a], topics[b]]))
    return _graph(ns, rvs, fs)


def hmln_evidence(ns):
    """Relation atoms all observed: the reference's category-gradient bug cannot fire."""
    return _paper_pop(ns, n_paper=4, n_topic=3, hidden_in=False, seed=11)


def hmln_hidden(ns):
    """Some PaperIn atoms hidden next to hidden popularities: H2 fires in the reference."""
    return _paper_pop(ns, n_paper=3, n_topic=2, hidden_in=True, seed=12)


def robot_like(ns):
    """Arity-5 boolean clause, boolean pair clauses and hybrid bool x real factors
    (shapes of reference Demo/Data/HMLN/GeneratorRobotMapping.py)."""
    d_bool = ns.Domain((0, 1))
    d_len = ns.Domain((0, 1), continuous=True)
    clause5 = ns.MLNPotential(
        lambda x: 1 - (x[0] == 1) * (x[1] == 1) * (x[2] == 0) * (x[3] == 1) * (1 - x[4]), w=1.591)
    excl = ns.MLNPotential(lambda x: ns.or_op(ns.neg_op(x[0]), ns.neg_op(x[1])), w=3)
    unit = ns.MLNPotential(lambda x: x[0], w=-0.737)
    door = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], 0.1), w=3.228)
    wall = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], 0.341), w=3.754)
    sw = [ns.RV(d_bool) for _ in range(2)]          # SegType(s, W)
    sd = [ns.RV(d_bool) for _ in range(2)]          # SegType(s, D)
    po = [ns.RV(d_bool), ns.RV(d_bool, 1)]          # PartOf(s, l), one observed
    al = ns.RV(d_bool, 0)                           # Aligned(s2, s1) observed
    ln = [ns.RV(d_len), ns.RV(d_len, 0.27)]         # Length(s), one observed
    rvs = sw + sd + po + [al] + ln
    fs = [ns.F(clause5, [sw[0], sw[1], po[0], po[1], al]),
          ns.F(excl, [sw[0], sd[0]]), ns.F(excl, [sw[1], sd[1]]),
          ns.F(unit, [sd[0]]), ns.F(unit, [sd[1]]),
          ns.F(door, [sd[0], ln[0]]), ns.F(door, [sd[1], ln[1]]),
          ns.F(wall, [sw[0], ln[0]]), ns.F(wall, [sw[1], ln[1]])]
    return _graph(ns, rvs, fs)


def rgm_small(ns, evidence="split"):
    """Relational Gaussian model recession -> market(c) -> loss(c,b) -> revenue(b)
    (reference Demo/Data/RGM/Generator.py) with 3 categories x 2 banks.

    ``evidence='split'``: observed losses take two well-separated groups of values so the
    C2F k-means split (k=2) is unambiguous; ``'exact'``: repeated exact values so plain
    colour passing lifts them."""
    d = ns.Domain((-50, 50), continuous=True)
    p1 = ns.GaussianPotential([0.0, 0.0], [[10.0, -7.0], [-7.0, 10.0]])
    p2 = ns.GaussianPotential([0.0, 0.0], [[10.0, 5.0], [5.0, 10.0]])
    p3 = ns.GaussianPotential([0.0, 0.0], [[10.0, 7.0], [7.0, 10.0]])
    n_cat, n_bank = 3, 2
    if evidence == "split":
        obs = {(0, 0): 4.0, (1, 0): 4.5, (2, 1): -6.0, (0, 1): -6.25}
    else:
        obs = {(0, 0): 4.0, (1, 0): 4.0, (2, 0): 4.0, (0, 1): -6.0}
    recession = ns.RV(d)
    market = [ns.RV(d) for _ in range(n_cat)]
    revenue = [ns.RV(d) for _ in range(n_bank)]
    loss = {(c, b): ns.RV(d, obs.get((c, b))) for c in range(n_cat) for b in range(n_bank)}
    rvs = [recession] + market + revenue + list(loss.values())
    fs = [ns.F(p1, [recession, m]) for m in market]
    for (c, b), l in loss.items():
        fs.append(ns.F(p2, [market[c], l]))
        fs.append(ns.F(p3, [l, revenue[b]]))
    return _graph(ns, rvs, fs)


def rgm_exact(ns):
    return rgm_small(ns, evidence="exact")


def smokers(ns):
    """Boolean friends-and-smokers MLN (reference Demo/old/DiscreteVarInferenceDemo.py),
    3 people; symmetric enough that colour passing merges variables."""
    d_bool = ns.Domain((0, 1))
    f_fs = ns.MLNPotential(lambda a: ns.imp_op(a[0], ns.bic_op(a[1], a[2])), 0.1)
    f_sc = ns.MLNPotential(lambda a: ns.imp_op(a[0], a[1]), 1)
    n = 3
    smoke = [ns.RV(d_bool, 1 if i == 2 else None) for i in range(n)]
    cancer = [ns.RV(d_bool) for _ in range(n)]
    rvs = smoke + cancer
    fs = [ns.F(f_sc, [smoke[i], cancer[i]]) for i in range(n)]
    for i in range(n):
        for j in range(n):
            if i > j:
                fr = ns.RV(d_bool, 1 if (i + j) % 2 else None)
                rvs.append(fr)
                fs.append(ns.F(f_fs, [fr, smoke[i], smoke[j]]))
    return _graph(ns, rvs, fs)


def ring_xy(ns):
    """Symmetric ring of identical continuous variables: colour passing collapses it to one
    class whose factor touches that class twice (exercises SURVEY H6, first-occurrence
    index + count)."""
    d = ns.Domain((-5, 5), continuous=True)
    n = 4
    X = [ns.RV(d) for _ in range(n)]
    xy = ns.XYPotential(-0.8, 1.0)
    x2 = ns.X2Potential(2.0, 1.0)
    fs = [ns.F(xy, [X[i], X[(i + 1) % n]]) for i in range(n)] + [ns.F(x2, [x]) for x in X]
    return _graph(ns, X, fs)


def hmln_demo(ns):
    """The reference's paper-popularity demo at its own size (Demo/HMLN/DemoPaperPopularity.py with
    Demo/Data/HMLN/GeneratorPaperPopularity.py: 300 papers x 10 topics, 3400 ground atoms, 3390
    factors) and its own evidence file (Demo/Data/HMLN/0, copied to hmln_demo_evidence.json):
    70 % of the popularities and about a third of the PaperIn atoms observed, the rest -- hidden
    booleans next to hidden reals -- inferred.  BASELINE config 1's model."""
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    data = {tuple(k): v for k, v in json.load(open(os.path.join(here, "hmln_demo_evidence.json")))["evidence"]}
    n_paper, n_topic = 300, 10
    d_bool = ns.Domain((0, 1))
    d_real = ns.Domain((-15, 15), continuous=True)
    prior = ns.MLNPotential(lambda x: ns.eq_op(x[0], 1), w=0.3)
    sess = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], x[2]), w=0.5)
    link = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], x[2]), w=1)
    topics = [ns.RV(d_real, data.get(("TopicPopularity", f"t{t}"))) for t in range(n_topic)]
    papers = [ns.RV(d_real, data.get(("PaperPopularity", f"p{p}"))) for p in range(n_paper)]
    rvs = topics + papers
    fs = [ns.F(prior, [p]) for p in papers]
    for a in range(n_topic):
        for b in range(n_topic):
            if a != b:
                same = ns.RV(d_bool, data.get(("SameSession", f"t{a}", f"t{b}")))
                rvs.append(same)
                fs.append(ns.F(sess, [same, topics[a], topics[b]]))
    for i, p in enumerate(papers):
        for j, t in enumerate(topics):
            pin = ns.RV(d_bool, data.get(("PaperIn", f"p{i}", f"t{j}")))
            rvs.append(pin)
            fs.append(ns.F(link, [pin, p, t]))
    return _graph(ns, rvs, fs)


def robot_demo(ns):
    """The reference's robot-mapping demo (Demo/HMLN/DemoRobotMapping.py with
    Demo/Data/HMLN/GeneratorRobotMapping.py and the laser-scan facts Demo/Data/HMLN/robot-map, copied
    to robot_demo_evidence.json): 37 segments, 2 lines, 3 segment types -- 1628 ground atoms, 3182
    factors, among them 2664 five-argument clauses over four hidden booleans -- with the demo's
    closed-world assumption on the Aligned atoms."""
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    data = {tuple(k): v for k, v in json.load(open(os.path.join(here, "robot_demo_evidence.json")))["evidence"]}
    seg = [f"A1_{i}" for i in range(1, 38)]
    typ, line = ["W", "D", "O"], ["LA1", "LA2"]
    d_bool = ns.Domain((0, 1))
    d_len = ns.Domain((0, 1), continuous=True)
    d_dep = ns.Domain((0, 0.5), continuous=True)
    rvs = []

    def atom(domain, key, closed_world=False):
        value = data.get(key, 0 if closed_world else None)
        rv = ns.RV(domain, value)
        rvs.append(rv)
        return rv
    part = {(s_, l): atom(d_bool, ("PartOf", s_, l)) for s_ in seg for l in line}
    stype = {(s_, t): atom(d_bool, ("SegType", s_, t)) for s_ in seg for t in typ}
    aligned = {(a, b): atom(d_bool, ("Aligned", a, b), closed_world=True) for a in seg for b in seg}
    length = {s_: atom(d_len, ("Length", s_)) for s_ in seg}
    depth = {s_: atom(d_dep, ("Depth", s_)) for s_ in seg}
    p0 = ns.MLNPotential(lambda x: ns.or_op(ns.neg_op(x[0]), ns.neg_op(x[1])), w=3)
    p1 = ns.MLNPotential(lambda x: 1 - (x[0] == 0) * (x[1] == 0) * (x[2] == 0), w=3)
    p2 = ns.MLNPotential(lambda x: 1 - (x[0] == 1) * (x[1] == 1) * (x[2] == 0) * (x[3] == 1) * (1 - x[4]), w=1.591)
    p3 = ns.MLNPotential(lambda x: x[0], w=0.3)
    p4 = ns.MLNPotential(lambda x: x[0], w=-0.737)
    p5 = ns.MLNPotential(lambda x: x[0], w=-0.077)
    p6 = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], 0.1), w=3.228)
    p7 = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], 0.02), w=2.668)
    p8 = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], 0.341), w=3.754)
    p9 = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], 0.001), w=2.532)
    fs = []
    for s_ in seg:
        fs += [ns.F(p0, [stype[s_, a], stype[s_, b]]) for a in typ for b in typ if a != b]
        fs.append(ns.F(p1, [stype[s_, "W"], stype[s_, "D"], stype[s_, "O"]]))
        fs += [ns.F(p3, [stype[s_, "W"]]), ns.F(p4, [stype[s_, "D"]]), ns.F(p5, [stype[s_, "O"]])]
        fs += [ns.F(p6, [stype[s_, "D"], length[s_]]), ns.F(p7, [stype[s_, "D"], depth[s_]]),
               ns.F(p8, [stype[s_, "W"], length[s_]]), ns.F(p9, [stype[s_, "W"], depth[s_]])]
    for s1 in seg:
        for s2 in seg:
            if s1 != s2:
                fs += [ns.F(p2, [stype[s1, "W"], stype[s2, "W"], part[s1, l], part[s2, l], aligned[s2, s1]])
                       for l in line]
    return _graph(ns, rvs, fs)


def edge_mix(ns):
    """Degenerate pieces in one graph: hidden variables with no factor at all (N = 0: the node
    term has scale -1), a variable with a single unary factor (N = 1: the node term vanishes), a
    factor that takes the same variable twice (the reference's ``f.nb.index(rv)`` first-occurrence
    rule), and a factor whose arguments are all observed (a constant of the energy)."""
    dc = ns.Domain((-10, 10), continuous=True)
    db = ns.Domain((0, 1))
    lonely_c, lonely_b = ns.RV(dc), ns.RV(db)
    single = ns.RV(dc)
    twice, other = ns.RV(dc), ns.RV(dc)
    seen_a, seen_b = ns.RV(dc, 0.7), ns.RV(dc, -1.1)
    flag = ns.RV(db)
    pg = ns.GaussianPotential([0.5, -0.5], [[2.0, 0.6], [0.6, 1.5]])
    xy = ns.XYPotential(-0.4, 1.5)
    x2 = ns.X2Potential(1.0, 2.0)
    mix = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], x[2]), w=0.4)
    fs = [ns.F(x2, [single]), ns.F(pg, [twice, twice]), ns.F(xy, [twice, other]), ns.F(x2, [other]),
          ns.F(pg, [seen_a, seen_b]), ns.F(mix, [flag, other, seen_a]), ns.F(mix, [flag, twice, twice])]
    return _graph(ns, [lonely_c, lonely_b, single, twice, other, seen_a, seen_b, flag], fs)


def denoise(ns):
    """3 x 3 crop of the reference's image-denoising model (Demo/old/DenoisingDemo.py:35-70): a hidden
    pixel per observed one, ``ImageNodePotential`` between the two, ``ImageEdgePotential`` on the grid
    edges -- the demo's (0, 3.5, 25) on the horizontal ones and, so that the distance term and the
    truncation are exercised at these magnitudes, (0.05, 1.2, 1.0) on the vertical ones."""
    rng = np.random.default_rng(11)
    dom = ns.Domain((0, 255), continuous=True)
    row = col = 3
    obs = rng.uniform(-3.0, 6.0, size=(row, col))
    x = [ns.RV(dom) for _ in range(row * col)]
    y = [ns.RV(dom, float(obs[i, j])) for i in range(row) for j in range(col)]
    pxo = ns.ImageNodePotential(0, 5)
    ph = ns.ImageEdgePotential(0, 3.5, 25)
    pv = ns.ImageEdgePotential(0.05, 1.2, 1.0)
    fs = [ns.F(pxo, [x[i], y[i]]) for i in range(row * col)]
    fs += [ns.F(ph, [x[i * col + j], x[i * col + j + 1]]) for i in range(row) for j in range(col - 1)]
    fs += [ns.F(pv, [x[i * col + j], x[(i + 1) * col + j]]) for i in range(row - 1) for j in range(col)]
    return _graph(ns, x + y, fs)


def hard_mln(ns):
    """``MLNHardPotential`` over continuous arguments (MLNPotential.py:43-49) next to soft factors: a
    half-plane constraint between two hidden reals, a band constraint switched by a hidden boolean
    against an observed real, and one against a discrete observation."""
    dc = ns.Domain((-10, 10), continuous=True)
    db = ns.Domain((0, 1))
    a, b, c = ns.RV(dc), ns.RV(dc), ns.RV(dc)
    seen = ns.RV(dc, 0.4)
    flag, on = ns.RV(db), ns.RV(db, 1)
    half = ns.MLNHardPotential(lambda x: x[0] - x[1] + 0.3)
    band = ns.MLNHardPotential(lambda x: x[0] * (4.0 + ns.eq_op(x[1], x[2])) + (1 - x[0]) * 0.5)
    gate = ns.MLNHardPotential(lambda x: x[0] * (2.0 - x[1] * x[1]) + 0.1)
    soft = ns.MLNPotential(lambda x: x[0] * ns.eq_op(x[1], x[2]), w=0.6)
    pg = ns.GaussianPotential([0.2, -0.1], [[1.5, 0.4], [0.4, 1.2]])
    x2 = ns.X2Potential(1.0, 2.0)
    fs = [ns.F(half, [a, b]), ns.F(band, [flag, c, seen]), ns.F(gate, [on, b]), ns.F(soft, [flag, a, c]),
          ns.F(pg, [b, c]), ns.F(x2, [a]), ns.F(x2, [c])]
    return _graph(ns, [a, b, c, seen, flag, on], fs)


def gabp_grid(ns, row=4, col=4):
    """Loopy grid for Gaussian belief propagation (reference GaBP.py): hidden x[i][j] with an observed
    y[i][j]; ``LinearGaussianPotential`` between the two, an ``X2Potential`` prior on every x,
    zero-mean ``GaussianPotential`` on the horizontal edges and ``XYPotential`` on the vertical ones."""
    rng = np.random.default_rng(3)
    dom = ns.Domain((-10, 10), continuous=True)
    x = [ns.RV(dom) for _ in range(row * col)]
    y = [ns.RV(dom, float(v)) for v in rng.uniform(-2.0, 2.0, size=row * col)]
    obs = ns.LinearGaussianPotential(0.9, 0.6)
    prior = ns.X2Potential(1.0, 2.5)
    edge = ns.GaussianPotential([0.0, 0.0], [[2.0, 0.9], [0.9, 1.5]])
    xy = ns.XYPotential(-0.5, 2.0)
    fs = [ns.F(obs, [x[i], y[i]]) for i in range(row * col)] + [ns.F(prior, [x[i]]) for i in range(row * col)]
    fs += [ns.F(edge, [x[i * col + j], x[i * col + j + 1]]) for i in range(row) for j in range(col - 1)]
    fs += [ns.F(xy, [x[i * col + j], x[(i + 1) * col + j]]) for i in range(row - 1) for j in range(col)]
    return _graph(ns, x + y, fs)


CASES = {
    # name: (builder, K, T, engines)
    "chain_table": (chain_table, 3, 3, ("ground", "lifted", "c2f")),
    "tri_table3": (tri_table3, 2, 3, ("ground",)),
    "gauss_net": (gauss_net, 2, 3, ("ground", "lifted")),
    "gauss_net_t5": (gauss_net, 1, 5, ("ground",)),
    "hmln_evidence": (hmln_evidence, 2, 3, ("ground", "lifted", "c2f")),
    "hmln_hidden": (hmln_hidden, 2, 3, ("ground",)),
    "robot_like": (robot_like, 2, 3, ("ground",)),
    "rgm_split": (rgm_small, 2, 3, ("ground", "lifted", "c2f")),
    "rgm_exact": (rgm_exact, 1, 3, ("lifted", "c2f")),
    "smokers": (smokers, 2, 4, ("ground", "lifted")),
    "ring_xy": (ring_xy, 2, 3, ("ground", "lifted")),
    "edge_mix": (edge_mix, 2, 3, ("ground", "lifted")),
    "hmln_demo": (hmln_demo, 2, 3, ("ground", "lifted")),
    "robot_demo": (robot_demo, 2, 3, ("ground",)),
    "denoise": (denoise, 2, 3, ("ground", "lifted")),
    "hard_mln": (hard_mln, 2, 3, ("ground",)),
}


def inject_values(rv_index, is_continuous, n_states, K, seed):
    """Deterministic initial parameters for the variable (or class) whose smallest member
    has creation index ``rv_index``; same distribution as ``init_param``
    (reference VarInference.py:197-208)."""
    rng = np.random.default_rng([seed, rv_index])
    if is_continuous:
        out = np.ones((K, 2))
        out[:, 0] = rng.random(K) * 3 - 1.5
        out[:, 1] = 0.5 + rng.random(K)      # vary the variances too (init_param uses 1)
        return out
    return rng.random((K, n_states)) * 10
