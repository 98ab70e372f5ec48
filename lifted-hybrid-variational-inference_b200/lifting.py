"""Array-native colour passing: the lifted engines at sizes where a
Python object per ground variable and factor is not an option.

The reference compresses a ground ``Graph`` with ``CompressedGraphWithObs.CompressedGraph.run``
(``:264-271``): variables start coloured by (domain, hidden | exact evidence value)
(``init_cluster`` ``:187-234``), factors by their potential (through the potentials' own
``__hash__`` / ``__eq__``), and until the number of variable classes stops growing

* factors are split by the tuple of their arguments' classes -- sorted if the potential is
  ``symmetric`` (``split_factors`` ``:152-175,260-262``),
* variables are split by the multiset of their neighbouring factor classes (positions are *not*
  part of the key, ``split_rvs`` ``:47-76,249-258``).

``colour_passing`` does exactly that on index arrays: in the host-side C++ library when it is built
(``csrc/host/lhvi_lift.cpp`` through ``_lift_native``: hash-table passes on all OpenMP threads, class
ids in order of first appearance), else with numpy sort / unique passes -- the same partition either
way.  ``lower_partition`` writes the compressed model's record columns straight from the colour
arrays -- class size, representative degree ``N`` (``:45``), neighbour counts ``count[f]`` (``:43``),
mean evidence value / variance (``:9-13``) by ``bincount`` / ``unique`` passes -- and is checked column
for column against ``lowering.lower_compressed`` over the class handles of ``quotient`` (the object
route, kept as that cross-check).  ``ArrayVI`` / ``C2FArrayVI`` are ``LiftedVarInference`` /
``C2FVarInference`` over such models; ``tests/test_lifting.py`` checks every piece against the
object-level implementation on the object-graph twins of the generators and on the golden graphs.
"""
from __future__ import annotations

import time
from collections import Counter
from dataclasses import dataclass, field

import numpy as np

from . import _lift_native, lowering


@dataclass
class FactorBlock:
    """``n`` ground factors sharing one potential object: ``args[i]`` are the variable indices of
    factor i in argument order."""
    potential: object
    args: np.ndarray              # int64 [n, arity]

    @property
    def n(self):
        return int(self.args.shape[0])

    @property
    def arity(self):
        return int(self.args.shape[1])


@dataclass
class GroundArrays:
    """A ground graph as arrays.  ``var_dom[v]`` indexes ``domains``; ``var_value[v]`` is the
    evidence value (NaN: hidden; for a discrete domain the value itself, as the reference stores
    it in ``RV.value``)."""
    domains: list
    var_dom: np.ndarray           # int32 [n_vars]
    var_value: np.ndarray         # float64 [n_vars]
    blocks: list = field(default_factory=list)

    @property
    def n_vars(self):
        return int(self.var_dom.shape[0])

    @property
    def n_factors(self):
        return sum(b.n for b in self.blocks)

    def degrees(self):
        """Ground degree of every variable (``Graph.init_nb``: one entry per argument position)."""
        deg = np.zeros(self.n_vars, dtype=np.int64)
        for b in self.blocks:
            deg += np.bincount(b.args.reshape(-1), minlength=self.n_vars)
        return deg


def _rank_rows(cols):
    """Dense ids (0..m-1, in lexicographic order of the rows) of the rows of ``cols`` (list of
    equal-length integer arrays)."""
    if len(cols) == 1:
        _, inv = np.unique(cols[0], return_inverse=True)
        return inv.astype(np.int64)
    # non-negative columns whose ranges multiply to less than 2^62 pack into one sortable key
    if all(c.dtype.kind in "iu" for c in cols) and all(c.size and int(c.min()) >= 0 for c in cols):
        spans = [int(c.max()) + 1 for c in cols]
        total = 1
        for sp in spans:
            total *= sp
        if total < (1 << 62):
            key = np.zeros(cols[0].shape[0], dtype=np.int64)
            for c, sp in zip(cols, spans):
                key = key * sp + c.astype(np.int64)
            _, inv = np.unique(key, return_inverse=True)
            return inv.astype(np.int64)
    order = np.lexsort(cols[::-1])
    stacked = np.stack([c[order] for c in cols])
    new = np.ones(stacked.shape[1], dtype=bool)
    new[1:] = (stacked[:, 1:] != stacked[:, :-1]).any(axis=0)
    ids_sorted = np.cumsum(new) - 1
    out = np.empty(stacked.shape[1], dtype=np.int64)
    out[order] = ids_sorted
    return out


def _rank_hashed(vcol, H1, H2):
    """Dense ids of the rows (vcol, H1, H2): one sort of a 64-bit key that mixes the own class
    into the first hash, then an exact check that every resulting group is constant in all three
    columns (so a key collision cannot merge classes); falls back to the general ranking if not."""
    with np.errstate(over="ignore"):
        key = _mix(vcol, 0xA4093822299F31D0) + H1
    uniq, first, ids = np.unique(key, return_index=True, return_inverse=True)
    ids = ids.astype(np.int64).reshape(-1)
    if (np.array_equal(vcol[first][ids], vcol) and np.array_equal(H1[first][ids], H1)
            and np.array_equal(H2[first][ids], H2)):
        return ids
    return _rank_rows([vcol, H1.view(np.int64), H2.view(np.int64)])


def _potential_ids(blocks):
    """Colour of each block's potential: equal potentials (``==``, the reference's dict keys)
    share an id."""
    seen = {}
    ids = []
    for b in blocks:
        ids.append(seen.setdefault(b.potential, len(seen)))
    return ids


def _mix(x, seed):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def initial_colouring(ga: GroundArrays, split_cont_evidence=True):
    """``init_cluster`` (``CompressedGraphWithObs.py:187-234``): one class per (domain, hidden |
    evidence value); with ``split_cont_evidence=False`` all continuous observations of a domain
    share one class."""
    hidden = np.isnan(ga.var_value)
    val = np.where(hidden, 0.0, ga.var_value)
    if not split_cont_evidence:
        cont = np.array([bool(d.continuous) for d in ga.domains])[ga.var_dom]
        val = np.where(cont, 0.0, val)
    _, val_id = np.unique(val, return_inverse=True)
    return _rank_rows([ga.var_dom.astype(np.int64), hidden.astype(np.int64), val_id.astype(np.int64)])


def colour_passing(ga: GroundArrays, split_cont_evidence=True, max_sweeps=1000, start=None, use_native=None):
    """Coarsest equitable partition the reference's ``CompressedGraph.run`` converges to.

    Returns ``(var_colour [n_vars], [factor_colour of each block], sweeps)`` with dense class ids.
    ``split_cont_evidence=False`` starts with all continuous observations of a domain in one
    class (the coarse start of C2F, ``init_cluster(False)``).  ``start`` (a variable colouring)
    replaces the initial classes: the result is the coarsest equitable refinement of it (C2F's
    ``cp_run`` after an evidence split, ``C2FVarInference.py:33-61``).

    The multiset of neighbouring factor classes of a variable is compared through two
    independent 64-bit multiset hashes (sums of mixed class ids, order-free like the
    reference's sorted tuple); a false merge needs a 128-bit collision."""
    nv = ga.n_vars
    pot_id = _potential_ids(ga.blocks)
    native = _lift_native.load() if use_native is None else (_lift_native.load() if use_native else None)
    native_ok = native is not None and all(1 <= b.arity <= 16 for b in ga.blocks)
    if start is None:
        vcol = initial_colouring(ga, split_cont_evidence)
    elif native_ok:
        vcol = np.asarray(start, dtype=np.int64)         # any labels >= 0: the library makes them dense
    else:
        vcol = _rank_rows([np.asarray(start, dtype=np.int64)])

    # blocks whose potentials compare equal are one colour to start with: rank them together
    if native_ok:
        # the same passes in C++ with hash tables (include/lhvi_lift.h); ids in order of first appearance
        first_colour = {}
        blocks = [(b.args, getattr(b.potential, "symmetric", False),
                   first_colour.setdefault((pid, b.arity), len(first_colour))) for b, pid in zip(ga.blocks, pot_id)]
        return _lift_native.colour_passing(native, vcol, blocks, max_sweeps, holder=ga)
    merged = {}
    for i, (b, pid) in enumerate(zip(ga.blocks, pot_id)):
        merged.setdefault((pid, b.arity), []).append(i)
    supers = []                                  # (member block indices, args, symmetric)
    for (pid, arity), members in merged.items():
        args = np.concatenate([ga.blocks[i].args for i in members]) if len(members) > 1 else ga.blocks[members[0]].args
        supers.append((members, args.astype(np.int64), bool(getattr(ga.blocks[members[0]].potential, "symmetric", False))))
    scols = [np.full(a.shape[0], j, dtype=np.int64) for j, (_, a, _) in enumerate(supers)]

    # incidence entries (one per argument position), sorted by variable once
    inc_var = np.concatenate([a.reshape(-1) for _, a, _ in supers]) if supers else np.zeros(0, np.int64)
    sizes = [a.shape[0] for _, a, _ in supers]
    foff = np.cumsum([0] + sizes)
    inc_fac = np.concatenate([np.repeat(np.arange(a.shape[0]) + foff[j], a.shape[1])
                              for j, (_, a, _) in enumerate(supers)]) if supers else np.zeros(0, np.int64)
    by_var = np.argsort(inc_var, kind="stable")
    inc_var_s, inc_fac_s = inc_var[by_var], inc_fac[by_var]
    seg_first = np.ones(inc_var_s.size, dtype=bool)
    seg_first[1:] = inc_var_s[1:] != inc_var_s[:-1]
    seg_start = np.flatnonzero(seg_first)
    seg_var = inc_var_s[seg_start]

    n_classes = -1
    sweeps = 0
    while n_classes != int(vcol.max()) + 1 and sweeps < max_sweeps:
        n_classes = int(vcol.max()) + 1
        sweeps += 1
        # ---- split factors by (own class, classes of the arguments)
        offset = 0
        for j, (_, args, sym) in enumerate(supers):
            arg_cols = vcol[args]
            if sym:
                arg_cols = np.sort(arg_cols, axis=1)
            ids = _rank_rows([scols[j]] + [arg_cols[:, a] for a in range(args.shape[1])])
            scols[j] = ids + offset
            offset += int(ids.max()) + 1 if args.shape[0] else 0
        # ---- split variables by (own class, multiset of neighbouring factor classes)
        if inc_var_s.size:
            fc = np.concatenate(scols)[inc_fac_s]
            with np.errstate(over="ignore"):
                h1 = np.add.reduceat(_mix(fc, 0x243F6A8885A308D3), seg_start)
                h2 = np.add.reduceat(_mix(fc, 0x13198A2E03707344), seg_start)
            H1 = np.zeros(nv, dtype=np.uint64)
            H2 = np.zeros(nv, dtype=np.uint64)
            H1[seg_var], H2[seg_var] = h1, h2
            vcol = _rank_hashed(vcol, H1, H2)
    fcols = [None] * len(ga.blocks)
    for j, (members, _, _) in enumerate(supers):
        pos = 0
        for i in members:
            fcols[i] = scols[j][pos:pos + ga.blocks[i].n]
            pos += ga.blocks[i].n
    return vcol, fcols, sweeps


def _domain_value(domain, x):
    """The element of a discrete domain's ``values`` that equals ``x`` (evidence travels as
    float64 in ``GroundArrays.var_value``; table potentials index with the original value)."""
    if not domain.continuous:
        for v in domain.values:
            if v == x:
                return v
    return float(x)


class _ClassRV:
    """Variable class handle for ``lowering.lower_graph`` (what it reads of a ``SuperRV``)."""

    def __init__(self, uid, domain, value, variance, size, degree, rep):
        self.uid, self.domain, self.value, self.variance = uid, domain, value, variance
        self.rvs = range(size)
        self.N = degree
        self.rep = rep                 # ground index of the representative (smallest member)
        self.count = Counter()
        self.nb = ()

    def __lt__(self, other):
        return self.uid < other.uid


class _ClassF:
    def __init__(self, uid, potential, size, rep):
        self.uid, self.potential = uid, potential
        self.factors = range(size)
        self.rep = rep
        self.nb = ()

    def __lt__(self, other):
        return self.uid < other.uid


class QuotientGraph:
    """The compressed graph as ``lower_compressed`` wants it: ``rvs`` / ``factors`` are class
    handles; ``var_class`` / ``factor_class`` map ground indices to them."""

    def __init__(self, rvs, factors, var_colour, factor_colours):
        self.rvs, self.factors = rvs, factors
        self.var_colour, self.factor_colours = var_colour, factor_colours

    @property
    def compression(self):
        ground = self.var_colour.size + sum(c.size for c in self.factor_colours)
        return ground / max(1, len(self.rvs) + len(self.factors))


def quotient(ga: GroundArrays, var_colour, factor_colours, ev_value=None) -> QuotientGraph:
    """Class handles of a partition (``SuperRV`` ``:8-45``, ``SuperF`` ``:133-150``): evidence
    classes carry the mean value and population variance of their members, every class its
    size, variable classes the representative's degree ``N`` and neighbour counts."""
    nv = ga.n_vars
    deg = ga.degrees()
    ncls = int(var_colour.max()) + 1
    order = np.argsort(var_colour, kind="stable")
    starts = np.searchsorted(var_colour[order], np.arange(ncls))
    sizes = np.diff(np.append(starts, nv))
    reps = order[starts]                                               # smallest ground index
    hidden = np.isnan(ga.var_value)
    vsum = np.bincount(var_colour, weights=np.where(hidden, 0.0, ga.var_value), minlength=ncls)
    mean = vsum / sizes
    dev = np.where(hidden, 0.0, ga.var_value - mean[var_colour])       # spread around the members' mean
    if ev_value:                                   # classes whose value was set by k-means (centroid)
        mean = mean.copy()
        for c, v in ev_value.items():
            mean[int(c)] = v
    variance = np.bincount(var_colour, weights=dev * dev, minlength=ncls) / sizes
    rvs = []
    for c in range(ncls):
        r = int(reps[c])
        is_hidden = bool(hidden[r])
        dom = ga.domains[int(ga.var_dom[r])]
        rvs.append(_ClassRV(c, dom, None if is_hidden else _domain_value(dom, mean[c]),
                            None if is_hidden else float(variance[c]), int(sizes[c]), int(deg[r]), r))
    factors = []
    is_rep = np.zeros(nv, dtype=bool)
    is_rep[reps] = True
    classes = {}                       # factor colour -> handle (a class may span several blocks)
    sizes_f = Counter()
    for fcol in factor_colours:
        ids, counts = np.unique(fcol, return_counts=True)
        for cid, cn in zip(ids, counts):
            sizes_f[int(cid)] += int(cn)
    for b, fcol in zip(ga.blocks, factor_colours):
        if b.n == 0:
            continue
        ids, first = np.unique(fcol, return_index=True)
        for cid, fi in zip(ids, first):
            if int(cid) not in classes:
                f = _ClassF(len(factors), b.potential, sizes_f[int(cid)], int(fi))
                f.nb = tuple(rvs[int(var_colour[v])] for v in b.args[fi])
                classes[int(cid)] = f
                factors.append(f)
        # neighbour counts of the representatives: factors of this block that touch one, counted
        # per (variable class, factor class) pair with one unique pass per argument position
        n_fc = int(fcol.max()) + 1
        for a in range(b.arity):
            col = b.args[:, a]
            hit = np.flatnonzero(is_rep[col])
            if hit.size == 0:
                continue
            pairs, cnt = np.unique(var_colour[col[hit]].astype(np.int64) * n_fc + fcol[hit], return_counts=True)
            for pk, cn in zip(pairs.tolist(), cnt.tolist()):
                rvs[pk // n_fc].count[classes[pk % n_fc]] += cn
    for rv in rvs:
        rv.nb = tuple(rv.count)
    return QuotientGraph(rvs, factors, var_colour, list(factor_colours))


def lift(ga: GroundArrays, split_cont_evidence=True) -> QuotientGraph:
    vcol, fcols, _ = colour_passing(ga, split_cont_evidence)
    return quotient(ga, vcol, fcols)


def lower_lifted(ga: GroundArrays, K, T, **kw):
    """``LiftedVarInference`` on arrays: colour passing, then the compressed lowering
    (``lowering.lower_compressed``: W_f = class size, gamma = ``count[f]`` at the first
    occurrence of the class in ``f.nb``, node terms weighted by the variable class size)."""
    q = lift(ga)
    return lowering.lower_compressed(q, K, T, **kw), q


def lower_ground_arrays(ga: GroundArrays, K, T):
    """``VarInference`` on arrays: every variable and factor its own class (all weights 1) --
    the trivial partition pushed through the same route (small models; the benchmark-scale
    ground models are emitted directly as record columns by ``synthetic.py``)."""
    vcol = np.arange(ga.n_vars, dtype=np.int64)
    fcols, off = [], 0
    for b in ga.blocks:
        fcols.append(np.arange(b.n, dtype=np.int64) + off)
        off += b.n
    q = quotient(ga, vcol, fcols)
    return lowering.lower_compressed(q, K, T), q


def _first_index(labels, n):
    """``first[c]`` = smallest position i with ``labels[i] == c`` (-1: label unused); no sort."""
    first = np.full(n, -1, dtype=np.int64)
    first[labels[::-1]] = np.arange(labels.size - 1, -1, -1, dtype=np.int64)      # the last write is the smallest index
    return first


def class_stats(ga: GroundArrays, var_colour, ev_value=None, degrees=None, base=None):
    """Per variable class, as arrays: size, representative (smallest member), hidden flag, mean
    evidence value (k-means centroid where ``ev_value`` names one), population variance around
    the members' mean, representative degree, domain (``SuperRV`` ``:8-45``)."""
    if base is not None:          # statistics of this very colouring without centroids: only the means change
        st = dict(base)
        ncls, mean = st["n"], st["mean"]
        if isinstance(ev_value, tuple):
            st["mean"] = np.where(ev_value[0][:ncls], ev_value[1][:ncls], mean)
        elif ev_value:
            st["mean"] = mean.copy()
            ids = np.fromiter(ev_value.keys(), dtype=np.int64, count=len(ev_value))
            st["mean"][ids] = np.fromiter(ev_value.values(), dtype=float, count=len(ev_value))
        return st
    nv = ga.n_vars
    ncls = int(var_colour.max()) + 1 if nv else 0
    sizes = np.bincount(var_colour, minlength=ncls)
    reps = _first_index(var_colour, ncls)                  # smallest member
    hidden = np.isnan(ga.var_value)
    mean = np.bincount(var_colour, weights=np.where(hidden, 0.0, ga.var_value), minlength=ncls) / sizes
    dev = np.where(hidden, 0.0, ga.var_value - mean[var_colour])
    variance = np.bincount(var_colour, weights=dev * dev, minlength=ncls) / sizes
    if isinstance(ev_value, tuple):                       # (has_centroid [ncls], centroid [ncls])
        mean = np.where(ev_value[0][:ncls], ev_value[1][:ncls], mean)
    elif ev_value:
        mean = mean.copy()
        ids = np.fromiter(ev_value.keys(), dtype=np.int64, count=len(ev_value))
        mean[ids] = np.fromiter(ev_value.values(), dtype=float, count=len(ev_value))
    return dict(n=ncls, size=sizes, rep=reps, hidden=hidden[reps], mean=mean, variance=variance,
                degree=(ga.degrees() if degrees is None else degrees)[reps], dom=ga.var_dom[reps].astype(np.int64))


def slot_layout(K, dom, domains):
    """Parameter slots of hidden variable classes with domains ``dom`` (indices into ``domains``),
    in the given order: ``(kind, dim, off, n_param)`` as ``lowering.lower_graph`` lays them out."""
    cont = np.array([bool(d.continuous) for d in domains], dtype=bool)
    ddim = np.array([2 if d.continuous else len(d.values) for d in domains], dtype=np.int64)
    if np.any(~cont & (ddim > lowering.MAX_DSTATES)):
        raise ValueError(f"discrete variable with more than {lowering.MAX_DSTATES} states")
    dim = ddim[dom]
    n = K * dim
    slot = np.where(n <= 2, 2, (n + 3) // 4 * 4)
    off = np.cumsum(slot) - slot
    return (~cont[dom]).astype(np.uint8), dim.astype(np.int32), off.astype(np.int32), int(slot.sum())


def lower_partition(ga: GroundArrays, var_colour, factor_colours, K, T, *, ev_value=None,
                    gaussian_obs=False, min_obs_var=0.0, degrees=None, stats=None) -> lowering.LoweredModel:
    """``lowering.lower_compressed(quotient(...))`` without an object per class: the record
    columns of the compressed model straight from the partition arrays (same groups, same
    records in the same order, same coefficient table -- ``tests/test_lifting.py`` compares the
    two column by column).  Weights as in ``LiftedVarInference.py:74,90,111-112,131-132``:
    W_f = factor class size, gamma = incidences of the class's factors on the variable class's
    representative, at the first position of that class in the representative factor's
    argument list; node terms scaled by the variable class size (energy, G_w) and 1 (gradients).
    ``model.slot_class[i]`` is the variable class behind parameter slot i (``handles`` is empty)."""
    if not 1 <= K <= lowering.MAX_K:
        raise ValueError(f"num_mixtures must be in 1..{lowering.MAX_K}, got {K}")
    if not 1 <= T <= lowering.MAX_T:
        raise ValueError(f"num_quadrature_points must be in 1..{lowering.MAX_T}, got {T}")
    HD, HC, EG, EC, ED = lowering.HD, lowering.HC, lowering.EG, lowering.EC, lowering.ED
    var_colour = np.asarray(var_colour, dtype=np.int64)
    st = class_stats(ga, var_colour, ev_value, degrees, base=stats)
    ncls, mean, variance = st["n"], st["mean"], st["variance"]
    cont_dom = np.array([bool(d.continuous) for d in ga.domains], dtype=bool)
    cls_cont = cont_dom[st["dom"]]
    cls_hidden = st["hidden"]
    cls_gauss = ~cls_hidden & (variance > min_obs_var) if gaussian_obs else np.zeros(ncls, dtype=bool)
    role_of = np.where(cls_hidden, np.where(cls_cont, HC, HD),
                       np.where(cls_gauss, EG, np.where(cls_cont, EC, ED))).astype(np.int64)

    slot_class = np.flatnonzero(cls_hidden)
    kind, dim, off, n_param = slot_layout(K, st["dom"][slot_class], ga.domains)
    class_off = np.full(ncls, -1, dtype=np.int64)
    class_off[slot_class] = off

    # ---- factor classes in the order the object route numbers them: block by block, ascending
    # colour inside a block, a class that spans blocks at its first one
    n_fc = max((int(c.max()) + 1 for c in factor_colours if c.size), default=0)
    size_f = np.zeros(n_fc, dtype=np.int64)
    for c in factor_colours:
        size_f += np.bincount(c, minlength=n_fc)
    # incidences of each factor class on the class representatives (rv.count[f])
    is_rep = np.zeros(ga.n_vars, dtype=bool)
    is_rep[st["rep"]] = True
    keys = []
    for b, fcol in zip(ga.blocks, factor_colours):
        for a in range(b.arity):
            col = b.args[:, a]
            hit = np.flatnonzero(is_rep[col])
            if hit.size:
                keys.append(var_colour[col[hit]] * n_fc + fcol[hit])
    if keys:
        cnt_key, cnt_val = np.unique(np.concatenate(keys), return_counts=True)
    else:
        cnt_key, cnt_val = np.zeros(0, np.int64), np.zeros(0, np.int64)

    def count_of(vcls, fcls):
        if cnt_key.size == 0:
            return np.zeros(vcls.shape, dtype=float)
        k = vcls * n_fc + fcls
        at = np.minimum(np.searchsorted(cnt_key, k), cnt_key.size - 1)
        return np.where(cnt_key[at] == k, cnt_val[at], 0).astype(float)

    table = lowering.PotentialTable()
    chunks = {}                    # group key -> list of column dicts
    unary_w = np.zeros(ncls)
    unary_g = np.zeros(ncls)
    seen = np.zeros(n_fc, dtype=bool)
    uid_next = 0
    for b, fcol in zip(ga.blocks, factor_colours):
        if b.n == 0:
            continue
        first_of = _first_index(fcol, n_fc)
        ids = np.flatnonzero((first_of >= 0) & ~seen)            # classes first met in this block, ascending id
        first = first_of[ids]
        seen[ids] = True
        if ids.size == 0:
            continue
        uid = uid_next + np.arange(ids.size)
        uid_next += ids.size
        nbcls = var_colour[b.args[first]]                       # [m, arity] classes of the representative's arguments
        roles = role_of[nbcls]
        arity = b.arity
        # records that share (roles, domains of the hidden discrete arguments, discrete evidence
        # values) share a coefficient block; distinct combinations in order of first appearance
        doms = np.where((roles == HD) | (roles == ED), st["dom"][nbcls], 0)
        vals = np.where(roles == ED, mean[nbcls], 0.0)
        # one integer code per row: roles and domains packed positionally, evidence values by rank
        code = np.zeros(ids.size, dtype=np.int64)
        n_dom = len(ga.domains) + 1
        for j in range(arity):
            code = (code * 8 + roles[:, j]) * n_dom + doms[:, j]
        cols = [code]
        if np.any(roles == ED):
            cols += [np.unique(vals[:, j], return_inverse=True)[1].reshape(-1).astype(np.int64) for j in range(arity)]
        inv = _rank_rows(cols)
        where = _first_index(inv, int(inv.max()) + 1)
        combo = np.concatenate([roles.astype(float), doms.astype(float), vals], axis=1)[where]
        for ci in np.argsort(where, kind="stable"):
            sel = np.flatnonzero(inv == ci)
            r = [int(x) for x in combo[ci, :arity]]
            pos = {q: [i for i in range(arity) if r[i] == q] for q in (HD, HC, EG, EC)}
            nd, nc, ng, ne = (len(pos[q]) for q in (HD, HC, EG, EC))
            if nd + nc + ng > lowering.MAX_ARITY:
                raise ValueError(f"factor with {nd + nc + ng} integrated arguments (max {lowering.MAX_ARITY})")
            args = []
            for i in range(arity):
                if r[i] == HD:
                    args.append(tuple(ga.domains[int(combo[ci, arity + i])].values))
                elif r[i] == ED:
                    args.append(_domain_value(ga.domains[int(combo[ci, arity + i])], combo[ci, 2 * arity + i]))
                else:
                    args.append(None)
            dims = tuple(len(args[i]) for i in pos[HD])
            hid = pos[HD] + pos[HC]
            nb = nbcls[sel]
            w_f = size_f[ids[sel]].astype(float)
            gam = np.zeros((len(hid), sel.size))
            for j, i in enumerate(hid):
                is_first = np.ones(sel.size, dtype=bool)
                for e in range(i):
                    is_first &= nb[:, e] != nb[:, i]
                gam[j] = np.where(is_first, count_of(nb[:, i], ids[sel]), 0.0)
            pure = (nd + nc + ng == 0) or (nd + nc == 1 and ng == 0)
            if pure and hid:
                np.add.at(unary_w, nb[:, hid[0]], w_f)
                np.add.at(unary_g, nb[:, hid[0]], gam[0])
            pot = table.block(b.potential, r, args)
            pkind = lowering.potential_kind(b.potential, nc + ng + ne)
            chunks.setdefault((nd, nc, ng, ne, dims, False, pure, pkind), []).append(dict(
                uid=uid[sel], pot=np.full(sel.size, pot, dtype=np.int32),
                poff=class_off[nb[:, hid]].T.reshape(len(hid), sel.size),
                egval=mean[nb[:, pos[EG]]].T.reshape(ng, sel.size),
                egvar=variance[nb[:, pos[EG]]].T.reshape(ng, sel.size),
                ecval=mean[nb[:, pos[EC]]].T.reshape(ne, sel.size),
                wf=w_f, gam=gam, nscale=np.zeros(sel.size)))

    # ---- node-entropy records, one per hidden class and per Gaussian-evidence class
    scale = (st["degree"] - 1).astype(float)
    c_v = st["size"].astype(float)
    for is_disc, d in sorted({(int(k_), int(d_)) for k_, d_ in zip(kind, dim)}):
        pick = slot_class[(kind == is_disc) & (dim == d)]
        key = (1, 0, 0, 0, (d,), True, False, 0) if is_disc else (0, 1, 0, 0, (), True, False, 0)
        chunks.setdefault(key, []).append(dict(
            uid=pick, pot=np.zeros(pick.size, dtype=np.int32), poff=class_off[pick][None, :],
            egval=np.zeros((0, pick.size)), egvar=np.zeros((0, pick.size)), ecval=np.zeros((0, pick.size)),
            wf=c_v[pick] * scale[pick] - unary_w[pick], gam=np.ones((1, pick.size)),
            nscale=scale[pick] - unary_g[pick]))
    pick = np.flatnonzero(cls_gauss)
    if pick.size:
        chunks.setdefault((0, 0, 1, 0, (), True, False, 0), []).append(dict(
            uid=pick, pot=np.zeros(pick.size, dtype=np.int32), poff=np.zeros((0, pick.size), dtype=np.int64),
            egval=mean[pick][None, :], egvar=variance[pick][None, :], ecval=np.zeros((0, pick.size)),
            wf=c_v[pick] * scale[pick], gam=np.zeros((0, pick.size)), nscale=scale[pick]))

    groups = []
    for (nd, nc, ng, ne, dims, node, pure, pkind), parts in chunks.items():
        cat = {k: np.concatenate([p[k] for p in parts], axis=-1) for k in parts[0]}
        order = np.argsort(cat["uid"], kind="stable")
        wf, gam = cat["wf"][order], cat["gam"][:, order]
        weighted = bool(node or np.any(wf != 1.0) or np.any(gam != 1.0))
        groups.append(lowering.RecordGroup(
            nd, nc, ng, ne, dims, node, cat["pot"][order].astype(np.int32),
            np.ascontiguousarray(cat["poff"][:, order].astype(np.int32)),
            np.ascontiguousarray(cat["egval"][:, order]), np.ascontiguousarray(cat["egvar"][:, order]),
            np.ascontiguousarray(cat["ecval"][:, order]), wf, np.ascontiguousarray(gam),
            cat["nscale"][order], weighted, pure, (), pkind))
    groups.sort(key=lambda g: (not g.node, not g.pure, g.nd + g.nc + g.ng, g.nd, g.nc, g.ng, g.ne, g.dims, g.kind))
    return lowering.LoweredModel(K, T, max(n_param, 2), kind, dim, off, table.array(), groups, [], {},
                                 slot_class.astype(np.int64))


def arrays_from_graph(g) -> tuple:
    """``GroundArrays`` of an object graph (any ``Graph`` of this repo or of the reference):
    variables in id order, one block per potential object.  Returns ``(arrays, rvs)`` with
    ``rvs[i]`` the object behind variable index i."""
    rvs = sorted(g.rvs, key=lambda rv: rv.id) if all(hasattr(rv, "id") for rv in g.rvs) else list(g.rvs)
    index = {id(rv): i for i, rv in enumerate(rvs)}
    domains, dom_index = [], {}
    var_dom = np.zeros(len(rvs), dtype=np.int32)
    var_value = np.full(len(rvs), np.nan)
    for i, rv in enumerate(rvs):
        if id(rv.domain) not in dom_index:
            dom_index[id(rv.domain)] = len(domains)
            domains.append(rv.domain)
        var_dom[i] = dom_index[id(rv.domain)]
        if rv.value is not None:
            var_value[i] = float(rv.value)
    by_pot = {}
    factors = sorted(g.factors, key=lambda f: f.id) if all(hasattr(f, "id") for f in g.factors) else list(g.factors)
    for f in factors:
        by_pot.setdefault(id(f.potential), (f.potential, []))[1].append([index[id(rv)] for rv in f.nb])
    blocks = [FactorBlock(pot, np.asarray(rows, dtype=np.int64)) for pot, rows in by_pot.values()]
    return GroundArrays(domains, var_dom, var_value, blocks), rvs


class PartitionInfo:
    """What callers ask of a partition: the colourings and the compression ratio."""

    def __init__(self, var_colour, factor_colours):
        self.var_colour, self.factor_colours = var_colour, list(factor_colours)

    @property
    def n_var_classes(self):
        return int(self.var_colour.max()) + 1 if self.var_colour.size else 0

    @property
    def n_factor_classes(self):
        return max((int(c.max()) + 1 for c in self.factor_colours if c.size), default=0)

    @property
    def compression(self):
        ground = self.var_colour.size + sum(c.size for c in self.factor_colours)
        return ground / max(1, self.n_var_classes + self.n_factor_classes)


def slot_elements(off, n):
    """Flat element indices of blocks of ``n[i]`` elements starting at ``off[i]``, concatenated."""
    off = np.asarray(off, dtype=np.int64)
    n = np.asarray(n, dtype=np.int64)
    start = np.cumsum(n) - n
    return np.repeat(off - start, n) + np.arange(int(n.sum()), dtype=np.int64)


def softmax_slots(tau, eta, K, kind, dim, off):
    """``eta = softmax(tau)`` row by row in every discrete slot (``VarInference.py:210-213``)."""
    for d in np.unique(dim[kind == 1]):
        o = off[(kind == 1) & (dim == d)].astype(np.int64)
        idx = o[:, None, None] + (np.arange(K) * int(d))[None, :, None] + np.arange(int(d))[None, None, :]
        e = np.e ** tau[idx]
        eta[idx] = e / e.sum(axis=2, keepdims=True)


def expand_map(engine, model, var_colour, ga: GroundArrays):
    """MAP value of every ground variable (reference ``VarInference.map``, ``:355-376``): one batched
    device call over the hidden classes (``lhvi_mixture_map``: safeguarded Newton from the best
    component mean; arg-max state for discrete classes), expanded to the members; observed variables
    carry their evidence value.  Discrete entries are the domain values themselves."""
    res = engine.mixture_map()
    res = np.asarray(res.double().cpu().numpy() if hasattr(res, "cpu") else res, dtype=float)
    n_cls = int(var_colour.max()) + 1 if var_colour.size else 0
    slot_of = np.full(n_cls, -1, dtype=np.int64)
    slot_of[model.slot_class] = np.arange(model.slot_class.size)
    out = np.array(ga.var_value, dtype=float)
    hidden = np.flatnonzero(np.isnan(ga.var_value))
    slots = slot_of[var_colour[hidden]]
    val = res[slots]
    disc = model.var_kind[slots] == 1
    if disc.any():                                   # state index -> domain value
        for d in np.unique(ga.var_dom[hidden[disc]]):
            values = np.asarray(ga.domains[int(d)].values, dtype=float)
            pick = disc & (ga.var_dom[hidden] == d)
            val[pick] = values[val[pick].astype(np.int64)]
    out[hidden] = val
    return out


class ArrayVI:
    """``LiftedVarInference`` (``lifted=True``) or ``VarInference`` over a ``GroundArrays`` model:
    colour passing on arrays, the array-native compressed lowering (``lower_partition``), and
    the device engine (``LiftedVarInference.py:14-26`` + ``VarInference.run`` ``:215-247``).
    Per-variable results are expanded from the classes back to the ground variables."""

    def __init__(self, ga: GroundArrays, K, T, *, lifted=True, dtype="float64", device=None, engine_factory=None,
                 device_passes=None):
        self.ga, self.K, self.T = ga, K, T
        # (a single colour passing: the host library is as fast as the device passes' first-use costs below
        # ~10^6 factors, so the resident passes are opt-in here; C2FArrayVI, which refines every ten
        # iterations, uses them by default)
        dev = _passes_device(False if device_passes is None else device_passes, device, engine_factory) if lifted else None
        if dev is not None:
            # colour passing, class statistics and lowering on the resident ground graph (lifting_torch):
            # the host library's class ids and record columns
            import torch
            from . import lifting_torch as lt
            tg = getattr(ga, "_torch_graph", None)
            if tg is None or tg.device != dev:
                tg = ga._torch_graph = lt.TorchGraph(ga, dev)
            cont, _ = lt.domain_tables(ga.domains, dev)
            vt, ft, _ = lt.colour_passing(tg, lt.initial_colouring(tg, cont, True))
            stats = lt.class_stats(tg, vt)
            self.model = lt.lower_partition(tg, ga, vt, ft, K, T, stats=stats)
            vcol, fcols = vt.cpu().numpy(), [f.cpu().numpy() for f in ft]
            self.class_rep = stats["rep"].cpu().numpy()
        else:
            if lifted:
                vcol, fcols, _ = colour_passing(ga)
            else:                                   # every variable and factor its own class
                vcol = np.arange(ga.n_vars, dtype=np.int64)
                sizes = np.cumsum([0] + [b.n for b in ga.blocks])
                fcols = [np.arange(b.n, dtype=np.int64) + o for b, o in zip(ga.blocks, sizes)]
            self.class_rep = class_stats(ga, vcol)["rep"]
            self.model = lower_partition(ga, vcol, fcols, K, T)
        self.quotient = PartitionInfo(vcol, fcols)
        if engine_factory is not None:
            self.engine = engine_factory(self.model)
        else:
            from .engine import DeviceEngine
            self.engine = DeviceEngine(self.model, dtype=dtype, device=device)
        self.init_param(0)

    def init_param(self, seed=0):
        """The reference's initial distributions (``VarInference.py:197-208``), drawn per class
        from a generator seeded by the class representative's ground index -- so a lifted and a
        ground run of the same model start from corresponding points."""
        m, K = self.model, self.K
        eta = np.zeros(m.n_param)
        tau = np.zeros(m.n_param)
        reps = self.class_rep[m.slot_class]
        if self.canonical is not None:
            reps = self.canonical[reps]
        for rep, off, kind, dim in zip(reps, m.var_off, m.var_kind, m.var_dim):
            rng = np.random.default_rng([seed, int(rep)])
            if kind == 0:
                eta[off:off + 2 * K:2] = rng.random(K) * 3 - 1.5
                eta[off + 1:off + 2 * K:2] = 1.0
            else:
                tau[off:off + K * dim] = (rng.random((K, dim)) * 10).reshape(-1)
        softmax_slots(tau, eta, K, m.var_kind, m.var_dim, m.var_off)
        self.engine.set_state(eta, tau, np.zeros(K))
        self.engine.reset_moments()

    canonical = None        # optional ground-variable -> canonical representative map (see tie_to)

    def tie_to(self, other: "ArrayVI"):
        """Start from the point that corresponds to ``other``'s (a coarser partition of the same
        model): every class here draws with the representative of the class of ``other`` it lies in."""
        self.canonical = other.class_rep[other.quotient.var_colour]
        self.init_param(0)

    def run(self, iteration=100, lr=0.1):
        self.engine.iterate(int(iteration), float(lr))

    def free_energy(self):
        return self.engine.free_energy()

    def map_values(self):
        """MAP value of every ground variable as one array (see ``expand_map``)."""
        return expand_map(self.engine, self.model, self.quotient.var_colour, self.ga)

    def ground_params(self):
        """``eta`` of every hidden ground variable: dict ground index -> ``[K, 2]`` (continuous) or
        ``[K, D]`` array, expanded from the classes; and the mixture weights."""
        eta, _, _, w = self.engine.get_state()
        m, K = self.model, self.K
        slot_of = np.full(self.quotient.n_var_classes, -1, dtype=np.int64)
        slot_of[m.slot_class] = np.arange(m.slot_class.size)
        tables = [eta[off:off + K * dim].reshape(K, dim) for off, dim in zip(m.var_off, m.var_dim)]
        col = self.quotient.var_colour
        hidden = np.flatnonzero(np.isnan(self.ga.var_value))
        return {int(v): tables[slot_of[col[v]]] for v in hidden}, w


def _passes_device(device_passes, device, engine_factory, k_mean_k=2):
    """Where the lifting passes run: ``None`` -> on the GPU next to the kernels when there is one and the
    engine is the real one (``lifting_torch``), else in the host library; ``False`` -> host; ``True`` -> the
    engine's GPU; anything else -> that torch device (``"cpu"`` in the tests)."""
    if device_passes is False:
        return None
    import torch
    if device_passes is None:
        if engine_factory is not None or not torch.cuda.is_available() or k_mean_k != 2:
            return None
        device_passes = True
    if device_passes is True:
        return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    return torch.device(device_passes)


def _kmeans_1d(values, k, iteration):
    """The reference's evidence k-means (``CompressedGraphWithObs.py:78-130``) on the members'
    values in member order: centroids start at the first ``k`` distinct values, ``iteration``
    Lloyd sweeps over the value histogram, members go to their nearest centroid.  Returns
    ``(owner [n], centroids [k])`` or ``None`` when there is nothing to split."""
    uniq, first, counts = np.unique(values, return_index=True, return_counts=True)
    k = min(k, uniq.size)
    if values.size <= 1 or k <= 1:
        return None
    seen = np.argsort(first, kind="stable")              # histogram in first-seen (member) order
    vals, cnts = uniq[seen].astype(float), counts[seen].astype(float)
    centroids = vals[:k].copy()
    for _ in range(iteration):
        owner = np.abs(vals[:, None] - centroids[None, :]).argmin(axis=1)
        mass = np.bincount(owner, weights=cnts, minlength=k)
        tot = np.bincount(owner, weights=vals * cnts, minlength=k)
        with np.errstate(invalid="ignore", divide="ignore"):
            centroids = tot / mass
    owner = np.abs(values[:, None] - centroids[None, :]).argmin(axis=1)
    return owner, centroids


class C2FArrayVI:
    """``C2FVarInference`` over a ``GroundArrays`` model (``C2FVarInference.py:301-352``).

    Continuous evidence starts lumped into one class per domain and is integrated as a fixed
    Gaussian (class mean, class variance); every ``update_obs_its`` iterations the evidence
    classes whose spread exceeds a shrinking threshold are split by 1-D k-means, colour passing
    refines the partition, the pieces of a hidden class inherit its parameters and Adam moments,
    the compressed model is lowered and uploaded again, and the kernels continue.  The Adam step
    counter runs across the rounds.  Class attributes are the reference's knobs."""

    k_mean_k = 2
    k_mean_its = 10
    update_obs_its = 10
    output_its = 0
    min_obs_var = 0
    gaussian_obs = True
    var_threshold = 0.1

    def __init__(self, ga: GroundArrays, K, T, *, dtype="float64", device=None, engine_factory=None,
                 init_fn=None, use_native=None, device_passes=None):
        self.ga, self.K, self.T = ga, int(K), int(T)
        # where the lifting passes of a round run (evidence split, colour passing, class statistics,
        # inheritance, lowering): None -> on the GPU next to the kernels when there is one and the engine is
        # the real one (``lifting_torch``: the ground graph is uploaded once and stays resident), else in the
        # host library; False -> host; a torch device (e.g. "cpu" in the tests) -> there
        self.device_passes = device_passes
        self.dtype, self.device = dtype, device
        self.engine_factory = engine_factory
        self.init_fn = init_fn            # (representative ground index, is_continuous, dim) -> [K, dim] or None
        self.layout = None
        self.hidden = np.isnan(ga.var_value)
        self.cont_dom = np.array([bool(d.continuous) for d in ga.domains])[ga.var_dom]
        self.t = 0.0
        self.use_native = use_native      # None: C++ passes when built; False: numpy passes
        self.history = []                 # (number of variable classes, free energy) per round

    # ---- engine ---------------------------------------------------------------------------
    def _make_engine(self, model):
        if self.engine_factory is not None:
            return self.engine_factory(model)
        from .engine import DeviceEngine
        return DeviceEngine(model, dtype=self.dtype, device=self.device, var_threshold=self.var_threshold)

    # ---- partition bookkeeping -------------------------------------------------------------
    def _evidence_stats(self):
        """(class ids, mean, variance, size) of the evidence classes of the current partition."""
        ev = np.flatnonzero(~self.hidden)
        col = self.vcol[ev]
        ids, inv, cnt = np.unique(col, return_inverse=True, return_counts=True)
        mean = np.bincount(inv, weights=self.ga.var_value[ev]) / cnt
        dev = self.ga.var_value[ev] - mean[inv]
        var = np.bincount(inv, weights=dev * dev) / cnt
        return ids, mean, var, cnt

    def _split_evidence(self, epsilon):
        """``CompressedGraph.split_evidence`` (``:236-247``) until nothing changes.  Book-keeping per
        class id: ``may_split`` (the reference's ``clustered_evidence``), ``ev_has`` / ``ev_val`` (classes
        whose value is a k-means centroid rather than the members' mean).  Runs in the host-side C++
        library when it is built (``lhvi_lift_split_evidence``), else in ``_split_evidence_numpy``;
        the two give the same arrays bit for bit (``tests/test_lifting.py``)."""
        n = int(self.vcol.max()) + 1
        room = n + int(np.count_nonzero(self.may_split[self.vcol]))     # every new class has a member
        may = np.zeros(room, dtype=np.uint8)
        has = np.zeros(room, dtype=np.uint8)
        val = np.zeros(room)
        may[:n], has[:n], val[:n] = self.may_split, self.ev_has, self.ev_val
        lib = _lift_native.load() if self.use_native is not False else None
        if lib is not None:
            vcol = np.ascontiguousarray(self.vcol, dtype=np.int64)
            n_new = _lift_native.split_evidence(lib, vcol, np.ascontiguousarray(self.ga.var_value, dtype=np.float64),
                                                n, may, has, val, epsilon, self.k_mean_k, self.k_mean_its)
            self.vcol = vcol
        else:
            n_new = self._split_evidence_numpy(epsilon, may, has, val)
        self.may_split, self.ev_has, self.ev_val = may[:n_new].astype(bool), has[:n_new].astype(bool), val[:n_new].copy()

    def _split_evidence_numpy(self, epsilon, may, has, val):
        changed = True
        next_id = int(self.vcol.max()) + 1
        while changed:
            changed = False
            n_now = next_id
            if not may[:n_now].any():
                break
            # members of the classes k-means may still split, grouped by class (classes are
            # disjoint: one grouping per pass); classes whose spread is within epsilon are skipped
            # after one vectorised variance pass
            cand = np.flatnonzero(may[self.vcol] != 0)
            cand = cand[np.argsort(self.vcol[cand], kind="stable")]
            ccol = self.vcol[cand]
            cids, cstart, ccnt = np.unique(ccol, return_index=True, return_counts=True)
            cval = self.ga.var_value[cand]
            cmean = np.add.reduceat(cval, cstart) / ccnt
            cdev = cval - np.repeat(cmean, ccnt)
            wide = np.sqrt(np.add.reduceat(cdev * cdev, cstart) / ccnt) > epsilon * (1 - 1e-9)
            wide &= np.maximum.reduceat(cval, cstart) > np.minimum.reduceat(cval, cstart)   # nothing to split otherwise
            for j in np.flatnonzero(wide):
                cid = int(cids[j])
                members = cand[cstart[j]:cstart[j] + ccnt[j]]
                vals = cval[cstart[j]:cstart[j] + ccnt[j]]
                if not np.sqrt(vals.var()) > epsilon:
                    continue
                res = _kmeans_1d(vals, self.k_mean_k, self.k_mean_its)
                if res is None:
                    if members.size == 1:
                        may[cid] = 0
                    continue
                owner, centroids = res
                pieces = [(cid, vals[owner == 0])]
                has[cid], val[cid] = 1, centroids[0]
                for c in range(1, centroids.size):
                    pick = owner == c
                    if pick.any():
                        self.vcol[members[pick]] = next_id
                        has[next_id], val[next_id], may[next_id] = 1, centroids[c], 0
                        pieces.append((next_id, vals[pick]))
                        next_id += 1
                if len(pieces) > 1:
                    changed = True
                    for pid, pv in pieces:
                        if pv.size and pv.var() > epsilon:
                            may[pid] = 1
                        elif pid != cid:
                            may[pid] = 0
        return next_id

    def _refine(self):
        """Colour passing from the current classes; hidden pieces inherit (``:39-61``)."""
        old = self.vcol
        new, self.fcols, _ = colour_passing(self.ga, start=old, use_native=self.use_native)
        # (old class, new class) pairs: the new partition refines the old one, so every new class
        # has one parent -- the old class of any of its members
        n_old, n_new = int(old.max()) + 1, int(new.max()) + 1
        parent = old[_first_index(new, n_new)]
        by_parent = np.argsort(parent, kind="stable")          # pairs ordered by (old, new)
        pair = np.stack([parent[by_parent], by_parent])
        kids_of = np.bincount(pair[0], minlength=n_old)
        # carry the evidence book-keeping over to the new ids: every piece of a class k-means may
        # split may be split; a class that colour passing left whole keeps its k-means centroid as
        # value, pieces of a structure split take the mean of their members
        # (SuperRV.split_by_structure, :57-62)
        may = np.zeros(n_new, dtype=bool)
        may[pair[1][self.may_split[pair[0]]]] = True
        first_kid = np.full(n_old, -1, dtype=np.int64)
        first_kid[pair[0][::-1]] = pair[1][::-1]
        keep = np.flatnonzero(self.ev_has[:n_old] & (kids_of == 1))
        has, val = np.zeros(n_new, dtype=bool), np.zeros(n_new)
        has[first_kid[keep]] = True
        val[first_kid[keep]] = self.ev_val[keep]
        self.may_split, self.ev_has, self.ev_val = may, has, val
        self._inherit(old, new)
        self.vcol = new

    def _layout(self, vcol):
        """Parameter slots of the hidden classes of a partition, in class order."""
        st = self.stats = class_stats(self.ga, vcol, degrees=self.degrees)     # of the colouring laid out last
        cls = np.flatnonzero(st["hidden"])
        kind, dim, off, n_param = slot_layout(self.K, st["dom"][cls], self.ga.domains)
        slot_of = np.full(st["n"], -1, dtype=np.int64)
        slot_of[cls] = np.arange(cls.size)
        return dict(cls=cls, rep=st["rep"][cls], kind=kind, dim=dim, off=off, n_param=max(n_param, 2), slot_of=slot_of)

    def _inherit(self, old, new):
        """The pieces of a hidden class start from its parameters and Adam moments
        (``C2FVarInference.py:39-61``): one gather from the old slot layout into the new one."""
        if self.layout is None:
            return
        lay_old, lay = self.layout, self._layout(new)
        parent = lay_old["slot_of"][old[lay["rep"]]]
        n = self.K * lay["dim"].astype(np.int64)
        dst = slot_elements(lay["off"], n)
        src = slot_elements(lay_old["off"][parent], n)
        for name in ("P", "m1", "m2"):
            fresh = np.zeros(lay["n_param"])
            fresh[dst] = getattr(self, name)[src]
            setattr(self, name, fresh)
        self.layout = lay

    # ---- parameters -----------------------------------------------------------------------
    def _init_params(self):
        """One table per initial hidden class (one class per domain), ``VarInference.py:197-208``:
        ``P`` holds (mu, var) pairs of continuous classes and the logits of discrete ones, in the
        slot layout of the current partition; ``m1`` / ``m2`` are the Adam moments beside it."""
        self._init_params_from(self._layout(self.vcol))

    def _init_params_from(self, lay):
        K = self.K
        self.layout = lay
        self.P = np.zeros(lay["n_param"])
        for r, kind, dim, off in zip(lay["rep"], lay["kind"], lay["dim"], lay["off"]):
            r, dim, cont = int(r), int(dim), kind == 0
            table = self.init_fn(r, cont, dim) if self.init_fn is not None else None
            if table is None:
                rng = np.random.default_rng([0, r])
                if cont:
                    table = np.ones((K, 2))
                    table[:, 0] = rng.random(K) * 3 - 1.5
                else:
                    table = rng.random((K, dim)) * 10          # logits
            self.P[off:off + K * dim] = np.asarray(table, dtype=float).reshape(-1)
        self.m1, self.m2 = np.zeros_like(self.P), np.zeros_like(self.P)
        self.w_tau = np.zeros(K)
        self.w = np.full(K, 1.0 / K)
        self.m_w, self.u_w = np.zeros(K), np.zeros(K)
        self.t = 0.0

    def _discrete_elements(self, lay):
        disc = lay["kind"] == 1
        return slot_elements(lay["off"][disc], self.K * lay["dim"][disc].astype(np.int64))

    def _push(self, model, engine):
        lay = self.layout
        assert np.array_equal(lay["cls"], model.slot_class) and np.array_equal(lay["off"], model.var_off)
        eta, tau = self.P.copy(), np.zeros_like(self.P)
        d = self._discrete_elements(lay)
        tau[d] = self.P[d]
        softmax_slots(tau, eta, self.K, lay["kind"], lay["dim"], lay["off"])
        engine.set_state(eta, tau, self.w_tau)
        engine.set_moments(self.m1, self.m2, self.m_w, self.u_w, self.t)

    def _pull(self, model, engine):
        eta, tau, w_tau, w = engine.get_state()
        m1, m2, m_w, u_w, t = engine.get_moments()
        d = self._discrete_elements(self.layout)
        self.P = np.array(eta, dtype=float)
        self.P[d] = np.asarray(tau)[d]
        self.m1, self.m2 = np.array(m1, dtype=float), np.array(m2, dtype=float)
        self.w_tau, self.w = w_tau, w
        self.m_w, self.u_w, self.t = m_w, u_w, t

    # ---- the run ---------------------------------------------------------------------------
    def _passes_device(self):
        return _passes_device(self.device_passes, self.device, self.engine_factory, self.k_mean_k)

    def run(self, iteration=100, lr=0.1, log_fe=False):
        """``log_fe=True`` evaluates the free energy at the end of every refinement round (one more
        pass over the records) into ``history`` -- the reference logs it after every iteration
        (``C2FVarInference.py:354-377``)."""
        dev = self._passes_device()
        if dev is not None:
            return self._run_resident(dev, iteration, lr, log_fe)
        ga = self.ga
        self.degrees = ga.degrees()
        # initial classes, parameters per initial class (one hidden class per domain), then the
        # first colour passing in which the pieces inherit (C2FVarInference.py:306-311)
        self.vcol = initial_colouring(ga, split_cont_evidence=False)
        n0 = int(self.vcol.max()) + 1
        self.may_split = np.zeros(n0, dtype=bool)                       # what k-means may split
        self.may_split[self.vcol[~self.hidden & self.cont_dom]] = True
        self.ev_has, self.ev_val = np.zeros(n0, dtype=bool), np.zeros(n0)
        self._init_params()
        self._refine()
        _, _, var, _ = self._evidence_stats()
        epsilon = float(np.sqrt(var.max())) if var.size else 0.0
        d = epsilon * self.update_obs_its / (iteration - self.output_its)
        epsilon -= d
        self.history = []
        self.timing = {"split": 0.0, "refine": 0.0, "lower": 0.0, "upload": 0.0, "iterate": 0.0, "pull": 0.0}

        self.timing_rounds = []           # the same per refinement round

        def clock(phase, t0):
            now = time.perf_counter()
            self.timing[phase] += now - t0
            self.timing_rounds[-1][phase] = now - t0
            return now
        for _ in range(int(iteration / self.update_obs_its)):          # remainder dropped (H10)
            self.timing_rounds.append({})
            t = time.perf_counter()
            self._split_evidence(epsilon)
            t = clock("split", t)
            self._refine()
            t = clock("refine", t)
            epsilon = max(epsilon - d, self.min_obs_var)
            self.quotient = PartitionInfo(self.vcol, self.fcols)
            self.model = lower_partition(ga, self.vcol, self.fcols, self.K, self.T, ev_value=(self.ev_has, self.ev_val),
                                         gaussian_obs=self.gaussian_obs, min_obs_var=self.min_obs_var,
                                         degrees=self.degrees, stats=self.stats)
            t = clock("lower", t)
            old = getattr(self, "engine", None)
            if old is not None and hasattr(old, "close"):
                old.close()                      # graphs, peer buffers of the previous round's engine
            self.engine = self._make_engine(self.model)
            self._push(self.model, self.engine)
            t = clock("upload", t)
            self.engine.iterate(self.update_obs_its, lr)
            if hasattr(self.engine, "synchronize"):
                self.engine.synchronize()
            t = clock("iterate", t)
            self._pull(self.model, self.engine)
            clock("pull", t)
            self.history.append((self.quotient.n_var_classes, float(self.engine.free_energy()) if log_fe else None))
        return self

    def _run_resident(self, dev, iteration, lr, log_fe):
        """``run`` with the lifting passes on ``dev`` (``lifting_torch``): the ground graph, the colourings
        and the evidence book-keeping stay there between the rounds; the parameters and Adam moments
        (one slot per hidden class) travel through the host as before.  Same partitions, class ids and
        record columns as the host route (``tests/test_lifting_torch.py``)."""
        import torch
        from . import lifting_torch as lt
        ga, K = self.ga, self.K
        sync = (lambda: torch.cuda.synchronize(dev)) if dev.type == "cuda" else (lambda: None)
        t_setup = time.perf_counter()
        tg = getattr(ga, "_torch_graph", None)
        if tg is None or tg.device != dev:
            tg = ga._torch_graph = lt.TorchGraph(ga, dev)
        cont, ddim = lt.domain_tables(ga.domains, dev)
        self.degrees = None
        to_host = lambda lay: {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in lay.items()}

        def refine(vcol, may, has, val, lay_t):
            new, fcols, _ = lt.colour_passing(tg, vcol)
            may, has, val = lt.refine_bookkeeping(vcol, new, may, has, val)
            new_lay_t, stats = lt.layout(tg, new, K, cont, ddim)
            if lay_t is not None:
                arrs = lt.inherit(lay_t, new_lay_t, vcol, K, [torch.as_tensor(a).to(dev) for a in (self.P, self.m1, self.m2)])
                self.P, self.m1, self.m2 = (a.cpu().numpy() for a in arrs)
            return new, fcols, may, has, val, new_lay_t, stats

        vcol = lt.initial_colouring(tg, cont, split_cont_evidence=False)
        n0 = int(vcol.max()) + 1
        may = torch.zeros(n0, dtype=torch.bool, device=dev)
        may[vcol[~tg.hidden & cont[tg.var_dom]]] = True
        has, val = torch.zeros(n0, dtype=torch.bool, device=dev), torch.zeros(n0, dtype=torch.float64, device=dev)
        # parameters per initial hidden class (C2FVarInference.py:306-311), then the first colour passing
        lay_t, _ = lt.layout(tg, vcol, K, cont, ddim)
        self.vcol = vcol.cpu().numpy()
        self.stats = None
        self._init_params_from(to_host(lay_t))
        vcol, fcols, may, has, val, lay_t, stats = refine(vcol, may, has, val, lay_t)
        self.layout = to_host(lay_t)
        ev = ~tg.hidden
        if bool(ev.any()):
            ecol = vcol[ev]
            ids, inv, cnt = torch.unique(ecol, return_inverse=True, return_counts=True)
            evar = lt._segment_var(tg.var_value[ev], inv, ids.numel(), cnt)
            epsilon = float(torch.sqrt(evar.max()))
        else:
            epsilon = 0.0
        d = epsilon * self.update_obs_its / (iteration - self.output_its)
        epsilon -= d
        self.history = []
        self.timing = {"split": 0.0, "refine": 0.0, "lower": 0.0, "upload": 0.0, "iterate": 0.0, "pull": 0.0}
        self.timing_rounds = []
        sync()
        self.timing["setup"] = time.perf_counter() - t_setup     # graph upload, initial classes, first colour passing

        def clock(phase, t0):
            sync()
            now = time.perf_counter()
            self.timing[phase] += now - t0
            self.timing_rounds[-1][phase] = now - t0
            return now
        for _ in range(int(iteration / self.update_obs_its)):          # remainder dropped (H10)
            self.timing_rounds.append({})
            sync()
            t = time.perf_counter()
            vcol, may, has, val = lt.split_evidence(tg, vcol, may, has, val, epsilon, self.k_mean_k, self.k_mean_its)
            t = clock("split", t)
            vcol, fcols, may, has, val, lay_t, stats = refine(vcol, may, has, val, lay_t)
            self.layout = to_host(lay_t)
            t = clock("refine", t)
            epsilon = max(epsilon - d, self.min_obs_var)
            self.model = lt.lower_partition(tg, ga, vcol, fcols, K, self.T, ev_value=(has, val),
                                            gaussian_obs=self.gaussian_obs, min_obs_var=self.min_obs_var, stats=stats)
            t = clock("lower", t)
            old = getattr(self, "engine", None)
            if old is not None and hasattr(old, "close"):
                old.close()
            self.engine = self._make_engine(self.model)
            self._push(self.model, self.engine)
            t = clock("upload", t)
            self.engine.iterate(self.update_obs_its, lr)
            if hasattr(self.engine, "synchronize"):
                self.engine.synchronize()
            t = clock("iterate", t)
            self._pull(self.model, self.engine)
            clock("pull", t)
            n_classes = int(stats["n"])
            self.history.append((n_classes, float(self.engine.free_energy()) if log_fe else None))
        # the public attributes of a finished run, as the host route leaves them
        self.vcol = vcol.cpu().numpy()
        self.fcols = [f.cpu().numpy() for f in fcols]
        self.may_split, self.ev_has, self.ev_val = may.cpu().numpy(), has.cpu().numpy(), val.cpu().numpy()
        self.stats = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in stats.items()}
        self.quotient = PartitionInfo(self.vcol, self.fcols)
        return self

    def free_energy(self):
        return self.engine.free_energy()

    def map_values(self):
        """MAP value of every ground variable as one array (see ``expand_map``), from the engine of
        the last refinement round."""
        return expand_map(self.engine, self.model, self.vcol, self.ga)

    def class_tables(self):
        """``(slot_of_class, tables)``: ``tables[slot_of_class[c]]`` is the ``[K, dim]`` table of hidden
        class c (continuous: (mu, var); discrete: probabilities)."""
        lay, K = self.layout, self.K
        eta = self.P.copy()
        tau = np.zeros_like(eta)
        d = self._discrete_elements(lay)
        tau[d] = self.P[d]
        softmax_slots(tau, eta, K, lay["kind"], lay["dim"], lay["off"])
        return lay["slot_of"], [eta[o:o + K * n].reshape(K, n) for o, n in zip(lay["off"], lay["dim"])]

    def ground_params(self):
        """Per hidden ground variable: its class's ``[K, dim]`` table (continuous: (mu, var);
        discrete: probabilities) -- the reference's ``eta[rv.cluster]``."""
        slot_of, tables = self.class_tables()
        return {int(v): tables[slot_of[self.vcol[v]]] for v in np.flatnonzero(self.hidden)}, self.w
