"""Device-resident lifting passes ("vectorised colour passing + evidence k-means as
torch sort / unique passes"): the same partitions, class ids and statistics as ``lifting.py`` +
``csrc/host/lhvi_lift.cpp``, computed on whatever device the index arrays live on -- the GPU for
``C2FArrayVI(device_passes=True)``, the CPU in ``tests/test_lifting_torch.py``, which compares every
function here bit for bit with the host passes.

Why: ``C2FVarInference.run`` re-compresses the graph every ``update_obs_its`` iterations
(``C2FVarInference.py:301-352``).  With the variational kernels at ~0.1 ms per iteration the host passes
over the ground graph (colour passing: one hash-table sweep over every factor and every incidence per
refinement) dominate a coarse-to-fine run by two orders of magnitude.  Here the ground graph is uploaded
once and every pass is a handful of gathers, 64-bit hashes, sorts and prefix sums on resident arrays.

Conventions shared with the host passes (``lhvi_lift.cpp``): class ids are dense and numbered in order of
first appearance (variables: ascending index; factors: all blocks concatenated in block order); a factor's
key is (arity, symmetric, own class, classes of its arguments -- sorted for a symmetric potential); a
variable's key is (own class, multiset of the classes of its incident factors), the multiset compared
through two independent 64-bit sums of mixed class ids.  Keys are ranked through one 64-bit hash and then
*verified* column by column (a hash collision cannot merge classes: the exact ranking takes over).
"""
from __future__ import annotations

import numpy as np
import torch

_M64 = (1 << 64) - 1


def _i64(c):
    """Python int (uint64 constant) -> the int64 with the same bits."""
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(z, k):
    """Logical right shift of int64 tensors (torch shifts arithmetically)."""
    return (z >> k) & ((1 << (64 - k)) - 1)


def mix(x, seed):
    """splitmix64 finaliser on int64 tensors with wrap-around arithmetic: the bits of ``lifting._mix`` /
    ``lhvi_lift.cpp: mix``."""
    z = (x + _i64(seed)) * _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def first_index(labels, n):
    """``first[c]`` = smallest position i with ``labels[i] == c`` (``n`` = number of labels; unused: size)."""
    pos = torch.arange(labels.numel(), device=labels.device, dtype=torch.int64)
    return torch.full((n,), labels.numel(), dtype=torch.int64, device=labels.device).scatter_reduce_(
        0, labels, pos, "amin", include_self=True)


def relabel_first(inv, n_uniq):
    """Dense ids ``inv`` (any order) -> the same classes numbered in order of first appearance."""
    first = first_index(inv, n_uniq)
    order = torch.argsort(first)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(n_uniq, device=inv.device, dtype=torch.int64)
    return rank[inv], first[order]


def rank_first(h, cols, h2=None):
    """Dense first-appearance ids of the rows ``cols`` (int64 tensors, or a callable returning them) through
    their 64-bit hash ``h``; returns ``(ids, n_classes, first)`` with ``first[c]`` the first row of class c.
    The classes of ``h`` are verified -- every row must agree with its class's first row in an independent
    second hash ``h2`` (a false merge then needs a 128-bit collision), or, without ``h2``, in every column --
    and ranked exactly, column by column, if they do not."""
    n = h.numel()
    if n == 0:
        return h.clone(), 0, h.clone()
    uniq, inv = torch.unique(h, return_inverse=True)
    ids, first = relabel_first(inv, uniq.numel())
    if h2 is not None:
        ok = bool((h2[first][ids] == h2).all())
    else:
        ok = True
        for c in (cols() if callable(cols) else cols):
            ok = ok and bool((c[first][ids] == c).all())
    if not ok:                                      # a hash collision: rank the rows exactly, column by column
        ids = torch.zeros(n, dtype=torch.int64, device=h.device)
        for c in (cols() if callable(cols) else cols):
            _, ci = torch.unique(c, return_inverse=True)
            _, ids = torch.unique(ids * (int(ci.max()) + 1) + ci, return_inverse=True)
        ids, first = relabel_first(ids, int(ids.max()) + 1)
    return ids, int(first.numel()), first


class TorchGraph:
    """The ground graph's index arrays on ``device``, with the incidence list (one entry per argument
    position, sorted by variable) prepared once (``lhvi_lift_graph_create``)."""

    def __init__(self, ga, device):
        from .lifting import _potential_ids
        self.device = torch.device(device)
        self.n_vars = ga.n_vars
        dev = lambda a, dt=torch.int64: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(self.device)
        self.args = [dev(b.args) for b in ga.blocks]
        self.arity = [b.arity for b in ga.blocks]
        self.symmetric = [bool(getattr(b.potential, "symmetric", False)) for b in ga.blocks]
        sizes = [b.n for b in ga.blocks]
        self.foff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self.n_fac = int(self.foff[-1])
        # blocks whose potentials compare equal start in one class (lifting.colour_passing)
        first_colour = {}
        self.colour0 = [first_colour.setdefault((pid, b.arity), len(first_colour))
                        for b, pid in zip(ga.blocks, _potential_ids(ga.blocks))]
        self.header = [a | (int(s) << 8) for a, s in zip(self.arity, self.symmetric)]
        # incidences sorted by variable (stable: block order, factor order, position order -- as lhvi_lift.cpp fills them)
        if self.n_fac:
            inc_var = torch.cat([a.reshape(-1) for a in self.args])
            inc_fac = torch.cat([(torch.arange(a.shape[0], device=self.device).repeat_interleave(a.shape[1]) + int(o))
                                 for a, o in zip(self.args, self.foff[:-1])])
            order = torch.argsort(inc_var, stable=True)
            self.inc_fac = inc_fac[order]
            deg = torch.bincount(inc_var, minlength=self.n_vars)
        else:
            self.inc_fac = torch.zeros(0, dtype=torch.int64, device=self.device)
            deg = torch.zeros(self.n_vars, dtype=torch.int64, device=self.device)
        self.degree = deg
        self.inc_ptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=self.device), torch.cumsum(deg, 0)])
        self.var_value = dev(ga.var_value, torch.float64)
        self.var_dom = dev(ga.var_dom)
        self.hidden = torch.isnan(self.var_value)

    def segment_sum(self, per_incidence):
        """Wrap-around int64 sums of ``per_incidence`` (sorted by variable) per variable: prefix sums and
        differences (exact modulo 2^64; a group-level variable with 10^6 incidences costs nothing extra)."""
        cs = torch.cat([torch.zeros(1, dtype=torch.int64, device=self.device), torch.cumsum(per_incidence, 0)])
        return cs[self.inc_ptr[1:]] - cs[self.inc_ptr[:-1]]


def colour_passing(tg: TorchGraph, start, max_sweeps=1000):
    """``lhvi_lift_graph_colour_passing`` on resident arrays: the coarsest equitable refinement of the
    variable colouring ``start`` (any labels >= 0).  Returns ``(var_colour, [factor colour per block],
    sweeps)`` -- the same ids as the host library."""
    start = start.to(torch.int64)
    vcol, n_classes, _ = rank_first(mix(start, 1), [start])
    fcol = [torch.full((a.shape[0],), c, dtype=torch.int64, device=tg.device) for a, c in zip(tg.args, tg.colour0)]
    before, sweeps = -1, 0
    fid = torch.zeros(0, dtype=torch.int64, device=tg.device)
    while before != n_classes and sweeps < max_sweeps:
        before = n_classes
        sweeps += 1
        # ---- factors: (header, own class, classes of the arguments; sorted for a symmetric potential)
        hs, hs2, keys = [], [], []
        width = max([2 + a for a in tg.arity], default=2)
        for args, own, hdr, sym in zip(tg.args, fcol, tg.header, tg.symmetric):
            k = vcol[args]
            if sym and k.shape[1] > 1:
                k = torch.sort(k, dim=1).values
            h = mix(own, 0x452821E638D01377 + hdr)
            g2 = mix(own, 0xC0AC29B7C97C50DD + hdr)                 # independent second hash (verification)
            for j in range(k.shape[1]):
                h = mix(h ^ k[:, j], 0xBE5466CF34E90C6C)
                g2 = mix(g2 ^ k[:, j], 0x3F84D5B5B5470917)
            hs.append(h)
            hs2.append(g2)
            keys.append((k, own, hdr))

        def exact_rows():                                            # only built if the hashes disagree
            rows = []
            for k, own, hdr in keys:
                pad = torch.zeros((k.shape[0], width - 2 - k.shape[1]), dtype=torch.int64, device=tg.device)
                rows.append(torch.cat([torch.full((k.shape[0], 1), hdr, dtype=torch.int64, device=tg.device),
                                       own[:, None], k, pad], dim=1))
            key = torch.cat(rows)
            return [key[:, j] for j in range(width)]
        if tg.n_fac:
            fid, _, _ = rank_first(torch.cat(hs), exact_rows, torch.cat(hs2))
            fcol = [fid[int(a):int(b)] for a, b in zip(tg.foff[:-1], tg.foff[1:])]
        del keys
        # ---- variables: (own class, multiset of incident factor classes) through two 64-bit sums
        if tg.n_fac:
            c = fid[tg.inc_fac]
            H1 = tg.segment_sum(mix(c, 0x243F6A8885A308D3))
            H2 = tg.segment_sum(mix(c, 0x13198A2E03707344))
        else:
            H1 = H2 = torch.zeros(tg.n_vars, dtype=torch.int64, device=tg.device)
        vh = mix(vcol, 0xA4093822299F31D0) + H1 + mix(H2, 0x082EFA98EC4E6C89)
        vh2 = mix(mix(vcol, 0x9216D5D98979FB1B) ^ H2, 0xD1310BA698DFB5AC) + mix(H1, 0x2FFD72DBD01ADFB7)
        vcol, n_classes, _ = rank_first(vh, [vcol, H1, H2], vh2)
    return vcol, fcol, sweeps


def initial_colouring(tg: TorchGraph, cont, split_cont_evidence=True):
    """``lifting.initial_colouring`` (``init_cluster``, ``CompressedGraphWithObs.py:187-234``): one class per
    (domain, hidden | evidence value), ids in lexicographic order of that triple as the host pass numbers
    them; ``split_cont_evidence=False`` puts all continuous observations of a domain in one class."""
    val = torch.where(tg.hidden, torch.zeros_like(tg.var_value), tg.var_value)
    if not split_cont_evidence:
        val = torch.where(cont[tg.var_dom], torch.zeros_like(val), val)
    uniq, val_id = torch.unique(val, return_inverse=True)
    key = (tg.var_dom * 2 + tg.hidden.to(torch.int64)) * max(int(uniq.numel()), 1) + val_id
    return torch.unique(key, return_inverse=True)[1]


# ---- class statistics, parameter slots, inheritance -----------------------------------------------------

def class_stats(tg: TorchGraph, vcol, ev_value=None, base=None):
    """``lifting.class_stats`` on resident arrays: per variable class its size, representative (smallest
    member), hidden flag, mean evidence value (k-means centroid where ``ev_value = (has, value)`` names one),
    population variance around the members' mean, representative degree and domain."""
    if base is not None:
        st = dict(base)
        if ev_value is not None:
            n = st["n"]
            st["mean"] = torch.where(ev_value[0][:n], ev_value[1][:n], st["mean"])
        return st
    ncls = int(vcol.max()) + 1 if tg.n_vars else 0
    sizes = torch.bincount(vcol, minlength=ncls)
    reps = first_index(vcol, ncls)
    val0 = torch.where(tg.hidden, torch.zeros_like(tg.var_value), tg.var_value)
    fs = sizes.to(torch.float64)
    mean = torch.zeros(ncls, dtype=torch.float64, device=tg.device).index_add_(0, vcol, val0) / fs
    dev = torch.where(tg.hidden, torch.zeros_like(val0), tg.var_value - mean[vcol])
    variance = torch.zeros(ncls, dtype=torch.float64, device=tg.device).index_add_(0, vcol, dev * dev) / fs
    if ev_value is not None:
        mean = torch.where(ev_value[0][:ncls], ev_value[1][:ncls], mean)
    return dict(n=ncls, size=sizes, rep=reps, hidden=tg.hidden[reps], mean=mean, variance=variance,
                degree=tg.degree[reps], dom=tg.var_dom[reps])


def domain_tables(domains, device):
    from . import lowering
    cont = torch.tensor([bool(d.continuous) for d in domains], dtype=torch.bool, device=device)
    ddim = torch.tensor([2 if d.continuous else len(d.values) for d in domains], dtype=torch.int64, device=device)
    if bool((~cont & (ddim > lowering.MAX_DSTATES)).any()):
        raise ValueError(f"discrete variable with more than {lowering.MAX_DSTATES} states")
    return cont, ddim


def slot_layout(K, dom, cont, ddim):
    """``lifting.slot_layout``: ``(kind, dim, off, n_param)`` of hidden classes with domains ``dom``."""
    dim = ddim[dom]
    n = K * dim
    slot = torch.where(n <= 2, torch.full_like(n, 2), (n + 3) // 4 * 4)
    off = torch.cumsum(slot, 0) - slot
    return (~cont[dom]).to(torch.uint8), dim, off, int(slot.sum()) if slot.numel() else 0


def slot_elements(off, n):
    """Flat element indices of blocks of ``n[i]`` elements starting at ``off[i]``, concatenated."""
    start = torch.cumsum(n, 0) - n
    total = int(n.sum()) if n.numel() else 0
    return torch.repeat_interleave(off - start, n) + torch.arange(total, device=off.device, dtype=torch.int64)


def layout(tg: TorchGraph, vcol, K, cont, ddim):
    """``C2FArrayVI._layout``: parameter slots of the hidden classes of a partition, in class order."""
    st = class_stats(tg, vcol)
    cls = torch.nonzero(st["hidden"]).reshape(-1)
    kind, dim, off, n_param = slot_layout(K, st["dom"][cls], cont, ddim)
    slot_of = torch.full((st["n"],), -1, dtype=torch.int64, device=tg.device)
    slot_of[cls] = torch.arange(cls.numel(), device=tg.device, dtype=torch.int64)
    return dict(cls=cls, rep=st["rep"][cls], kind=kind, dim=dim, off=off, n_param=max(n_param, 2), slot_of=slot_of), st


def refine_bookkeeping(old, new, may_split, ev_has, ev_val):
    """``C2FArrayVI._refine`` after the colour passing: the evidence book-keeping carried over to the
    ids of the refinement ``new`` of ``old`` (every new class has one parent)."""
    n_old, n_new = int(old.max()) + 1, int(new.max()) + 1
    parent = old[first_index(new, n_new)]                 # [n_new] parent of every new class
    kids_of = torch.bincount(parent, minlength=n_old)
    may = may_split[parent]                               # every piece of a class k-means may split may be split
    kid = torch.arange(n_new, device=old.device, dtype=torch.int64)
    first_kid = torch.full((n_old,), n_new, dtype=torch.int64, device=old.device).scatter_reduce_(0, parent, kid, "amin")
    keep = torch.nonzero(ev_has[:n_old] & (kids_of == 1)).reshape(-1)
    has = torch.zeros(n_new, dtype=torch.bool, device=old.device)
    val = torch.zeros(n_new, dtype=torch.float64, device=old.device)
    has[first_kid[keep]] = True
    val[first_kid[keep]] = ev_val[keep]
    return may, has, val


def inherit(lay_old, lay, old, K, arrays):
    """``C2FArrayVI._inherit``: the pieces of a hidden class start from its parameters and Adam moments."""
    parent = lay_old["slot_of"][old[lay["rep"]]]
    n = K * lay["dim"]
    dst = slot_elements(lay["off"], n)
    src = slot_elements(lay_old["off"][parent], n)
    out = []
    for a in arrays:
        fresh = torch.zeros(lay["n_param"], dtype=a.dtype, device=a.device)
        fresh[dst] = a[src]
        out.append(fresh)
    return out


# ---- the compressed model's record columns from the partition --------------------------------------------

def _dense_rows(cols):
    """Dense ids (any order) of the rows of ``cols`` (int64 tensors), exact."""
    ids = torch.zeros_like(cols[0])
    for c in cols:
        _, ci = torch.unique(c, return_inverse=True)
        _, ids = torch.unique(ids * (int(ci.max()) + 1) + ci, return_inverse=True)
    return ids


def lower_partition(tg: TorchGraph, ga, var_colour, factor_colours, K, T, *, ev_value=None, gaussian_obs=False,
                    min_obs_var=0.0, stats=None):
    """``lifting.lower_partition`` with every pass over the ground graph and over the classes on ``tg.device``
    (same groups, same records in the same order, same coefficient table); the finished columns are
    handed to the engine as host arrays."""
    from . import lowering
    from .lifting import _domain_value
    if not 1 <= K <= lowering.MAX_K:
        raise ValueError(f"num_mixtures must be in 1..{lowering.MAX_K}, got {K}")
    if not 1 <= T <= lowering.MAX_T:
        raise ValueError(f"num_quadrature_points must be in 1..{lowering.MAX_T}, got {T}")
    HD, HC, EG, EC, ED = lowering.HD, lowering.HC, lowering.EG, lowering.EC, lowering.ED
    dev = tg.device
    i64, f64 = torch.int64, torch.float64
    st = class_stats(tg, var_colour, ev_value, base=stats)
    ncls, mean, variance = st["n"], st["mean"], st["variance"]
    cont_dom, ddim = domain_tables(ga.domains, dev)
    cls_cont = cont_dom[st["dom"]]
    cls_hidden = st["hidden"]
    cls_gauss = (~cls_hidden & (variance > min_obs_var)) if gaussian_obs else torch.zeros(ncls, dtype=torch.bool, device=dev)
    role_of = torch.where(cls_hidden, torch.where(cls_cont, HC, HD),
                          torch.where(cls_gauss, EG, torch.where(cls_cont, EC, ED))).to(i64)

    slot_class = torch.nonzero(cls_hidden).reshape(-1)
    kind, dim, off, n_param = slot_layout(K, st["dom"][slot_class], cont_dom, ddim)
    class_off = torch.full((ncls,), -1, dtype=i64, device=dev)
    class_off[slot_class] = off

    n_fc = max((int(c.max()) + 1 for c in factor_colours if c.numel()), default=0)
    size_f = torch.zeros(n_fc, dtype=i64, device=dev)
    for c in factor_colours:
        size_f += torch.bincount(c, minlength=n_fc)
    # incidences of each factor class on the class representatives (rv.count[f]): the representatives' segments
    # of the incidence list (prepared once, sorted by variable) instead of a scan over every argument column
    rep = st["rep"]
    deg = tg.degree[rep]
    total = int(deg.sum()) if rep.numel() else 0
    if total:
        fid_all = torch.cat(list(factor_colours))
        seg = torch.repeat_interleave(torch.arange(rep.numel(), device=dev, dtype=i64), deg)
        within = torch.arange(total, device=dev, dtype=i64) - (torch.cumsum(deg, 0) - deg)[seg]
        f = tg.inc_fac[tg.inc_ptr[rep][seg] + within]
        cnt_key, cnt_val = torch.unique(seg * n_fc + fid_all[f], return_counts=True)       # (class c has representative index c)
    else:
        cnt_key, cnt_val = torch.zeros(0, dtype=i64, device=dev), torch.zeros(0, dtype=i64, device=dev)

    def count_of(vcls, fcls):
        if cnt_key.numel() == 0:
            return torch.zeros(vcls.shape, dtype=f64, device=dev)
        k = vcls * n_fc + fcls
        at = torch.clamp(torch.searchsorted(cnt_key, k), max=cnt_key.numel() - 1)
        return torch.where(cnt_key[at] == k, cnt_val[at], torch.zeros_like(k)).to(f64)

    table = lowering.PotentialTable()
    chunks = {}
    unary_w = torch.zeros(ncls, dtype=f64, device=dev)
    unary_g = torch.zeros(ncls, dtype=f64, device=dev)
    seen = torch.zeros(n_fc, dtype=torch.bool, device=dev)
    uid_next = 0
    n_dom = len(ga.domains) + 1
    for b, args_t, fcol in zip(ga.blocks, tg.args, factor_colours):
        if b.n == 0:
            continue
        first_of = first_index(fcol, n_fc)                                   # (unused class: b.n)
        ids = torch.nonzero((first_of < b.n) & ~seen).reshape(-1)           # classes first met in this block, ascending id
        first = first_of[ids]
        seen[ids] = True
        if ids.numel() == 0:
            continue
        uid = uid_next + torch.arange(ids.numel(), device=dev, dtype=i64)
        uid_next += int(ids.numel())
        nbcls = var_colour[args_t[first]]                                    # [m, arity]
        roles = role_of[nbcls]
        arity = b.arity
        doms = torch.where((roles == HD) | (roles == ED), st["dom"][nbcls], torch.zeros_like(nbcls))
        vals = torch.where(roles == ED, mean[nbcls], torch.zeros(nbcls.shape, dtype=f64, device=dev))
        code = torch.zeros(ids.numel(), dtype=i64, device=dev)
        for j in range(arity):
            code = (code * 8 + roles[:, j]) * n_dom + doms[:, j]
        cols = [code]
        if bool((roles == ED).any()):
            cols += [torch.unique(vals[:, j], return_inverse=True)[1].reshape(-1) for j in range(arity)]
        inv = _dense_rows(cols)
        n_combo = int(inv.max()) + 1
        where = first_index(inv, n_combo)
        combo = torch.cat([roles.to(f64), doms.to(f64), vals], dim=1)[where].cpu().numpy()
        for ci in torch.argsort(where, stable=True).tolist():
            sel = torch.nonzero(inv == ci).reshape(-1)
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
            w_f = size_f[ids[sel]].to(f64)
            gam = torch.zeros((len(hid), sel.numel()), dtype=f64, device=dev)
            for j, i in enumerate(hid):
                is_first = torch.ones(sel.numel(), dtype=torch.bool, device=dev)
                for e in range(i):
                    is_first &= nb[:, e] != nb[:, i]
                gam[j] = torch.where(is_first, count_of(nb[:, i], ids[sel]), torch.zeros(sel.numel(), dtype=f64, device=dev))
            pure = (nd + nc + ng == 0) or (nd + nc == 1 and ng == 0)
            if pure and hid:
                unary_w.index_add_(0, nb[:, hid[0]], w_f)
                unary_g.index_add_(0, nb[:, hid[0]], gam[0])
            pot = table.block(b.potential, r, args)
            pkind = lowering.potential_kind(b.potential, nc + ng + ne)
            m = sel.numel()
            chunks.setdefault((nd, nc, ng, ne, dims, False, pure, pkind), []).append(dict(
                uid=uid[sel], pot=torch.full((m,), pot, dtype=torch.int32, device=dev),
                poff=class_off[nb[:, hid]].T.reshape(len(hid), m),
                egval=mean[nb[:, pos[EG]]].T.reshape(ng, m),
                egvar=variance[nb[:, pos[EG]]].T.reshape(ng, m),
                ecval=mean[nb[:, pos[EC]]].T.reshape(ne, m),
                wf=w_f, gam=gam, nscale=torch.zeros(m, dtype=f64, device=dev)))

    # ---- node-entropy records, one per hidden class and per Gaussian-evidence class
    scale = (st["degree"] - 1).to(f64)
    c_v = st["size"].to(f64)

    def empty(rows, m):
        return torch.zeros((rows, m), dtype=f64, device=dev)
    kd = torch.unique(kind.to(i64) * 1024 + dim).tolist() if kind.numel() else []
    for code in sorted(kd):
        is_disc, d = int(code) // 1024, int(code) % 1024
        pick = slot_class[(kind == is_disc) & (dim == d)]
        key = (1, 0, 0, 0, (d,), True, False, 0) if is_disc else (0, 1, 0, 0, (), True, False, 0)
        m = pick.numel()
        chunks.setdefault(key, []).append(dict(
            uid=pick, pot=torch.zeros(m, dtype=torch.int32, device=dev), poff=class_off[pick][None, :],
            egval=empty(0, m), egvar=empty(0, m), ecval=empty(0, m),
            wf=c_v[pick] * scale[pick] - unary_w[pick], gam=torch.ones((1, m), dtype=f64, device=dev),
            nscale=scale[pick] - unary_g[pick]))
    pick = torch.nonzero(cls_gauss).reshape(-1)
    if pick.numel():
        m = pick.numel()
        chunks.setdefault((0, 0, 1, 0, (), True, False, 0), []).append(dict(
            uid=pick, pot=torch.zeros(m, dtype=torch.int32, device=dev), poff=torch.zeros((0, m), dtype=i64, device=dev),
            egval=mean[pick][None, :], egvar=variance[pick][None, :], ecval=empty(0, m),
            wf=c_v[pick] * scale[pick], gam=empty(0, m), nscale=scale[pick]))

    groups = []
    host = lambda t: t.cpu().numpy()
    for (nd, nc, ng, ne, dims, node, pure, pkind), parts in chunks.items():
        cat = {k: torch.cat([p[k] for p in parts], dim=-1) for k in parts[0]}
        order = torch.argsort(cat["uid"], stable=True)
        wf, gam = cat["wf"][order], cat["gam"][:, order]
        weighted = bool(node or bool((wf != 1.0).any()) or bool((gam != 1.0).any()))
        groups.append(lowering.RecordGroup(
            nd, nc, ng, ne, dims, node, host(cat["pot"][order].to(torch.int32)),
            np.ascontiguousarray(host(cat["poff"][:, order].to(torch.int32))),
            np.ascontiguousarray(host(cat["egval"][:, order])), np.ascontiguousarray(host(cat["egvar"][:, order])),
            np.ascontiguousarray(host(cat["ecval"][:, order])), host(wf), np.ascontiguousarray(host(gam)),
            host(cat["nscale"][order]), weighted, pure, (), pkind))
    groups.sort(key=lambda g: (not g.node, not g.pure, g.nd + g.nc + g.ng, g.nd, g.nc, g.ng, g.ne, g.dims, g.kind))
    return lowering.LoweredModel(K, T, max(n_param, 2), host(kind), host(dim).astype(np.int32), host(off).astype(np.int32),
                                 table.array(), groups, [], {}, host(slot_class))


# ---- evidence k-means (CompressedGraphWithObs.py:78-130, 236-247) ----------------------------------------

def _segment_var(values, seg, n_seg, counts):
    fc = counts.to(torch.float64)
    mean = torch.zeros(n_seg, dtype=torch.float64, device=values.device).index_add_(0, seg, values) / fc
    dev = values - mean[seg]
    return torch.zeros(n_seg, dtype=torch.float64, device=values.device).index_add_(0, seg, dev * dev) / fc


def split_evidence(tg: TorchGraph, vcol, may_split, ev_has, ev_val, epsilon, k, iterations):
    """``C2FArrayVI._split_evidence`` (two centroids) on resident arrays: every evidence class k-means may
    split and whose spread exceeds ``epsilon`` is cut in two by 1-D 2-means (centroids start at the first two
    distinct values in member order, ``iterations`` Lloyd sweeps), repeated until nothing changes.  All
    classes of a pass are processed at once; new ids are handed out in ascending order of the class they
    were cut from, as the host pass does.  Returns ``(vcol, may_split, ev_has, ev_val)`` (new tensors)."""
    if k != 2:
        raise NotImplementedError("the device pass implements the reference's default of two centroids (k_mean_k = 2)")
    dev, f64, i64 = tg.device, torch.float64, torch.int64
    vcol = vcol.clone()
    n = int(vcol.max()) + 1
    room = n + int(may_split[vcol].sum())
    may = torch.zeros(room, dtype=torch.bool, device=dev)
    has = torch.zeros(room, dtype=torch.bool, device=dev)
    val = torch.zeros(room, dtype=f64, device=dev)
    may[:n], has[:n], val[:n] = may_split, ev_has, ev_val
    next_id = n
    while True:
        if not bool(may[:next_id].any()):
            break
        cand = torch.nonzero(may[vcol]).reshape(-1)
        cand = cand[torch.argsort(vcol[cand], stable=True)]            # grouped by class, members in index order
        ccol = vcol[cand]
        cids, ccnt = torch.unique_consecutive(ccol, return_counts=True)
        n_seg = cids.numel()
        seg = torch.repeat_interleave(torch.arange(n_seg, device=dev, dtype=i64), ccnt)
        cstart = torch.cumsum(ccnt, 0) - ccnt
        cval = tg.var_value[cand]
        var = _segment_var(cval, seg, n_seg, ccnt)
        vmax = torch.full((n_seg,), -float("inf"), dtype=f64, device=dev).scatter_reduce_(0, seg, cval, "amax")
        vmin = torch.full((n_seg,), float("inf"), dtype=f64, device=dev).scatter_reduce_(0, seg, cval, "amin")
        wide = (torch.sqrt(var) > epsilon) & (vmax > vmin)
        if not bool(wide.any()):
            break
        # the members of the classes that are cut, regrouped
        pick = wide[seg]
        cand, cval = cand[pick], cval[pick]
        wid = torch.nonzero(wide).reshape(-1)
        cids, ccnt = cids[wid], ccnt[wid]
        n_seg = cids.numel()
        seg = torch.repeat_interleave(torch.arange(n_seg, device=dev, dtype=i64), ccnt)
        cstart = torch.cumsum(ccnt, 0) - ccnt
        pos = torch.arange(cand.numel(), device=dev, dtype=i64)
        c0 = cval[cstart]                                              # first value, first different value
        other = torch.where(cval != c0[seg], pos, torch.full_like(pos, cand.numel()))
        c1 = cval[torch.full((n_seg,), cand.numel(), dtype=i64, device=dev).scatter_reduce_(0, seg, other, "amin")]
        for _ in range(iterations):
            own1 = (torch.abs(cval - c1[seg]) < torch.abs(cval - c0[seg]))      # ties go to the first centroid (argmin)
            w1 = own1.to(f64)
            m1 = torch.zeros(n_seg, dtype=f64, device=dev).index_add_(0, seg, w1)
            t1 = torch.zeros(n_seg, dtype=f64, device=dev).index_add_(0, seg, w1 * cval)
            tall = torch.zeros(n_seg, dtype=f64, device=dev).index_add_(0, seg, cval)
            c0, c1 = (tall - t1) / (ccnt.to(f64) - m1), t1 / m1
        own1 = (torch.abs(cval - c1[seg]) < torch.abs(cval - c0[seg]))
        n1 = torch.zeros(n_seg, dtype=i64, device=dev).index_add_(0, seg, own1.to(i64))
        second = n1 > 0
        new_id = next_id + torch.cumsum(second.to(i64), 0) - 1
        has[cids], val[cids] = True, c0
        if bool(second.any()):
            moved = own1                                               # (own1 implies second[seg])
            vcol[cand[moved]] = new_id[seg[moved]]
            nid = new_id[second]
            has[nid], val[nid], may[nid] = True, c1[second], False
            # a piece stays a candidate when its variance (not its deviation: the reference's test) exceeds epsilon
            two = seg * 2 + own1.to(i64)
            cnt2 = torch.bincount(two, minlength=2 * n_seg)
            var2 = _segment_var(cval, two, 2 * n_seg, torch.clamp(cnt2, min=1)).reshape(n_seg, 2)
            big = var2 > epsilon
            keep0 = second & big[:, 0]
            may[cids[keep0]] = True
            may[nid] = big[second, 1]
            next_id += int(second.sum())
        else:
            break
    return vcol, may[:next_id].clone(), has[:next_id].clone(), val[:next_id].clone()
