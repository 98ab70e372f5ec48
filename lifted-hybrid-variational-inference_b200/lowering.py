"""Lowering: object graphs -> structure-of-arrays factor tables for the CUDA kernels.

The reference walks Python objects and calls ``potential.get`` at every quadrature point
(``VarInference.py:40-55,74-88``).  Here the graph is flattened once per ``run`` (or once
per C2F refinement round) into

* a flat parameter vector: one slot per *hidden* variable -- continuous ``[K][mu, var]``
  interleaved, discrete ``[K][D]`` row-major -- plus its slot offset table;
* a coefficient table ``ptab``: every potential becomes, per discrete configuration,
  ``log psi = c + b'x + x'Ax`` over its continuous arguments (SURVEY section 8 a-P), or a
  single pre-computed ``log(psi + 1e-100)`` when it has no continuous argument;
* *record groups*: factors with the same canonical signature
  ``(#hidden-discrete, #hidden-continuous, #Gaussian-evidence, #point-evidence-continuous)``
  stored as columns (``pot`` offset, parameter offsets, evidence values, lifted weights).
  Arguments are permuted into that canonical order (the coefficient block is permuted to
  match); discrete point evidence is folded into the ``pot`` offset, continuous point
  evidence travels as a per-record value column.

The variables' own ``(N-1) E[log b]`` terms (``VarInference.py:60-72,96-106,138-139``) are
emitted as *node groups*: unary records with ``F = log b`` and two per-record scales (one for
the energy / G_w, one for the parameter gradients).

*Unary split.*  A factor with a single integrated argument v has
``F = log psi(x) - log b_v(x)``; its ``-log b_v`` part depends on the variable only, so by
linearity of the expectation it is moved into v's node record (its scales become
``C_v (N_v-1) - sum W_f`` and ``(N_v-1) - sum gamma_f``) and the factor is stored as a *pure*
record that evaluates ``log psi`` alone -- no density, exp or log per record.  The sums the
reference computes are unchanged; only their grouping differs.

Nothing here touches the GPU; ``engine.py`` uploads the result.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field

import numpy as np

# argument roles, in canonical record order
HD, HC, EG, EC, ED = 0, 1, 2, 3, 4

MAX_ARITY = 6      # hidden + Gaussian-evidence axes walked by the kernels
MAX_DSTATES = 16   # states of one hidden discrete argument
MAX_K = 8
MAX_T = 32
LOG_FLOOR_EPS = 1e-100   # the reference's "+1e-100" inside both logs


def slot_size(n_elems: int) -> int:
    """Padded size of one variable's parameter slot (keeps 16-byte vector alignment)."""
    return 2 if n_elems <= 2 else (n_elems + 3) // 4 * 4


def ncoef_for(nct: int) -> int:
    """Coefficients per discrete configuration: c, b[nct], upper-triangular A."""
    return 1 if nct == 0 else (nct + 1) * (nct + 2) // 2


@dataclass
class RecordGroup:
    """One signature's worth of factor records, column-major."""
    nd: int
    nc: int
    ng: int
    ne: int
    dims: tuple            # states of each hidden discrete argument
    node: bool             # node-entropy pseudo factors: F = log b, energy x wf, gradients x nscale
    pot: np.ndarray        # int32 [n]            offset of the coefficient block in ptab
    poff: np.ndarray       # int32 [nd+nc, n]     parameter-slot offsets
    egval: np.ndarray      # f64   [ng, n]        Gaussian-evidence mean
    egvar: np.ndarray      # f64   [ng, n]        Gaussian-evidence variance
    ecval: np.ndarray      # f64   [ne, n]        point-evidence values (continuous args)
    wf: np.ndarray         # f64   [n]            W_f: energy / g_w weight (node: energy scale)
    gam: np.ndarray        # f64   [nd+nc, n]     gamma: parameter-gradient weight per arg
    nscale: np.ndarray     # f64   [n]            node groups: parameter-gradient scale
    weighted: bool         # False -> wf == gam == 1 everywhere (columns not shipped)
    pure: bool = False     # True -> F = log psi only (the -log b part lives in the node records)
    dvals: tuple = ()      # domain values of each hidden discrete argument (tuple of tuples of float; NaN for a
                           # non-numeric value): the nodes the reference's H2 path puts on the OTHER arguments
    kind: int = 0          # POT_QUADRATIC | POT_HARD | POT_IMAGE_EDGE: how a coefficient block becomes log psi

    @property
    def n(self) -> int:
        return int(self.wf.shape[0])

    @property
    def nh(self) -> int:
        return self.nd + self.nc

    @property
    def nct(self) -> int:
        return self.nc + self.ng + self.ne

    @property
    def ncfg(self) -> int:
        return int(np.prod(self.dims)) if self.dims else 1

    @property
    def signature(self):
        return (self.node, self.pure, self.nd, self.nc, self.ng, self.ne, tuple(self.dims), self.weighted)

    def take(self, sel) -> "RecordGroup":
        """Sub-group (used to shard records across ranks)."""
        return RecordGroup(self.nd, self.nc, self.ng, self.ne, self.dims, self.node,
                           self.pot[sel], self.poff[:, sel], self.egval[:, sel],
                           self.egvar[:, sel], self.ecval[:, sel], self.wf[sel],
                           self.gam[:, sel], self.nscale[sel], self.weighted, self.pure, self.dvals, self.kind)


@dataclass
class LoweredModel:
    K: int
    T: int
    n_param: int                      # elements in the flat parameter vector
    var_kind: np.ndarray              # uint8 [V]  0 = continuous, 1 = discrete
    var_dim: np.ndarray               # int32 [V]  2 for continuous, D for discrete
    var_off: np.ndarray               # int32 [V]  slot offset
    ptab: np.ndarray                  # f64 coefficient table
    groups: list
    handles: list = field(default_factory=list)      # hidden variable objects, slot order
    index: dict = field(default_factory=dict)        # handle -> position in `handles`
    slot_class: np.ndarray = None     # int64 [V]  variable class behind each slot (array-native lowering)

    @property
    def n_vars(self) -> int:
        return int(self.var_off.shape[0])

    @property
    def n_records(self) -> int:
        return sum(g.n for g in self.groups if not g.node)

    @property
    def n_node_records(self) -> int:
        return sum(g.n for g in self.groups if g.node)

    def shard(self, rank: int, world: int) -> "LoweredModel":
        """Contiguous 1/world slice of every group's records; parameters stay replicated
        (SURVEY section 8 e)."""
        if world == 1:
            return self
        parts = []
        for g in self.groups:
            lo = g.n * rank // world
            hi = g.n * (rank + 1) // world
            if hi > lo:
                parts.append(g.take(slice(lo, hi)))
        return LoweredModel(self.K, self.T, self.n_param, self.var_kind, self.var_dim,
                            self.var_off, self.ptab, parts, self.handles, self.index, self.slot_class)


# ----------------------------------------------------------------------------------------
# potentials -> coefficient blocks
# ----------------------------------------------------------------------------------------

# How a coefficient block of ``ptab`` becomes log psi at a grid point (include/lhvi.h, LHVI_POT_*).
POT_QUADRATIC, POT_HARD, POT_IMAGE_EDGE = 0, 1, 2


def potential_kind(potential, n_continuous):
    """``POT_*`` of a potential whose factor has ``n_continuous`` continuous arguments (hidden or
    observed).  Everything exp-quadratic -- and every potential over discrete arguments only, which
    is tabulated through ``get`` -- is ``POT_QUADRATIC``.  Two potentials of the reference are not
    log-quadratic in their continuous arguments and are evaluated point by point on the device:

    * ``MLNHardPotential`` (``MLNPotential.py:43-49``: 1 where ``formula(x) > 0``, else 0) -- recognised
      by a ``formula`` without a weight ``w``; the formula itself must be (at most) quadratic in the
      continuous arguments, as every formula of the reference is (a-P);
    * ``ImageEdgePotential`` (``Potential.py:411-424``) -- recognised by its three coefficients."""
    if n_continuous == 0:
        return POT_QUADRATIC
    if getattr(potential, "formula", None) is not None and not hasattr(potential, "w"):
        return POT_HARD
    if all(hasattr(potential, a) for a in ("distant_cof", "scaling_cof", "max_threshold")):
        return POT_IMAGE_EDGE
    return POT_QUADRATIC


def _log_psi_fn(potential):
    """Callable x -> log psi(x) for exp-type potentials (exact for MLN formulas)."""
    formula = getattr(potential, "formula", None)
    if formula is not None and hasattr(potential, "w"):
        w = potential.w
        return lambda x: float(formula(x) * w)
    if formula is not None:               # POT_HARD: the block holds the formula, the kernels test its sign
        return lambda x: float(formula(x))

    def via_get(x):
        v = float(np.asarray(potential.get(x)).reshape(-1)[0])
        if not v > 0:
            raise NotImplementedError(
                f"{type(potential).__name__}: psi <= 0 at a continuous probe point; "
                "not an exp-quadratic potential")
        return math.log(v)
    return via_get


def _fit_quadratic(logpsi, template, cpos):
    """Exact second-order finite differences of ``logpsi`` in the continuous positions
    ``cpos`` of the argument ``template`` (discrete positions already filled in).
    Returns [c, b_0.., A_00, A_01, .., A_(n-1)(n-1)] with A upper-triangular (cross terms
    carry the full coefficient), then verifies the fit on random probes."""
    nct = len(cpos)

    def at(vec):
        x = list(template)
        for p, v in zip(cpos, vec):
            x[p] = float(v)
        return logpsi(tuple(x))

    zero = np.zeros(nct)
    f0 = at(zero)
    unit = np.eye(nct)
    fp = [at(unit[i]) for i in range(nct)]
    fm = [at(-unit[i]) for i in range(nct)]
    coef = [f0]
    coef += [(fp[i] - fm[i]) * 0.5 for i in range(nct)]
    for i in range(nct):
        for j in range(i, nct):
            if i == j:
                coef.append((fp[i] + fm[i]) * 0.5 - f0)
            else:
                coef.append(at(unit[i] + unit[j]) - fp[i] - fp[j] + f0)
    coef = np.array(coef, dtype=float)

    rng = np.random.default_rng(12345)
    for _ in range(6):
        probe = rng.uniform(-3.0, 3.0, size=nct)
        want = at(probe)
        got = eval_quadratic(coef, probe)
        if abs(want - got) > 1e-8 * max(1.0, abs(want)):
            raise NotImplementedError(
                "potential is not (log-)quadratic in its continuous arguments: "
                f"probe {probe} gives {want}, quadratic fit gives {got}")
    return coef


def eval_quadratic(coef, x):
    """c + b'x + sum_{i<=j} A_ij x_i x_j for the packed layout used in ``ptab``."""
    nct = len(x)
    val = coef[0]
    for i in range(nct):
        val += coef[1 + i] * x[i]
    p = 1 + nct
    for i in range(nct):
        for j in range(i, nct):
            val += coef[p] * x[i] * x[j]
            p += 1
    return val


def _pack_quadratic(A, b, c, order):
    """Packed coefficients of x'Ax + b'x + c after permuting arguments by ``order``
    (``order[i]`` = original position of canonical continuous argument i)."""
    A = np.asarray(A, dtype=float)
    b = np.asarray(b, dtype=float).reshape(-1)
    n = len(order)
    out = [float(c)] + [float(b[o]) for o in order]
    for i in range(n):
        for j in range(i, n):
            oi, oj = order[i], order[j]
            out.append(float(A[oi, oi]) if i == j else float(A[oi, oj] + A[oj, oi]))
    return np.array(out, dtype=float)


class PotentialTable:
    """Builds and caches coefficient blocks keyed by (potential, roles, evidence config)."""

    def __init__(self):
        self.chunks = []
        self.size = 0
        self.cache = {}

    def block(self, potential, roles, args):
        """Offset of the block for a factor whose argument ``i`` has role ``roles[i]``;
        ``args[i]`` is the tuple of domain values for a hidden-discrete argument and the
        evidence value for a discrete-evidence argument (ignored otherwise)."""
        ed_vals = tuple(args[i] for i, r in enumerate(roles) if r == ED)
        hd_vals = tuple(tuple(args[i]) for i, r in enumerate(roles) if r == HD)
        key = (potential, tuple(roles), ed_vals, hd_vals)
        try:
            hit = self.cache.get(key)
        except TypeError:       # unhashable potential: fall back to identity
            key = (id(potential), tuple(roles), ed_vals, hd_vals)
            hit = self.cache.get(key)
        if hit is not None:
            return hit
        data = self._build(potential, roles, args)
        off = self.size
        self.chunks.append(data)
        self.size += data.size
        self.cache[key] = off
        return off

    @staticmethod
    def _build(potential, roles, args):
        n = len(roles)
        hd_pos = [i for i in range(n) if roles[i] == HD]
        c_pos = ([i for i in range(n) if roles[i] == HC] + [i for i in range(n) if roles[i] == EG]
                 + [i for i in range(n) if roles[i] == EC])
        nct = len(c_pos)
        template = [None] * n
        for i in range(n):
            if roles[i] == ED:
                template[i] = args[i]
        out = []
        quad = None
        if potential_kind(potential, nct) == POT_IMAGE_EDGE:
            if nct != 2 or n != 2:
                raise NotImplementedError("ImageEdgePotential takes two continuous arguments")
            v = float(getattr(potential, "v", math.exp(-potential.max_threshold / potential.scaling_cof)))
            return np.array([float(potential.distant_cof), float(potential.scaling_cof),
                             float(potential.max_threshold), v, 0.0, 0.0])
        if nct and not hd_pos and all(r != ED for r in roles) and hasattr(potential, "get_quadratic_params"):
            A, b, c = potential.get_quadratic_params()
            A = np.asarray(A, dtype=float)
            if A.shape == (n, n):
                quad = _pack_quadratic(A, b, c, c_pos)
        logpsi = None
        for cfg in itertools.product(*[args[p] for p in hd_pos]):
            for p, v in zip(hd_pos, cfg):
                template[p] = v
            if nct == 0:
                psi = float(np.asarray(potential.get(tuple(template))).reshape(-1)[0])
                out.append(np.array([math.log(psi + LOG_FLOOR_EPS)]))
                continue
            if logpsi is None:
                logpsi = _log_psi_fn(potential)
            if quad is not None:
                # trust-but-verify the plugin's closed form against get()
                probe = np.linspace(-0.7, 0.9, nct)
                x = list(template)
                for p, v in zip(c_pos, probe):
                    x[p] = float(v)
                if abs(logpsi(tuple(x)) - eval_quadratic(quad, probe)) > 1e-9:
                    quad = None
            out.append(quad if quad is not None else _fit_quadratic(logpsi, template, c_pos))
        return np.concatenate(out)

    def array(self):
        return np.concatenate(self.chunks) if self.chunks else np.zeros(1)


def fold_unary(g: RecordGroup, ptab: np.ndarray):
    """Streaming form of a pure group with one hidden continuous argument: per record the
    quadratic ``log psi(x) = c0 + l0 x + a0 x^2`` in that argument with the point evidence
    substituted (coefficient layout of ``ptab``: c, b[nct], upper-triangular A row-major;
    canonical argument order [hidden | evidence...]).  Returns ``[3, n]`` float64 or ``None``
    when the group has another shape."""
    if g.node or not g.pure or g.nd != 0 or g.nc != 1 or g.ng != 0 or g.kind != POT_QUADRATIC:
        return None
    nct = 1 + g.ne
    ncoef = ncoef_for(nct)
    coef = ptab[g.pot.astype(np.int64)[:, None] + np.arange(ncoef)[None, :]]      # [n, ncoef]
    ev = g.ecval                                                                   # [ne, n]
    c0 = coef[:, 0].copy()
    l0 = coef[:, 1].copy()
    for e in range(g.ne):
        c0 += coef[:, 2 + e] * ev[e]
    p = 1 + nct
    a0 = None
    for i in range(nct):
        for j in range(i, nct):
            a = coef[:, p]
            p += 1
            if i == 0 and j == 0:
                a0 = a.copy()
            elif i == 0:
                l0 += a * ev[j - 1]
            else:
                c0 += a * ev[i - 1] * ev[j - 1]
    return np.stack([c0, l0, a0])


# ----------------------------------------------------------------------------------------
# graphs -> record groups
# ----------------------------------------------------------------------------------------

class _GroupBuilder:
    def __init__(self, nd, nc, ng, ne, dims, node, pure=False, dvals=(), kind=POT_QUADRATIC):
        self.sig = (nd, nc, ng, ne, tuple(dims), node)
        self.pure = pure
        self.dvals = tuple(dvals)
        self.kind = kind
        self.pot, self.poff, self.egval, self.egvar, self.ecval = [], [], [], [], []
        self.wf, self.gam, self.nscale = [], [], []

    def add(self, pot, poff, eg, ec, wf, gam, nscale=0.0):
        self.pot.append(pot)
        self.poff.append(poff)
        self.egval.append([v for v, _ in eg])
        self.egvar.append([s for _, s in eg])
        self.ecval.append(ec)
        self.wf.append(wf)
        self.gam.append(gam)
        self.nscale.append(nscale)

    def finish(self) -> RecordGroup:
        nd, nc, ng, ne, dims, node = self.sig
        n = len(self.pot)

        def cols(rows, width, dtype):
            a = np.asarray(rows, dtype=dtype).reshape(n, width)
            return np.ascontiguousarray(a.T)

        wf = np.asarray(self.wf, dtype=float)
        gam = cols(self.gam, nd + nc, float)
        weighted = bool(node or np.any(wf != 1.0) or np.any(gam != 1.0))
        return RecordGroup(nd, nc, ng, ne, dims, node,
                           np.asarray(self.pot, dtype=np.int32), cols(self.poff, nd + nc, np.int32),
                           cols(self.egval, ng, float), cols(self.egvar, ng, float),
                           cols(self.ecval, ne, float), wf, gam,
                           np.asarray(self.nscale, dtype=float), weighted, self.pure, self.dvals, self.kind)


def lower_graph(rvs, factors, K, T, *, factor_weight=None, arg_weight=None, var_weight=None,
                gaussian_evidence=None):
    """Flatten a (ground or compressed) graph.

    ``rvs`` / ``factors``: iterables of variable / factor handles exposing
    ``domain, value, N`` and ``potential, nb``.  Hooks supply the lifted weights:

    * ``factor_weight(f)``      -> W_f   (``len(f.factors)``; 1 for a ground graph)
    * ``arg_weight(f, i, rv)``  -> gamma (``rv.count[f]`` at the first position of ``rv`` in
      ``f.nb``, 0 at later ones; ground: multiplicity at the first position)
    * ``var_weight(rv)``        -> C_v   (``len(rv.rvs)``; 1 for a ground graph)
    * ``gaussian_evidence(rv)`` -> ``(value, variance)`` if this evidence variable is to be
      integrated as a fixed Gaussian (C2F, ``C2FVarInference.py:110-113``) else ``None``.
    """
    if not 1 <= K <= MAX_K:
        raise ValueError(f"num_mixtures must be in 1..{MAX_K}, got {K}")
    if not 1 <= T <= MAX_T:
        raise ValueError(f"num_quadrature_points must be in 1..{MAX_T}, got {T}")
    factor_weight = factor_weight or (lambda f: 1.0)
    var_weight = var_weight or (lambda rv: 1.0)
    gaussian_evidence = gaussian_evidence or (lambda rv: None)

    def default_arg_weight(f, i, rv):
        first = next(j for j, r in enumerate(f.nb) if r is rv)
        return float(sum(1 for r in f.nb if r is rv)) if i == first else 0.0
    arg_weight = arg_weight or default_arg_weight

    # ---- parameter slots for hidden variables
    handles, kind, dim, off = [], [], [], []
    cursor = 0
    for rv in rvs:
        if rv.value is not None:
            continue
        if rv.domain.continuous:
            k_, d_ = 0, 2
        else:
            k_, d_ = 1, len(rv.domain.values)
            if d_ > MAX_DSTATES:
                raise ValueError(f"discrete variable with {d_} states (max {MAX_DSTATES})")
        handles.append(rv)
        kind.append(k_)
        dim.append(d_)
        off.append(cursor)
        cursor += slot_size(K * d_)
    index = {rv: i for i, rv in enumerate(handles)}

    table = PotentialTable()
    builders = {}

    def builder(nd, nc, ng, ne, dims, node, pure=False, dvals=(), kind=POT_QUADRATIC):
        # (the domain values are part of the key: a group's hidden discrete arguments share them)
        key = (nd, nc, ng, ne, tuple(dims), node, pure, tuple(dvals), kind)
        if key not in builders:
            builders[key] = _GroupBuilder(nd, nc, ng, ne, dims, node, pure, dvals, kind)
        return builders[key]

    def numeric(values):
        out = []
        for v in values:
            try:
                out.append(float(v))
            except (TypeError, ValueError):
                out.append(float("nan"))
        return tuple(out)

    # ---- factor records
    unary_w = {}     # hidden variable -> sum of W_f over its unary (single integrated arg) factors
    unary_g = {}     # hidden variable -> sum of gamma over the same factors
    for f in factors:
        nb = list(f.nb)
        roles, args = [], []
        for rv in nb:
            if rv.value is None:
                if rv.domain.continuous:
                    roles.append(HC)
                    args.append(None)
                else:
                    roles.append(HD)
                    args.append(tuple(rv.domain.values))
            else:
                ge = gaussian_evidence(rv)
                if ge is not None:
                    roles.append(EG)
                    args.append(ge)
                elif rv.domain.continuous:
                    roles.append(EC)
                    args.append(float(rv.value))
                else:
                    roles.append(ED)
                    args.append(rv.value)
        pos = {r: [i for i, x in enumerate(roles) if x == r] for r in (HD, HC, EG, EC)}
        nd, nc, ng, ne = (len(pos[r]) for r in (HD, HC, EG, EC))
        if nd + nc + ng > MAX_ARITY:
            raise ValueError(f"factor with {nd + nc + ng} integrated arguments (max {MAX_ARITY})")
        dims = tuple(len(args[i]) for i in pos[HD])
        hidden = pos[HD] + pos[HC]
        w_f = float(factor_weight(f))
        gam = [float(arg_weight(f, i, nb[i])) for i in hidden]
        # unary split: a single integrated argument (or none) -> pure log-psi record
        pure = (nd + nc + ng == 0) or (nd + nc == 1 and ng == 0)
        if pure and hidden:
            v = nb[hidden[0]]
            unary_w[v] = unary_w.get(v, 0.0) + w_f
            unary_g[v] = unary_g.get(v, 0.0) + gam[0]
        dvals = tuple(numeric(args[i]) for i in pos[HD])
        builder(nd, nc, ng, ne, dims, False, pure, dvals, potential_kind(f.potential, nc + ng + ne)).add(
            table.block(f.potential, roles, args),
            [off[index[nb[i]]] for i in hidden],
            [args[i] for i in pos[EG]],
            [args[i] for i in pos[EC]],
            w_f, gam)

    # ---- node-entropy pseudo factors: F = log b_v, energy scale wf, gradient scale nscale
    for rv in rvs:
        scale = float(rv.N - 1)
        c_v = float(var_weight(rv))
        if rv.value is None:
            h = index[rv]
            s_e = c_v * scale - unary_w.get(rv, 0.0)
            s_g = scale - unary_g.get(rv, 0.0)
            if kind[h] == 0:
                b = builder(0, 1, 0, 0, (), True)
            else:
                b = builder(1, 0, 0, 0, (dim[h],), True, False, (numeric(rv.domain.values),))
            b.add(0, [off[h]], [], [], s_e, [1.0], s_g)
        else:
            ge = gaussian_evidence(rv)
            if ge is not None:
                builder(0, 0, 1, 0, (), True).add(0, [], [ge], [], c_v * scale, [], scale)
            # point evidence: b = sum_k w_k = 1, log term vanishes (VarInference.py:65-66)

    groups = [b.finish() for b in builders.values()]
    groups.sort(key=lambda g: (not g.node, not g.pure, g.nd + g.nc + g.ng, g.nd, g.nc, g.ng, g.ne, g.dims, g.kind))
    return LoweredModel(K, T, max(cursor, 2), np.asarray(kind, dtype=np.uint8),
                        np.asarray(dim, dtype=np.int32), np.asarray(off, dtype=np.int32),
                        table.array(), groups, handles, index)


def lower_ground(g, K, T):
    """Ground graph (``VarInference``): all weights are 1."""
    rvs = sorted(g.rvs, key=lambda rv: rv.id)
    factors = sorted(g.factors, key=lambda f: f.id)
    return lower_graph(rvs, factors, K, T)


def lower_compressed(cg, K, T, *, gaussian_obs=False, min_obs_var=0.0):
    """Compressed graph (``LiftedVarInference`` / ``C2FVarInference``).

    Weights follow ``LiftedVarInference.py:74,90,111-112,131-132``:
    energy and g_w scale with the class sizes, parameter gradients with
    ``rv.count[f]`` at the first occurrence of the class in ``f.nb``."""
    rvs = sorted(cg.rvs)
    factors = sorted(cg.factors)

    def arg_weight(f, i, rv):
        first = next(j for j, r in enumerate(f.nb) if r is rv)
        return float(rv.count[f]) if i == first else 0.0

    def gaussian_evidence(rv):
        if gaussian_obs and rv.value is not None and rv.variance > min_obs_var:
            return (float(rv.value), float(rv.variance))
        return None

    return lower_graph(rvs, factors, K, T,
                       factor_weight=lambda f: float(len(f.factors)),
                       arg_weight=arg_weight,
                       var_weight=lambda rv: float(len(rv.rvs)),
                       gaussian_evidence=gaussian_evidence)
