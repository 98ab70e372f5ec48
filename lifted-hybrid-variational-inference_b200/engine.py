"""Device-side state and launch sequence of one VI model (the hot path behind the drop-in
classes).

``DeviceEngine`` owns PyTorch tensors for the flat parameter vector, Adam moments, the record
groups and the gradient buffer, and drives the CUDA kernels of ``liblhvi.so`` through the
C ABI (``include/lhvi.h``).  One iteration (reference ``ADAM_update`` body,
``VarInference.py:249-287``) is:

    for every record group:  lhvi_factor_expect_grad   (node-entropy groups included; the
                                                        launches are independent and run on
                                                        parallel branches of the CUDA graph)
    lhvi_finish                                         -> G_w[K], free energy, step counter, and
                                                        (records sharded over GPUs) the sum over
                                                        ranks of [G_w | energy | gradients of the
                                                        shared variables] through peer memory
    lhvi_param_step                                     -> new eta, tau, w, moments; clears the
                                                        gradient slots it consumed

PyTorch is plumbing only (allocation, streams, torch.distributed).  There is no CPU path:
constructing the engine without CUDA or without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import gc
import os

import numpy as np
import torch

from . import _cabi
from .dist import PeerExchange, ShardPlan
from .lowering import LoweredModel, RecordGroup, fold_unary


def hermgauss_scaled(T):
    """Gauss-Hermite nodes and weights / sqrt(pi) (VarInference.py:18-19)."""
    x, w = np.polynomial.hermite.hermgauss(T)
    return x, w / np.sqrt(np.pi)


def hub_mask(g):
    """Bit a set when hidden argument a mostly repeats the previous record's variable (runs of
    a hub variable): the kernels then accumulate its gradient per block in shared memory."""
    mask = 0
    if g.n < 64:
        return mask
    for a in range(g.nh):
        col = g.poff[a]
        if np.count_nonzero(col[1:] == col[:-1]) > 0.5 * (g.n - 1):
            mask |= 1 << a
    return mask


FOLD_ALIGN_MIN_TILES = 4   # runs of a streamed group at least this long are padded to whole tiles
FOLD_BLOCK_SLOTS = 148 * 2  # resident blocks of the streaming kernel on one B200 (a hint: launch_unary_fold, LHVI_FOLD_BLOCKS)
ITER_BLOCK_SLOTS = 148 * 2  # resident blocks of the persistent iteration kernel (lhvi_iterate)
ITER_TUNE_ROUNDS = 2
PERSISTENT_MAX_RECORDS = 3_500_000   # LHVI_PERSISTENT=auto: larger rank-local models use the per-group launches


def align_runs(g, null_pot, tile, FOLD_BLOCK_SLOTS=FOLD_BLOCK_SLOTS):
    """Pad the long runs (records on one variable) of a streamed unary group to whole tiles with
    null records -- coefficient block ``null_pot`` (log psi = 0), zero weights, zero evidence, the
    run's own offset -- so that a hub boundary never falls inside a tile of the streaming kernel
    (a split tile sends its threads through the per-thread path: 4-8 us per launch).  The padded
    group is an ordinary group: every kernel, the generic one included, sums the same values."""
    key = g.poff[0]
    starts = np.concatenate([[0], np.flatnonzero(key[1:] != key[:-1]) + 1, [g.n]])
    lens = np.diff(starts)
    # a small group is launched with whole chunks of ceil(tiles / resident blocks) tiles per block
    # (launch_unary_fold): pad the long runs to that chunk, so that no block crosses a hub boundary
    unit = tile
    tiles0 = -(-g.n // tile)
    if tiles0 <= 4 * FOLD_BLOCK_SLOTS:
        for _ in range(4):
            chunk = max(1, -(-tiles0 // FOLD_BLOCK_SLOTS))
            unit = chunk * tile
            padded = np.where(lens >= FOLD_ALIGN_MIN_TILES * tile, (lens + unit - 1) // unit * unit, lens)
            tiles1 = -(-int(padded.sum()) // unit) * chunk
            if -(-tiles1 // FOLD_BLOCK_SLOTS) == chunk:
                break
            tiles0 = tiles1
    plen = np.where(lens >= FOLD_ALIGN_MIN_TILES * tile, (lens + unit - 1) // unit * unit, lens)
    total = int(plen.sum())
    plen[-1] += (total + unit - 1) // unit * unit - total           # the tail
    if int(plen.sum()) == g.n:
        return g
    dst0 = np.cumsum(plen) - plen
    pos = np.repeat(dst0 - starts[:-1], lens) + np.arange(g.n)      # where each record goes
    n2 = int(plen.sum())

    def spread(arr, fill):
        out = np.full(arr.shape[:-1] + (n2,), fill, dtype=arr.dtype)
        out[..., pos] = arr
        return out
    poff = np.repeat(key[starts[:-1]], plen)[None, :].astype(g.poff.dtype)
    return RecordGroup(g.nd, g.nc, g.ng, g.ne, g.dims, g.node, spread(g.pot, null_pot), poff,
                       spread(g.egval, 0.0), spread(g.egvar, 1.0), spread(g.ecval, 0.0),
                       spread(g.wf, 0.0), spread(g.gam, 0.0), spread(g.nscale, 0.0), g.weighted, g.pure, g.dvals, g.kind)


RUN_SPLIT = 64          # a longer run is split (one thread walks a run)
RUN_ACC_BYTES = 96 * 1024


RUN_THREAD_SLOTS = 148 * 2 * 256     # resident threads of the run-major kernel on one B200


def run_layout(g, K, elem_bytes):
    """Run-major form of a full group with two hidden continuous arguments, one of which takes
    few distinct values (lhvi.h, lhvi_group::run_*): returns (sorted group, run_start, run_key,
    run_hid, hub_keys, hub_arg) or None when the group does not qualify."""
    if g.node or g.pure or g.nd != 0 or g.nc != 2 or g.ng != 0 or g.ne > 1 or g.n < 2 or g.kind != 0:
        return None
    uniq = [np.unique(g.poff[a]) for a in (0, 1)]
    hub_arg = 0 if uniq[0].size < uniq[1].size else 1
    hubs = uniq[hub_arg]
    if hubs.size > _cabi.LHVI_RUN_MAX_HUBS or hubs.size * 2 * K * 256 * elem_bytes > RUN_ACC_BYTES:
        return None
    run_arg = 1 - hub_arg
    if g.n < 1.5 * uniq[run_arg].size:          # hardly any run: the record-major kernel is as good
        return None
    order = np.argsort(g.poff[run_arg], kind="stable")
    sg = g.take(order)
    keys = sg.poff[run_arg]
    starts = np.concatenate([[0], np.flatnonzero(keys[1:] != keys[:-1]) + 1, [sg.n]])
    # one thread walks a run: cap the run length, and when the group is small (a rank's shard of a
    # strong-scaled model) cut the runs further so that every resident thread has one
    split = int(min(RUN_SPLIT, max(2, -(-sg.n // RUN_THREAD_SLOTS))))
    if np.diff(starts).max() > split:
        pieces = [np.arange(a, b, split) for a, b in zip(starts[:-1], starts[1:])]
        starts = np.concatenate(pieces + [[sg.n]])
    run_key = keys[starts[:-1]]
    hid = np.searchsorted(hubs, sg.poff[hub_arg])
    return sg, starts.astype(np.int32), run_key.astype(np.int32), hid.astype(np.int32), hubs.astype(np.int32), hub_arg


def fuse_run_extras(groups, layouts):
    """Move the records of a run-major group's run variables that the run-major kernel can evaluate once per
    run -- the variable's node-entropy record and one pure unary quadratic factor (nd=0, nc=1, ng=0, ne=0) --
    out of their own groups and into per-run columns of the run-major group (``lhvi_group::run_node /
    run_una_pot / run_una_w``): the kernel has the variable's parameters and axis tables in registers at that
    point, while as separate groups they are two more launches that the register-bound big kernels do not
    overlap with (7 + 10 us of a 137 us iteration on the bench model, ``tools/small_group_probe.py``).
    Returns ``(groups, extras)``: the reduced group list (same length and order) and ``{group index:
    (run_node [2, n_runs] | None, run_una_pot [n_runs] | None, run_una_w [2, n_runs] | None, counts)}``.
    Sums are unchanged: every record is still evaluated once per iteration."""
    groups = list(groups)
    extras = {}
    for gi, lay in enumerate(layouts):
        if lay is None:
            continue
        _, _, run_key, _, _, _ = lay
        n_runs = int(run_key.size)
        first = np.ones(n_runs, dtype=bool)
        first[1:] = run_key[1:] != run_key[:-1]            # a run cut into pieces: the first piece carries the extras
        pos_first = np.flatnonzero(first)
        keys = run_key[pos_first].astype(np.int64)         # ascending (the group is sorted by this argument)
        node = np.zeros((2, n_runs))
        una_pot = np.full(n_runs, -1, dtype=np.int32)
        una_w = np.zeros((2, n_runs))
        n_node = n_una = 0
        for gj, h in enumerate(groups):
            if gj == gi or h.n == 0 or h.nd != 0 or h.ng != 0 or h.nc != 1 or h.ne != 0 or h.kind != 0:
                continue
            if not (h.node or h.pure):
                continue
            off = h.poff[0].astype(np.int64)
            at = np.minimum(np.searchsorted(keys, off), keys.size - 1)
            hit = keys[at] == off
            # one record per variable goes along; further records on the same variable stay where they are
            uniq_first = np.zeros(h.n, dtype=bool)
            uniq_first[np.unique(off, return_index=True)[1]] = True
            run = pos_first[at]
            free = (node[0, run] == 0) & (node[1, run] == 0) if h.node else (una_pot[run] < 0)
            take = hit & uniq_first & free
            if h.node:
                take &= (h.wf != 0) | (h.nscale != 0)       # (an all-zero node record contributes nothing either way)
            if not take.any():
                continue
            r = run[take]
            if h.node:
                node[0, r], node[1, r] = h.wf[take], h.nscale[take]
                n_node += int(take.sum())
            else:
                una_pot[r] = h.pot[take]
                una_w[0, r], una_w[1, r] = h.wf[take], h.gam[0][take]
                n_una += int(take.sum())
            groups[gj] = h.take(np.flatnonzero(~take))
        if n_node or n_una:
            extras[gi] = (node if n_node else None, una_pot if n_una else None, una_w if n_una else None, (n_node, n_una))
    return groups, extras


def fuse_constants(groups, streamed, ptab):
    """Records without any integrated argument (every argument observed: constants of the free energy and of
    G_w) leave their group and travel as extra columns of the largest streamed group (``lhvi_group::cst_*``):
    the streaming kernel evaluates one of them per thread and tile beside its own records, instead of one more
    launch behind the register-bound kernels (8.6 us of the bench iteration, ``tools/small_group_probe.py``).
    Like ``fold_unary`` for the streamed records, the point evidence is substituted into the record's quadratic
    here (``cst_q``); the floor and the weight are applied to every record in every pass on the device.
    Returns ``(groups, {streamed group index: (q, wf | None)})``."""
    groups = list(groups)
    target = max((i for i in range(len(groups)) if streamed[i]), key=lambda i: groups[i].n, default=None)
    if target is None:
        return groups, {}
    qs, wfs, weighted = [], [], False
    for gj, h in enumerate(groups):
        if gj == target or h.n == 0 or h.node or not h.pure or h.nd or h.nc or h.ng or h.kind != 0 or h.ne < 1:
            continue
        nct = h.ne
        ncoef = (nct + 1) * (nct + 2) // 2
        coef = ptab[h.pot.astype(np.int64)[:, None] + np.arange(ncoef)[None, :]]
        q = coef[:, 0].copy()
        for i in range(nct):
            q += coef[:, 1 + i] * h.ecval[i]
        p = 1 + nct
        for i in range(nct):
            for j in range(i, nct):
                q += coef[:, p] * h.ecval[i] * h.ecval[j]
                p += 1
        qs.append(q)
        wfs.append(h.wf)
        weighted = weighted or bool(h.weighted)
        groups[gj] = h.take(np.zeros(0, dtype=np.int64))
    if not qs:
        return groups, {}
    return groups, {target: (np.concatenate(qs), np.concatenate(wfs) if weighted else None)}


# ---- schedule of the persistent iteration kernel (lhvi_iterate) ---------------------------------
# Per record group: nanoseconds per record for ONE block of 256 threads, and a fixed prologue +
# epilogue latency per block and iteration in microseconds (hub tables, shared-memory set-up, block
# reduction, flushes) -- rough figures from the per-group kernels timed alone on a B200 (fp32, K = 3;
# profiles/r2_iter_plan.md).  Only the split of the grid depends on them, never a result.
ITER_COST_NS = {"run": 6.5, "fold": 1.1, "node": 5.6, "pun": 8.8, "const": 3.0, "full": 10.0}
ITER_FIXED_US = {"run": 3.5, "fold": 2.5, "node": 2.5, "pun": 2.5, "const": 2.5, "full": 2.5}


def iteration_kind(g, runs, streamed):
    if g.node:
        return "node"
    if runs is not None:
        return "run"
    if streamed:
        return "fold"
    if g.pure:
        return "pun" if g.nc == 1 else "const"
    return "full"


def iteration_cost(kind, n, K, esize):
    """Estimated block-microseconds of one pass over a group of ``n`` records (before any measurement)."""
    c = ITER_COST_NS[kind]
    if kind in ("run", "node", "full"):             # walks over K x K cross densities
        c *= (K / 3.0) ** 2 * (2.5 if esize == 8 else 1.0)
    else:                                            # streaming / closed forms: bytes and a few FMAs
        c *= (K / 3.0) * (2.0 if esize == 8 else 1.0)
    return c * 1e-3 * n


def iteration_plan(work, fixed, blocks, pinned=None):
    """Blocks of the persistent grid per record group so that the groups, running side by side, end
    together: the smallest T with sum_p ceil(work_p / (T - fixed_p)) <= blocks (``work`` in
    block-microseconds, ``fixed`` = per-block prologue + epilogue in microseconds).  ``pinned``
    {index: blocks} keeps those groups' shares (a streamed group whose runs were padded to its
    share of the grid)."""
    pinned = pinned or {}
    free = [i for i in range(len(work)) if i not in pinned]
    budget = blocks - sum(pinned.values())
    plan = [pinned.get(i, 1) for i in range(len(work))]
    if not free:
        return plan
    budget = max(budget, len(free))

    def need(T):
        return [max(1, int(np.ceil(work[i] / (T - fixed[i])))) for i in free]
    lo, hi = max(fixed[i] for i in free) + 1e-3, max(fixed[i] for i in free) + sum(work[i] for i in free) + 1.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if sum(need(mid)) <= budget:
            hi = mid
        else:
            lo = mid
    for i, b in zip(free, need(hi)):
        plan[i] = b
    # hand the blocks the ceilings left over to the groups with the most work per block
    spare = budget - sum(plan[i] for i in free)
    while spare > 0:
        i = max(free, key=lambda j: work[j] / plan[j])
        plan[i] += 1
        spare -= 1
    return plan


def h2_tables(g):
    """``lhvi_h2`` of a full group with hidden discrete arguments (compat="reference"): the domain
    values of every such argument and, per pair (a, b), the state of b whose value is a's j-th value.
    Raises where the unmodified reference's result is not reproduced (include/lhvi.h)."""
    if g.ng != 0:
        raise NotImplementedError("compat='reference': factors with Gaussian evidence next to a hidden discrete argument")
    if len(g.dvals) != g.nd or any(len(v) != D for v, D in zip(g.dvals, g.dims)):
        raise NotImplementedError("compat='reference' needs the domain values of the hidden discrete arguments "
                                  "(models lowered from object graphs carry them)")
    h = _cabi.LhviH2()
    for a in range(g.nd):
        if any(np.isnan(v) for v in g.dvals[a]):
            raise NotImplementedError("compat='reference': non-numeric domain values")
        if g.nc > 0 and g.dims[a] != 2:
            raise NotImplementedError(
                "compat='reference': a hidden discrete argument with more than two states next to a hidden "
                "continuous one (the reference pairs the domain values with (mu, var) by position)")
        for j, v in enumerate(g.dvals[a]):
            h.dvals[a][j] = float(v)
        for b in range(g.nd):
            if b == a:
                continue
            if g.dims[b] != g.dims[a]:
                raise NotImplementedError("compat='reference': discrete arguments of different cardinality in one "
                                          "factor (the reference's pairing depends on the argument order)")
            for j, v in enumerate(g.dvals[a]):
                if v not in g.dvals[b]:
                    raise ValueError("compat='reference': a value of one discrete argument is not in the domain of "
                                     "another argument of the same factor (the unmodified reference raises here)")
                h.xmap[a][b][j] = g.dvals[b].index(v)
    return h



class DeviceEngine:
    def __init__(self, model: LoweredModel, dtype="float64", device=None, var_threshold=0.1,
                 process_group=None, shard=True, force_generic=False, run_major=True, compat=None):
        if not torch.cuda.is_available():
            raise RuntimeError(
                "lhvi: no CUDA device visible. The variational-inference update loop runs only "
                "on the GPU (hand-written sm_100a kernels); there is no CPU fallback.")
        self.lib = _cabi.load()
        self.full_model = model
        self.K, self.T = model.K, model.T
        self.dtype_name = "float64" if str(dtype) in ("float64", "torch.float64", "f64", "double") else "float32"
        self.tdtype = torch.float64 if self.dtype_name == "float64" else torch.float32
        self.dcode = _cabi.LHVI_F64 if self.dtype_name == "float64" else _cabi.LHVI_F32
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.var_threshold = float(var_threshold)
        self.force_generic = bool(force_generic)
        self.run_major = bool(run_major)     # False: keep every group record-major (tests)
        # compat="reference": the categorical gradients as the UNMODIFIED reference computes them (its
        # gradient_category_tau builds the other arguments' axes from the wrong domain, VarInference.py:147-150;
        # SURVEY hazard H2) instead of the intended mathematics -- lhvi_category_grad_reference
        if compat not in (None, "reference"):
            raise ValueError(f"compat must be None or 'reference', got {compat!r}")
        self.compat = compat
        self.b1, self.b2, self.eps = 0.9, 0.999, 1e-8      # VarInference.py:223-225
        self.profile_group = None
        self.dom_events = []
        self.use_graph = True          # replay one captured iteration instead of ~8 launches
        # lhvi_iterate: n iterations in one persistent cooperative launch whenever every record group
        # has a body in that kernel (else the captured per-group launches); LHVI_PERSISTENT=0 turns it off
        # ("auto", the default: for rank-local models up to PERSISTENT_MAX_RECORDS records -- above that
        # the per-group kernels, which run at twice the occupancy, win: profiles/r2_iter_plan.md)
        self.persistent_mode = os.environ.get("LHVI_PERSISTENT", "auto")
        self.use_persistent = self.persistent_mode != "0" and not self.force_generic and self.compat is None
        self.persistent_min_iters = int(os.environ.get("LHVI_PERSISTENT_MIN_ITERS", "2"))
        self._iter_adapt = os.environ.get("LHVI_ITER_ADAPT", "1") != "0"
        self._iter_trace_env = bool(os.environ.get("LHVI_ITER_TRACE"))
        self._optim_cache = {}
        self.launch_count = 0          # kernels of liblhvi.so launched by iterate() so far
        self.parallel_groups = True    # independent group launches on parallel graph branches
        # lhvi_finish_step instead of lhvi_finish + lhvi_param_step: on for one GPU (measured 143.1 ->
        # 141.7 us per iteration), off for several unless LHVI_FUSED_STEP=1 (measured 89.5 -> 91.6 us
        # at two GPUs: block 0 then carries reduction + exchange + shared variables alone)
        self.fused_step = os.environ.get("LHVI_FUSED_STEP", "auto")
        self._graphs = {}
        self._side_streams = []
        self._grad_clean = False       # True: every gradient slot is zero (param_step cleared them)

        self.plan = ShardPlan(process_group, enabled=shard)
        self.world, self.rank = self.plan.world, self.plan.rank
        self.model = self.plan.shard(model)
        if self.persistent_mode == "auto" and self.model.n_records > PERSISTENT_MAX_RECORDS:
            self.use_persistent = False
        self.reduce_grads = self.plan.active

        self._upload()
        self.exchange = None           # "p2p" | "collective" when the records are sharded
        self.peer = None
        self._closed = False
        if self.plan.active:
            self._setup_exchange()

    def _setup_exchange(self):
        """Exchange of [G_w | energy | shared-variable gradients] between the ranks: inside
        lhvi_finish over NVLink peer memory when every rank can map its peers' buffers (one
        node), else one all-reduce of the compact vector (LHVI_EXCHANGE=collective forces it)."""
        K = self.K
        idx = np.asarray(self.plan.shared_idx, dtype=np.int64)
        if idx.size and idx.max() >= 2 ** 31:
            raise ValueError("parameter vector too large for 32-bit exchange offsets")
        self.xidx = self._dev(idx.astype(np.int32))
        tail = np.arange(self.n_param, self.n_param + K + 1, dtype=np.int64)
        self.xidx_all = self._dev(np.concatenate([tail, idx]))
        self.xbuf = torch.zeros(int(idx.size) + K + 1, dtype=self.tdtype, device=self.device)
        mode = os.environ.get("LHVI_EXCHANGE", "auto")
        self.exchange = "collective"
        if mode != "collective" and self.device.type == "cuda":
            self.peer = PeerExchange.create(self.plan, self.lib, int(idx.size), K,
                                            self.xidx.data_ptr(), self.tdtype, self.device)
            if self.peer is not None:
                self.exchange = "p2p"
            elif mode == "p2p":
                raise RuntimeError("lhvi: LHVI_EXCHANGE=p2p but the peers' buffers could not be mapped")

    def close(self):
        """Release what the engine holds outside PyTorch's allocator: the captured CUDA graphs and
        the peer-visible exchange buffer with its IPC mappings.  Collective when the records are
        sharded (every rank must call it: the peers' mappings are closed behind a barrier).  Called
        by the drop-in classes and the array engines before they replace an engine."""
        if getattr(self, "_closed", True):
            return
        self._closed = True
        torch.cuda.synchronize(self.device)
        self._graphs.clear()
        if self.peer is not None:
            import torch.distributed as dist
            dist.barrier(group=self.plan.group)       # nobody is still writing into a peer's buffer
            self.peer.close()
            self.peer = None

    def __del__(self):
        # graphs are released by their own destructors; the peer buffer needs the explicit call
        # (close() is collective), so only the local, non-collective part is done here
        try:
            if not getattr(self, "_closed", True) and self.peer is not None:
                torch.cuda.synchronize(self.device)
                self.peer.close()
        except Exception:
            pass

    # ---- buffers --------------------------------------------------------------------------
    def _dev(self, arr, dtype=None):
        t = torch.as_tensor(np.ascontiguousarray(arr))
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=False).contiguous()

    def _upload(self):
        m, K = self.model, self.K
        n_param = int(m.n_param)
        self.n_param = n_param
        qx, qw = hermgauss_scaled(self.T)
        self.quad = self._dev(np.concatenate([qx, qw]), self.tdtype)
        ptab_host = np.asarray(m.ptab, dtype=np.float64)
        # variable table of the optimiser step: the variables only this rank touches first, the
        # shared ones (whose gradients are completed by the exchange) last -- lhvi_finish_step
        order = np.arange(int(m.n_vars))
        self.n_owned = int(m.n_vars)
        if self.plan.active:
            shared_off = self.full_model.var_off[self.plan.part.shared]
            is_shared = np.isin(m.var_off, shared_off)
            order = np.concatenate([np.flatnonzero(~is_shared), np.flatnonzero(is_shared)])
            self.n_owned = int((~is_shared).sum())
        # every variable continuous, slots contiguous in variable order, one GPU: lhvi_finish_step steps one
        # 16-byte vector per thread without the variable table (include/lhvi.h, "uniform continuous slots")
        slot = 2 if 2 * K <= 2 else (2 * K + 3) // 4 * 4
        self.uniform_slots = bool(
            not self.plan.active and int(m.n_vars) > 0 and not np.any(m.var_kind != 0)
            and (K >= 2 or self.dtype_name == "float64")
            and np.array_equal(m.var_off.astype(np.int64), np.arange(int(m.n_vars), dtype=np.int64) * slot)
            and int(m.n_vars) * slot <= n_param and os.environ.get("LHVI_UNIFORM_STEP", "1") != "0")
        self.var_kind = self._dev(m.var_kind[order].astype(np.uint8))
        self.var_dim = self._dev(m.var_dim[order].astype(np.int32))
        self.var_off = self._dev(m.var_off[order].astype(np.int32))
        self.n_vars = int(m.n_vars)

        def z(n, dt=None):
            return torch.zeros(int(n), dtype=dt or self.tdtype, device=self.device)
        self.eta = z(n_param)
        self.tau = z(n_param)
        self.mom1 = z(n_param)
        self.mom2 = z(n_param)
        self.grad = z(n_param + K + 1)
        self.wstate = z(5 * K)
        self.wstate[K:2 * K] = 1.0 / K
        self.step = z(4, torch.float64)

        # the streaming unary kernel uses only the even moments of the quadrature rule
        self.symmetric_rule = bool(abs(np.dot(qw, qx)) < 1e-13 and abs(np.dot(qw, qx ** 3)) < 1e-13)
        # exact mirror symmetry (what hermgauss returns): the specialised walk pairs the mirror nodes
        self.mirror_rule = bool(np.array_equal(qx, -qx[::-1]) and np.array_equal(qw, qw[::-1])
                                and (self.T % 2 == 0 or qx[self.T // 2] == 0.0))

        self.groups = []      # (descriptor struct, tensors kept alive, RecordGroup)
        esize = 8 if self.dtype_name == "float64" else 4
        null_pot = {}           # nct -> offset of an all-zero coefficient block appended to ptab
        tile = _cabi.LHVI_FOLD_TILE

        def is_streamed(g):
            return (self.symmetric_rule and g.n > 0 and not g.node and g.pure and g.nd == 0 and g.nc == 1
                    and g.ng == 0 and g.kind == 0 and K <= 3 and bool((hub_mask(g) >> g.nd) & 1))
        layouts = [run_layout(g, K, esize) if (self.run_major and self.mirror_rule and self.T == 3 and K <= 3) else None
                   for g in self.model.groups]
        # records of the run variables themselves go into the run-major kernel (fuse_run_extras)
        model_groups, self.fused_extras = list(self.model.groups), {}
        if os.environ.get("LHVI_FUSE_RUN_EXTRAS", "1") != "0" and not self.force_generic and self.compat is None:
            model_groups, self.fused_extras = fuse_run_extras(model_groups, layouts)
        self.fused_constants = {}
        if os.environ.get("LHVI_FUSE_CONSTANTS", "1") != "0" and not self.force_generic and self.compat is None:
            model_groups, self.fused_constants = fuse_constants(model_groups, [is_streamed(g) for g in model_groups],
                                                                np.asarray(m.ptab, dtype=np.float64))
        # schedule of the persistent kernel: "split" gives every group blocks of its own (the groups run
        # side by side, as on the branches of the CUDA graph), "slice" lets every block walk all groups
        self.iter_plan = None
        mode = os.environ.get("LHVI_ITER_PLAN", "split")
        live = [i for i, g in enumerate(model_groups) if g.n > 0]
        if self.use_persistent and mode != "slice" and live and all(model_groups[i].nd == 0 for i in live):
            kinds = [iteration_kind(model_groups[i], layouts[i], is_streamed(model_groups[i])) for i in live]
            blocks = int(os.environ.get("LHVI_ITER_BLOCKS", ITER_BLOCK_SLOTS))
            if len(live) <= blocks:
                work = [iteration_cost(k, model_groups[i].n, K, esize) for k, i in zip(kinds, live)]
                shares = iteration_plan(work, [ITER_FIXED_US[k] for k in kinds], blocks)
                self.iter_plan = dict(zip(live, shares))
                self.iter_kinds = dict(zip(live, kinds))
                self.iter_work = dict(zip(live, work))
        for gi, g in enumerate(model_groups):
            keep = {}
            g_report = g        # what bench.py and callers see: the group as lowered
            d = _cabi.LhviGroup()
            runs = layouts[gi]
            if runs is not None:
                g = runs[0]
            d.nd, d.nc, d.ng, d.ne = g.nd, g.nc, g.ng, g.ne
            for i in range(_cabi.LHVI_MAX_AXES):
                d.dims[i] = int(g.dims[i]) if i < len(g.dims) else 0
            d.node, d.weighted, d.n = int(g.node), int(g.weighted), int(g.n)
            d.hub_mask = hub_mask(g)
            d.pure = int(g.pure)
            d.pot_kind = int(g.kind)
            h2 = None
            if self.compat == "reference" and g.nd > 0 and not g.node and not g.pure and g.n > 0:
                h2 = h2_tables(g)
                d.no_category_grad = 1

            d.iter_blocks = int(self.iter_plan[gi]) if (self.iter_plan and gi in self.iter_plan) else 0
            streamed = is_streamed(g)
            if streamed:
                nct = g.nc + g.ne
                if nct not in null_pot:
                    null_pot[nct] = ptab_host.size
                    ptab_host = np.concatenate([ptab_host, np.zeros((nct + 1) * (nct + 2) // 2)])
                # the blocks that will stream this group: its share of the persistent grid, else the
                # whole grid (sliced schedule) or the streaming kernel's own resident blocks
                slots = d.iter_blocks or (ITER_BLOCK_SLOTS if self.use_persistent else FOLD_BLOCK_SLOTS)
                g = align_runs(g, null_pot[nct], tile, slots)
                d.n = int(g.n)
            fold = fold_unary(g, ptab_host) if (self.symmetric_rule and g.n > 0) else None
            n_pad = (g.n + tile - 1) // tile * tile if fold is not None else g.n

            def put(name, arr, dt, pad=None):
                if arr.size == 0:
                    return None
                if pad is not None and n_pad > g.n:          # single-column arrays of a folded group
                    arr = np.concatenate([arr.reshape(-1), np.full(n_pad - g.n, pad, dtype=arr.dtype)])
                keep[name] = self._dev(arr, dt)
                return keep[name].data_ptr()
            padded = fold is not None
            d.pot = None if g.node else put("pot", g.pot, torch.int32)
            d.poff = put("poff", g.poff, torch.int32, pad=g.poff[0, -1] if padded else None)
            d.egval = put("egval", g.egval, self.tdtype)
            d.egvar = put("egvar", g.egvar, self.tdtype)
            d.ecval = put("ecval", g.ecval, self.tdtype)
            d.wf = put("wf", g.wf, self.tdtype, pad=0.0 if padded else None) if g.weighted else None
            d.gam = put("gam", g.gam, self.tdtype, pad=0.0 if padded else None) if g.weighted else None
            d.nscale = put("nscale", g.nscale, self.tdtype) if g.node else None
            d.fold, d.n_pad = None, 0
            if padded:
                cols = np.zeros((3, n_pad))
                cols[:, :g.n] = fold
                d.fold, d.n_pad = put("fold", cols, self.tdtype), n_pad
            if gi in self.fused_constants and padded:
                cq, cwf = self.fused_constants[gi]
                keep["cst_q"] = self._dev(cq, self.tdtype)
                d.cst_q, d.cst_n = keep["cst_q"].data_ptr(), int(cq.size)
                if cwf is not None:
                    keep["cst_wf"] = self._dev(cwf, self.tdtype)
                    d.cst_wf = keep["cst_wf"].data_ptr()
                keep["fused_constants"] = int(cq.size)
            d.run_start, d.n_runs, d.n_hubs, d.run_hub_arg = None, 0, 0, 0
            if runs is not None:
                _, starts, run_key, hid, hubs, hub_arg = runs
                for name, arr in (("run_start", starts), ("run_key", run_key), ("run_hid", hid), ("hub_keys", hubs)):
                    keep[name] = self._dev(arr, torch.int32)
                    setattr(d, name, keep[name].data_ptr())
                d.n_runs, d.n_hubs, d.run_hub_arg = int(run_key.size), int(hubs.size), int(hub_arg)
                if gi in self.fused_extras:
                    node, una_pot, una_w, counts = self.fused_extras[gi]
                    if node is not None:
                        keep["run_node"] = self._dev(node, self.tdtype)
                        d.run_node = keep["run_node"].data_ptr()
                    if una_pot is not None:
                        keep["run_una_pot"] = self._dev(una_pot, torch.int32)
                        keep["run_una_w"] = self._dev(una_w, self.tdtype)
                        d.run_una_pot, d.run_una_w = keep["run_una_pot"].data_ptr(), keep["run_una_w"].data_ptr()
                    keep["fused"] = counts
            if h2 is not None:
                keep["h2"] = h2
            self.groups.append((d, keep, g_report))

        self.ptab = self._dev(ptab_host, self.tdtype)
        rows = max(1, len(self.groups)) * _cabi.LHVI_PARTIAL_ROWS
        self.partial_rows = rows
        self.partials = z(rows * (K + 1), torch.float64)

        md = _cabi.LhviModel()
        md.dtype, md.K, md.T, md.n_param = self.dcode, K, self.T, n_param
        md.rule_symmetric = int(self.mirror_rule)
        for i, v in enumerate(np.concatenate([qx, qw])):
            md.quad_host[i] = float(v)
        md.quad, md.ptab = self.quad.data_ptr(), self.ptab.data_ptr()
        md.eta, md.w = self.eta.data_ptr(), self.wstate[K:2 * K].data_ptr()
        md.grad, md.partials = self.grad.data_ptr(), self.partials.data_ptr()
        self.desc = md
        self.launches_per_pass = 0
        self.group_table = (_cabi.LhviGroup * max(1, len(self.groups)))()
        for i, (d, _, _) in enumerate(self.groups):
            C.memmove(C.byref(self.group_table, i * C.sizeof(_cabi.LhviGroup)), C.byref(d), C.sizeof(_cabi.LhviGroup))
        self.sm_count = torch.zeros(512, dtype=torch.int32, device=self.device)
        self.iter_accum = torch.zeros(K + 1, dtype=torch.float64, device=self.device)
        self._persistent = None        # lhvi_iterate_supported, asked once
        self._iter_tuned = 0           # how often the grid split has been re-planned from a measured launch

    # ---- state exchange with the host -----------------------------------------------------
    @property
    def w_tau(self):
        return self.wstate[:self.K]

    @property
    def w(self):
        return self.wstate[self.K:2 * self.K]

    def set_state(self, eta, tau, w_tau):
        """``eta``: (mu,var) on continuous slots, probabilities on discrete ones; ``tau``:
        logits on discrete slots; ``w_tau``: mixture logits.  w = softmax(w_tau) is derived."""
        K = self.K
        self.eta.copy_(torch.as_tensor(np.asarray(eta, dtype=np.float64)).to(self.tdtype))
        self.tau.copy_(torch.as_tensor(np.asarray(tau, dtype=np.float64)).to(self.tdtype))
        wt = np.asarray(w_tau, dtype=np.float64)
        e = np.e ** wt
        self.wstate[:K].copy_(torch.as_tensor(wt).to(self.tdtype))
        self.wstate[K:2 * K].copy_(torch.as_tensor(e / e.sum()).to(self.tdtype))

    def get_state(self):
        """(eta, tau, w_tau, w) on the host; with sharded records the ranks' pieces (owned
        variables from their owner, shared ones are identical everywhere) are merged first."""
        K = self.K
        self.check_exchange()
        ws = self.wstate.double().cpu().numpy()
        return (self.plan.merge(self.eta).double().cpu().numpy(),
                self.plan.merge(self.tau).double().cpu().numpy(),
                ws[:K].copy(), ws[K:2 * K].copy())

    # ---- compact host layout (what a reference caller holds: eta[rv] as K x 2 / K x D arrays) ------
    def packed_map(self, local=False):
        """Element offsets of the used slot elements, ascending (device int32) -- the compact
        layout of ``pack_state`` / ``unpack_state``; host copy in ``self.packed_index``.
        ``local=True`` (sharded records): only the variables this rank steps -- the ones it owns
        plus the shared ones -- so that N ranks move 1/N of the state each instead of all of it."""
        if getattr(self, "_packed_map", None) is None or getattr(self, "_packed_local", None) != bool(local):
            m, K = (self.model if local else self.full_model), self.K
            sizes = K * m.var_dim.astype(np.int64)
            off = m.var_off.astype(np.int64)
            idx = np.repeat(off - (np.cumsum(sizes) - sizes), sizes) + np.arange(int(sizes.sum()))
            self.packed_index = idx
            self._packed_map = self._dev(idx.astype(np.int32))
            self._packed_local = bool(local)
        return self._packed_map

    def unpack_state(self, packed, which="eta"):
        """Device tensor ``packed`` (compact layout) -> the padded slot vector ``eta`` / ``tau``."""
        mp = self.packed_map(getattr(self, "_packed_local", False))
        if packed.numel() != mp.numel():
            raise ValueError(f"lhvi: packed state has {packed.numel()} elements, the compact layout {mp.numel()}")
        dst = self.eta if which == "eta" else self.tau
        _cabi.check(self.lib.lhvi_state_unpack(self.dcode, mp.numel(), mp.data_ptr(), packed.data_ptr(),
                                               dst.data_ptr(), self._stream()), self.lib)

    def pack_state(self, packed, which="eta"):
        """The padded slot vector ``eta`` / ``tau`` -> device tensor ``packed`` (compact layout)."""
        mp = self.packed_map(getattr(self, "_packed_local", False))
        if packed.numel() != mp.numel():
            raise ValueError(f"lhvi: packed state has {packed.numel()} elements, the compact layout {mp.numel()}")
        src = self.eta if which == "eta" else self.tau
        _cabi.check(self.lib.lhvi_state_pack(self.dcode, mp.numel(), mp.data_ptr(), src.data_ptr(),
                                             packed.data_ptr(), self._stream()), self.lib)

    def check_exchange(self):
        if self.peer is not None and self.peer.timed_out():
            raise _cabi.LhviError("lhvi: a rank did not reach the gradient exchange within the "
                                  "spin limit (peer exchange timed out)")

    def reset_moments(self):
        K = self.K
        self.mom1.zero_()
        self.mom2.zero_()
        self.wstate[2 * K:].zero_()
        self.step.zero_()

    def get_moments(self):
        K = self.K
        ws = self.wstate.double().cpu().numpy()
        return (self.plan.merge(self.mom1).double().cpu().numpy(),
                self.plan.merge(self.mom2).double().cpu().numpy(),
                ws[2 * K:3 * K].copy(), ws[3 * K:4 * K].copy(), float(self.step[0].item()))

    def set_moments(self, mom1, mom2, m_w, u_w, t):
        K = self.K
        self.mom1.copy_(torch.as_tensor(np.asarray(mom1, dtype=np.float64)).to(self.tdtype))
        self.mom2.copy_(torch.as_tensor(np.asarray(mom2, dtype=np.float64)).to(self.tdtype))
        self.wstate[2 * K:3 * K].copy_(torch.as_tensor(np.asarray(m_w, dtype=np.float64)).to(self.tdtype))
        self.wstate[3 * K:4 * K].copy_(torch.as_tensor(np.asarray(u_w, dtype=np.float64)).to(self.tdtype))
        t = float(t)
        self.step.copy_(torch.tensor([t, 1 - self.b1 ** t, 1 - self.b2 ** t, 0.0], dtype=torch.float64))

    # ---- the hot path ---------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _launch_group(self, i, stream):
        d, keep, _ = self.groups[i]
        if int(d.n) == 0:
            # a group whose records all went into another group's kernel (fuse_run_extras / fuse_constants): its
            # region of `partials` keeps the zero header it was allocated with -- no valid rows, nothing to launch
            return
        _cabi.check(self.lib.lhvi_factor_expect_grad(
            C.byref(self.desc), C.byref(d), i * _cabi.LHVI_PARTIAL_ROWS, int(self.force_generic),
            C.c_void_p(stream.cuda_stream)), self.lib)
        if "h2" in keep:
            _cabi.check(self.lib.lhvi_category_grad_reference(
                C.byref(self.desc), C.byref(d), C.byref(keep["h2"]), C.c_void_p(stream.cuda_stream)), self.lib)

    def _tick(self, stream):
        _cabi.check(self.lib.lhvi_step_tick(self.step.data_ptr(), self.b1, self.b2,
                                            C.c_void_p(stream.cuda_stream)), self.lib)

    def _launch_groups(self, tick=False):
        """One launch per record group (and, with ``tick``, the step-counter kernel beside them).  The launches only share atomics into ``grad``, so
        they are forked onto side streams (parallel branches once captured in a CUDA graph),
        largest group first; with ``profile_group`` set they run in order on the current
        stream with CUDA events around that group's launch (bench.py roofline)."""
        main = torch.cuda.current_stream(self.device)
        n = len(self.groups)
        n_launched = sum(1 for d, _, _ in self.groups if int(d.n) > 0)      # (empty groups are not launched)
        if self.profile_group is not None or not self.parallel_groups or n < 2:
            if tick:
                self._tick(main)
            for i in range(n):
                timed = self.profile_group == i or self.profile_group == "all"
                if timed:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record(main)
                self._launch_group(i, main)
                if timed:
                    ev[1].record(main)
                    if int(self.groups[i][0].n) > 0:
                        self.dom_events.append((i, ev[0], ev[1]))
            return n_launched
        while len(self._side_streams) < n:
            self._side_streams.append(torch.cuda.Stream(device=self.device))
        order = sorted(range(n), key=lambda i: -self.groups[i][2].n)
        fork = torch.cuda.Event()
        fork.record(main)
        self._launch_group(order[0], main)
        for j, i in enumerate(order[1:]):
            side = self._side_streams[j]
            side.wait_event(fork)
            self._launch_group(i, side)
            join = torch.cuda.Event()
            join.record(side)
            main.wait_event(join)
        if tick:
            side = self._side_streams[n - 1]
            side.wait_event(fork)
            self._tick(side)
            join = torch.cuda.Event()
            join.record(side)
            main.wait_event(join)
        return n_launched

    def _finish(self, tick, exchange):
        lib, st = self.lib, self._stream()
        x = C.byref(self.peer.desc) if (exchange and self.exchange == "p2p") else None
        step = self.step.data_ptr() if tick else None
        _cabi.check(lib.lhvi_finish(C.byref(self.desc), self.partial_rows, step, self.b1, self.b2, x, st), lib)
        if exchange and self.exchange == "collective":
            torch.index_select(self.grad, 0, self.xidx_all, out=self.xbuf)
            self.plan.all_reduce(self.xbuf)
            self.grad.index_copy_(0, self.xidx_all, self.xbuf)

    def grad_pass(self):
        """Fill ``self.grad`` = [parameter gradients | raw G_w | free energy] at the current
        parameters (summed over the process group when the records are sharded)."""
        self.grad.zero_()
        self._grad_clean = False
        launches = self._launch_groups()
        self._finish(tick=False, exchange=False)
        if self.reduce_grads:
            self.plan.all_reduce(self.grad)
        self.launches_per_pass = launches + 1
        return self.grad

    def param_step(self, lr, sgd=False, zero_grad=False):
        lib, st = self.lib, self._stream()
        _cabi.check(lib.lhvi_param_step(
            self.dcode, self.K, self.n_vars, self.var_kind.data_ptr(), self.var_dim.data_ptr(),
            self.var_off.data_ptr(), self.eta.data_ptr(), self.tau.data_ptr(), self.grad.data_ptr(),
            self.n_param, self.mom1.data_ptr(), self.mom2.data_ptr(), self.wstate.data_ptr(),
            self.step.data_ptr(), float(lr), self.b1, self.b2, self.eps, self.var_threshold,
            int(bool(sgd)), int(bool(zero_grad)), st), lib)

    def _can_fuse(self):
        if self.fused_step == "0":
            return False
        if not self.plan.active:
            return True
        return self.fused_step == "1" and self.exchange == "p2p" and int(self.peer.desc.blocks) == 1

    def _iteration(self, lr, sgd):
        """One Jacobi iteration; expects clean gradient slots and leaves them clean."""
        if self._can_fuse():
            # the step counter is advanced beside the factor kernels; finish + step are one launch
            launches = self._launch_groups(tick=not sgd)
            x = C.byref(self.peer.desc) if self.plan.active else None
            tables = ((None, None, None) if self.uniform_slots else
                      (self.var_kind.data_ptr(), self.var_dim.data_ptr(), self.var_off.data_ptr()))
            _cabi.check(self.lib.lhvi_finish_step(
                C.byref(self.desc), self.partial_rows, x, self.n_vars, self.n_owned, *tables,
                self.tau.data_ptr(), self.mom1.data_ptr(), self.mom2.data_ptr(), self.wstate.data_ptr(),
                self.step.data_ptr(), float(lr), self.b1, self.b2, self.eps, self.var_threshold,
                int(bool(sgd)), self._stream()), self.lib)
            self.launches_per_pass = launches + (0 if sgd else 1)
            return
        launches = self._launch_groups()
        self._finish(tick=not sgd, exchange=self.plan.active)
        self.param_step(lr, sgd=sgd, zero_grad=True)
        self.launches_per_pass = launches + 1

    def persistent(self):
        """True when ``iterate`` runs as one persistent launch (``lhvi_iterate``)."""
        if not self.use_persistent:
            return False
        if self._persistent is None:
            x = C.byref(self.peer.desc) if (self.plan.active and self.exchange == "p2p") else None
            ok = (not self.plan.active) or self.exchange == "p2p"       # the collective exchange is a host-side call
            self._persistent = bool(ok and self.lib.lhvi_iterate_supported(
                C.byref(self.desc), self.group_table, len(self.groups), x))
            if self._persistent:
                resident = int(self.lib.lhvi_iterate_blocks(C.byref(self.desc), self.group_table, len(self.groups), x))
                if resident < 0:
                    _cabi.check(resident, self.lib)
                if self.persistent_mode == "auto" and 0 < resident < ITER_BLOCK_SLOTS:
                    # a group whose shared memory leaves one block per SM (the fp64 run-major tables: 122 KB)
                    # halves the persistent grid: the per-group launches are faster then (1.4 M records,
                    # fp64: 167 us against 119 us)
                    self._persistent = False
                elif self.iter_plan and 0 < resident < sum(self.iter_plan.values()):
                    # the split was planned for 2 blocks per SM: plan again for what is resident
                    self._replan(max(resident, len(self.iter_plan)))
        return self._persistent

    def _replan(self, blocks, pinned=None):
        live = sorted(self.iter_plan)
        pin = {live.index(i): b for i, b in (pinned or {}).items()}
        shares = iteration_plan([self.iter_work[i] for i in live], [ITER_FIXED_US[self.iter_kinds[i]] for i in live],
                                blocks, pin)
        self.iter_plan = dict(zip(live, shares))
        for i in live:
            self.group_table[i].iter_blocks = int(self.iter_plan[i])

    def _retune(self, trace, iters):
        """Re-split the persistent grid from what the groups' blocks measured in the launch just done
        (``lhvi_optim::trace``, last traced iteration): work_p = (mean time in the group - fixed) x its
        blocks.  A streamed group keeps its share (its hub runs are padded to it)."""
        blocks = sum(self.iter_plan.values())
        grid = int(np.count_nonzero(trace.reshape(-1, 16)[:, 0])) // iters       # blocks actually launched
        if grid < 1:
            return
        t = trace[:iters * grid * 16].reshape(iters, grid, 16)[-1]
        live = sorted(self.iter_plan)
        for p, i in enumerate(live):
            if p >= 11:
                continue
            rows = t[t[:, 1 + p] > 0]                  # the blocks that worked on group p
            if rows.shape[0] == 0:
                continue
            dur = float(np.mean(rows[:, 1 + p] - rows[:, 0])) * 1e-3          # microseconds
            fixed = ITER_FIXED_US[self.iter_kinds[i]]
            self.iter_work[i] = max(dur - fixed, 0.05 * dur, 0.05) * rows.shape[0]
        pinned = {i: self.iter_plan[i] for i in live if self.iter_kinds[i] == "fold"}
        self._replan(blocks, pinned)

    def _iterate_persistent(self, n, lr, sgd):
        while (self.iter_plan and self._iter_tuned < ITER_TUNE_ROUNDS and len(self.iter_plan) > 1 and n > 0
               and self._iter_adapt):
            # the first launches measure: two iterations with per-block timestamps, then the grid is
            # re-split so that the groups end together (the iterations themselves are ordinary ones;
            # twice, because a group's speed depends on what shares its SMs)
            self._iter_tuned += 1
            k = min(int(n), 2)
            self._launch_persistent(k, lr, sgd, trace=True)
            torch.cuda.synchronize(self.device)
            self._retune(self.iter_trace.cpu().numpy(), k)
            n -= k
        if n <= 0:
            return
        self._launch_persistent(n, lr, sgd, trace=self._iter_trace_env)

    def _launch_persistent(self, n, lr, sgd, trace=False):
        key = (float(lr), bool(sgd), self.b1, self.b2, self.eps, self.var_threshold)
        o = self._optim_cache.get(key)
        if o is None:
            o = _cabi.LhviOptim()
            o.n_vars, o.n_owned = self.n_vars, self.n_owned
            o.var_kind, o.var_dim, o.var_off = self.var_kind.data_ptr(), self.var_dim.data_ptr(), self.var_off.data_ptr()
            o.tau, o.mom1, o.mom2 = self.tau.data_ptr(), self.mom1.data_ptr(), self.mom2.data_ptr()
            o.wstate, o.step, o.sm_count = self.wstate.data_ptr(), self.step.data_ptr(), self.sm_count.data_ptr()
            o.lr, o.b1, o.b2, o.eps, o.var_threshold = float(lr), self.b1, self.b2, self.eps, self.var_threshold
            o.sgd = int(bool(sgd))
            o.accum = self.iter_accum.data_ptr()
            self._optim_cache[key] = o
        o.trace = None
        self.iter_trace = None
        if trace:                  # per-block timestamps of the launch (_retune, tools/iter_trace.py)
            self.iter_trace = torch.zeros(int(n) * 2 * 160 * 16, dtype=torch.int64, device=self.device)
            o.trace = self.iter_trace.data_ptr()
        x = C.byref(self.peer.desc) if self.plan.active else None
        rc = self.lib.lhvi_iterate(C.byref(self.desc), self.group_table, len(self.groups), x, C.byref(o), int(n),
                                   self._stream())
        if rc == 1:
            raise _cabi.LhviError("lhvi_iterate refused a model lhvi_iterate_supported accepted")
        _cabi.check(rc, self.lib)
        self.launch_count += 1
        self.launches_per_pass = 0

    def _graph_for(self, lr, sgd):
        """CUDA graph of one iteration for these hyper-parameters (captured once).  The step
        counter and bias corrections live on the device, so the same graph serves every t."""
        key = (float(lr), bool(sgd), self.b1, self.b2, self.eps, self.var_threshold)
        graph = self._graphs.get(key)
        if graph is None:
            # lazy kernel set-up (occupancy queries, module load) must happen outside capture;
            # this pass has no side effects on the parameters
            self.grad_pass()
            self.grad.zero_()
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            # Python's cyclic collector must not run inside the capture: a dead cycle that holds
            # another engine (its CUDA graphs, peer buffers) would be finalised there, and
            # cudaGraphExecDestroy / cudaFree under capture invalidate it ("operation not permitted
            # when stream is capturing"; seen intermittently when the previous model's drop-in
            # object -- compressed graphs are full of cycles -- died mid-capture).  torch's context
            # manager only collects when torch.compiler.config.force_cudagraph_gc is set.
            gc.collect()
            gc_was_enabled = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    self._iteration(lr, sgd)
            finally:
                if gc_was_enabled:
                    gc.enable()
            # capture records the launches without running them: nothing has changed yet
            self._graphs[key] = graph
        return graph

    def synchronize(self):
        """Wait for everything this engine has launched (timing probes)."""
        torch.cuda.synchronize(self.device)

    def iterate(self, n, lr, sgd=False):
        """``n`` Jacobi iterations: all gradients at the old parameters, then the step."""
        n = int(n)
        if n <= 0:
            return
        # One launch per call only pays off when the call holds several iterations: a cooperative launch
        # costs tens of microseconds more on the host than replaying a captured graph, which a caller who
        # synchronises after every single iteration (bench.py's e2e loop, is_log=True) would pay each time
        # (measured at 1.4 M records: 210 us per synchronised step against 133 us)
        if self.profile_group is None and n >= self.persistent_min_iters and self.persistent():
            if not self._grad_clean:
                self.grad[:self.n_param].zero_()
                self._grad_clean = True
            self._iterate_persistent(n, lr, sgd)
            return
        graph = None
        if self.use_graph and self.profile_group is None:
            # a failed capture is an error, not a reason to fall back silently (use_graph = False
            # asks for eager launches explicitly)
            graph = self._graph_for(lr, sgd)
        if not self._grad_clean:
            self.grad[:self.n_param].zero_()
            self._grad_clean = True
        for _ in range(n):
            if graph is not None:
                graph.replay()
            else:
                self._iteration(lr, sgd)
        self.launch_count += n * self.launches_per_iteration

    @property
    def launches_per_iteration(self):
        """Kernels of liblhvi.so per iteration on the per-group path: the group launches, lhvi_finish,
        lhvi_param_step (the persistent path is one launch per ``iterate`` call: ``launch_count``)."""
        return self.launches_per_pass + 1

    def last_free_energy(self):
        """Free energy computed by the most recent pass (at the parameters before its step)."""
        return float(self.grad[self.n_param + self.K].item())

    def free_energy(self):
        self.grad_pass()
        return self.last_free_energy()

    def gradients(self):
        """Host copies of (raw parameter gradients [n_param], raw G_w [K], free energy)."""
        g = self.grad_pass().double().cpu().numpy()
        return g[:self.n_param].copy(), g[self.n_param:self.n_param + self.K].copy(), float(g[-1])

    def mixture_belief(self, q_off, q_dim, q_kind, x):
        """Batched ``belief(x, rv)`` (VarInference.py:333-353); discrete x are state indices."""
        n = len(q_off)
        out = torch.empty(n, dtype=self.tdtype, device=self.device)
        if n == 0:
            return out
        off = self._dev(np.asarray(q_off, dtype=np.int32))
        dim = self._dev(np.asarray(q_dim, dtype=np.int32))
        kind = self._dev(np.asarray(q_kind, dtype=np.uint8))
        xs = self._dev(np.asarray(x, dtype=np.float64), self.tdtype)
        eta = self.plan.merge(self.eta)      # sharded records: a rank steps only the variables it owns
        _cabi.check(self.lib.lhvi_mixture_belief(
            self.dcode, self.K, n, off.data_ptr(), dim.data_ptr(), kind.data_ptr(), xs.data_ptr(),
            eta.data_ptr(), self.w.data_ptr(), out.data_ptr(), self._stream()), self.lib)
        return out

    def mixture_map(self, q_off=None, q_dim=None, q_kind=None):
        """Batched ``map`` (VarInference.py:355-376) of the given slots (default: every hidden
        variable of the model, in slot-table order): continuous -> MAP position, discrete ->
        arg-max state index.  With sharded records the state is merged first."""
        m = self.full_model
        if q_off is None:
            q_off, q_dim, q_kind = m.var_off, m.var_dim, m.var_kind
        n = len(q_off)
        out = torch.empty(n, dtype=self.tdtype, device=self.device)
        if n == 0:
            return out
        off = self._dev(np.asarray(q_off, dtype=np.int32))
        dim = self._dev(np.asarray(q_dim, dtype=np.int32))
        kind = self._dev(np.asarray(q_kind, dtype=np.uint8))
        eta = self.plan.merge(self.eta)
        _cabi.check(self.lib.lhvi_mixture_map(
            self.dcode, self.K, n, off.data_ptr(), dim.data_ptr(), kind.data_ptr(),
            eta.data_ptr(), self.w.data_ptr(), out.data_ptr(), self._stream()), self.lib)
        return out
