"""Multi-GPU plan: owner-computes partition of the factor records (SURVEY section 8 e).

Every gradient and the free energy are sums over factor records, so the records can be split
over the ranks freely.  What decides the cost of a split is which *variables* end up being
touched by more than one rank, because only their gradients have to cross the NVLink:

* ``partition`` gives every non-hub variable a home rank (contiguous blocks of variables,
  balanced by record incidences), sends each record to the home of its first non-hub argument
  (records that touch only hubs are dealt out evenly), and then classifies every variable as
  *owned* (all its records live on one rank) or *shared* (hubs, block boundaries);
* per iteration the ranks exchange only ``[G_w | energy | gradients of shared variables]``
  -- a few hundred bytes for a relational model with a handful of group variables, one row of
  boundary variables per rank for a grid -- inside ``lhvi_finish`` over peer memory
  (``PeerExchange``), or through one NCCL / gloo all-reduce of that compact vector;
* each rank runs the optimiser step for the variables it owns plus the shared ones (whose
  summed gradients are bit-identical on every rank), so no parameter ever crosses the link
  during the loop; ``merge`` assembles the full vector when the host asks for the state.

The same ``ShardPlan`` object drives NCCL on GPUs and gloo in the CPU tests.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


@dataclass
class Partition:
    world: int
    rec_rank: list            # per group: int16 [n] rank of every record
    var_owner: np.ndarray     # int32 [V]: owning rank, -1 for shared variables
    hub: np.ndarray           # bool  [V]: excluded from anchoring (degree above the hub threshold)

    @property
    def shared(self) -> np.ndarray:
        return np.flatnonzero(self.var_owner < 0)


def _var_of_slot(model):
    lut = np.full(int(model.n_param) + 1, -1, dtype=np.int64)
    lut[model.var_off.astype(np.int64)] = np.arange(model.n_vars)
    return lut


def partition(model, world: int) -> Partition:
    """Assign every record of ``model`` to one of ``world`` ranks (see the module docstring)."""
    V = model.n_vars
    lut = _var_of_slot(model)
    factor_groups = [g for g in model.groups if not g.node]
    n_records = sum(g.n for g in factor_groups)

    degree = np.zeros(V, dtype=np.int64)
    for g in factor_groups:
        for a in range(g.nh):
            degree += np.bincount(lut[g.poff[a]], minlength=V)
    hub = degree > max(32, n_records // (8 * world))

    # home rank of every non-hub variable: contiguous blocks with equal incidence counts
    work = np.where(hub, 0, degree + 1)
    cum = np.cumsum(work) - work
    total = max(1, int(work.sum()))
    home = np.minimum(world - 1, cum * world // total).astype(np.int32)

    rec_rank = {}
    for gi, g in enumerate(model.groups):
        if g.node:
            continue
        rank = np.full(g.n, -1, dtype=np.int32)
        for a in range(g.nh):
            v = lut[g.poff[a]]
            cand = np.where(hub[v], -1, home[v])
            rank = np.where(rank < 0, cand, rank)
        loose = np.flatnonzero(rank < 0)              # records that touch hubs (or nothing) only
        if loose.size:
            rank[loose] = (np.arange(loose.size, dtype=np.int64) * world // loose.size).astype(np.int32)
        rec_rank[gi] = rank

    # which ranks touch each variable
    touched_by = np.zeros(V, dtype=np.int32)
    first = np.full(V, -1, dtype=np.int32)
    for r in range(world):
        seen = np.zeros(V, dtype=bool)
        for gi, g in enumerate(model.groups):
            if g.node:
                continue
            sel = rec_rank[gi] == r
            for a in range(g.nh):
                seen[lut[g.poff[a][sel]]] = True
        touched_by += seen
        first = np.where((first < 0) & seen, r, first)
    owner = np.where(touched_by > 1, -1, np.where(touched_by == 1, first, home)).astype(np.int32)

    # node records follow their variable (a shared variable's node record goes to the first
    # rank that touches it; records without a hidden variable are dealt out evenly)
    out = []
    for gi, g in enumerate(model.groups):
        if not g.node:
            out.append(rec_rank[gi].astype(np.int16))
        elif g.nh == 1:
            v = lut[g.poff[0]]
            out.append(np.where(owner[v] >= 0, owner[v], first[v]).astype(np.int16))
        else:
            out.append((np.arange(g.n, dtype=np.int64) * world // max(1, g.n)).astype(np.int16))
    return Partition(world, out, owner, hub)


def local_model(model, part: Partition, rank: int):
    """The records of ``rank`` and the variables it steps (owned + shared), in the global
    parameter layout (offsets are unchanged, so every rank addresses the same flat vector)."""
    groups = []
    for g, rr in zip(model.groups, part.rec_rank):
        sel = np.flatnonzero(rr == rank)
        if sel.size:
            groups.append(g.take(sel))
    keep = np.flatnonzero((part.var_owner == rank) | (part.var_owner < 0))
    return dataclasses.replace(model, var_kind=model.var_kind[keep], var_dim=model.var_dim[keep],
                               var_off=model.var_off[keep], groups=groups, handles=[], index={})


def slot_elements(model, variables):
    """Element offsets (into the flat parameter vector) of the K x dim blocks of ``variables``."""
    if len(variables) == 0:
        return np.zeros(0, dtype=np.int64)
    off = model.var_off[variables].astype(np.int64)
    size = model.K * model.var_dim[variables].astype(np.int64)
    within = np.arange(int(size.sum()), dtype=np.int64) - np.repeat(np.cumsum(size) - size, size)
    return np.repeat(off, size) + within


class ShardPlan:
    def __init__(self, process_group=None, enabled=True):
        self.group = process_group
        self.world, self.rank = 1, 0
        if enabled and dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(process_group)
            self.rank = dist.get_rank(process_group)
        self.part = None
        self.full = None
        self.shared_idx = np.zeros(0, dtype=np.int64)     # element offsets of shared variables
        self.keep_idx = None                              # elements this rank contributes to merge()

    @property
    def active(self) -> bool:
        return self.world > 1

    def shard(self, model):
        """Rank-local model.  Deterministic: every rank computes the same partition."""
        if not self.active:
            return model
        self.full = model
        self.part = partition(model, self.world)
        self.shared_idx = slot_elements(model, self.part.shared)
        mine = np.flatnonzero(self.part.var_owner == self.rank)
        keep = slot_elements(model, mine)
        if self.rank == 0:
            keep = np.concatenate([keep, self.shared_idx])
        self.keep_idx = keep
        return local_model(model, self.part, self.rank)

    def describe(self):
        if not self.active:
            return "1 rank"
        p = self.part
        return (f"{self.world} ranks, owner-computes: {int((p.var_owner >= 0).sum())} owned + "
                f"{p.shared.size} shared variables ({int(p.hub.sum())} hubs), "
                f"{self.shared_idx.size} exchanged gradient elements")

    # ---- collectives ----------------------------------------------------------------------
    def all_reduce(self, flat: torch.Tensor) -> torch.Tensor:
        """Dense sum over ranks (queries outside the loop, and the compact exchange vector)."""
        if self.active:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat

    def merge(self, flat: torch.Tensor) -> torch.Tensor:
        """Full parameter-layout vector from the ranks' pieces: owned slots from their owner,
        shared slots (identical everywhere) from rank 0."""
        if not self.active:
            return flat
        out = torch.zeros_like(flat)
        idx = torch.as_tensor(self.keep_idx, device=flat.device)
        out[idx] = flat[idx]
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out


class PeerExchange:
    """Peer-memory buffers behind ``lhvi_finish``'s in-kernel exchange (``include/lhvi.h``,
    ``lhvi_exchange``): every rank allocates one receive buffer + flag array, publishes its
    CUDA IPC handle through the process group, and maps the other ranks' buffers."""

    ELEMS_PER_BLOCK = 4096

    def __init__(self):
        self.desc = None
        self.local = None
        self.opened = []
        self.lib = None

    @classmethod
    def create(cls, plan, lib, n_idx, K, idx_ptr, tdtype, device):
        """Returns a ready exchange, or ``None`` if any rank could not map its peers (then
        every rank falls back to the collective together)."""
        import ctypes as C

        from . import _cabi
        self = cls()
        self.lib = lib
        world, rank = plan.world, plan.rank
        n_x = n_idx + K + 1
        itemsize = torch.empty(0, dtype=tdtype).element_size()
        blocks = max(1, min(32, -(-n_x // cls.ELEMS_PER_BLOCK)))
        recv_bytes = (2 * world * n_x * itemsize + 255) // 256 * 256
        flag_bytes = world * blocks * 8
        ok = world <= _cabi.LHVI_MAX_PEERS
        handle = C.create_string_buffer(_cabi.LHVI_IPC_HANDLE_BYTES)
        ptr = C.c_void_p()
        if ok:
            ok = lib.lhvi_peer_alloc(recv_bytes + flag_bytes, C.byref(ptr), handle) == 0
            if ok:
                self.local = ptr.value
        handles = [None] * world
        dist.all_gather_object(handles, (bool(ok), handle.raw, int(torch.cuda.current_device())),
                               group=plan.group)
        ok = all(h[0] for h in handles)
        bases = [None] * world
        if ok:
            for p, (_, raw, _) in enumerate(handles):
                if p == rank:
                    bases[p] = self.local
                    continue
                q = C.c_void_p()
                if lib.lhvi_peer_open(raw, C.byref(q)) != 0:
                    ok = False
                    break
                self.opened.append(q.value)
                bases[p] = q.value
        flag = torch.tensor([1 if ok else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=plan.group)
        if int(flag.item()) == 0:
            self.close()
            return None
        self.seq = torch.zeros(blocks, dtype=torch.int64, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        d = _cabi.LhviExchange()
        d.world, d.rank, d.blocks, d.n_idx = world, rank, blocks, n_idx
        d.idx = idx_ptr if n_idx else None
        for p in range(world):
            d.recv[p] = bases[p]
            d.flags[p] = bases[p] + recv_bytes
        d.seq, d.status = self.seq.data_ptr(), self.status.data_ptr()
        self.desc = d
        torch.cuda.synchronize(device)
        dist.barrier(group=plan.group)       # every buffer is mapped and zeroed before first use
        return self

    def timed_out(self) -> bool:
        return bool(int(self.status.item()) != 0)

    def close(self):
        if self.lib is None:
            return
        for q in self.opened:
            self.lib.lhvi_peer_close(q)
        self.opened = []
        if self.local:
            self.lib.lhvi_peer_free(self.local)
            self.local = None
