"""Multi-GPU plan: shard the factor records, replicate the parameters, one all-reduce of the
flat gradient vector per iteration (SURVEY section 8 e).

Every gradient and the free energy are sums over records, so each rank processes a contiguous
1/world slice of every record group and the ranks exchange ``[parameter grads | G_w | energy]``
with a single ``all_reduce(sum)``; all ranks then apply the identical parameter step, keeping
the replicas bit-identical.  The same object drives NCCL on GPUs and gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class ShardPlan:
    def __init__(self, process_group=None, enabled=True):
        self.group = process_group
        self.world, self.rank = 1, 0
        if enabled and dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(process_group)
            self.rank = dist.get_rank(process_group)

    @property
    def active(self) -> bool:
        return self.world > 1

    def shard(self, model):
        return model.shard(self.rank, self.world) if self.active else model

    def all_reduce(self, flat: torch.Tensor) -> torch.Tensor:
        if self.active:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat
