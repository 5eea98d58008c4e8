"""Batch-sharded Chamfer loss for one process per GPU.

Every cloud pair is independent, so the batch is split contiguously over the ranks and each rank runs the
single-GPU path on its shard; point gradients stay rank-local. The only exchange the path has is the scalar the
reference logs and back-propagates from -- ``pc_metric(syn, ref).sum()`` (src/models/few_shot.py:110,117) -- which
is all-reduced (SUM) over NCCL/NVLink. The all-reduce is launched asynchronously right after the forward and
waited for only when the value is read, so it overlaps the backward kernel.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a batch for ``rank``; the first ``batch % world`` ranks take one extra pair."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    q, r = divmod(batch, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class PendingScalar:
    """A scalar whose all-reduce may still be in flight; ``.value()`` waits and returns the 0-d tensor."""

    def __init__(self, tensor: torch.Tensor, work=None):
        self._t = tensor
        self._work = work

    def value(self) -> torch.Tensor:
        if self._work is not None:
            self._work.wait()
            self._work = None
        return self._t

    def item(self) -> float:
        return float(self.value().item())


class ShardedChamferLoss:
    """loss_local, total = ShardedChamferLoss(group)(p1_shard, p2_shard)

    ``loss_local`` (B_local,) is differentiable w.r.t. the local shards; ``total`` is the global
    ``sum_b loss[b]`` over all ranks (a PendingScalar). ``loss_fn`` defaults to ``fpsg_b200.chamfer_loss``;
    CPU tests inject the oracle to exercise the sharding/reduction logic under gloo.
    """

    def __init__(self, group: Optional[dist.ProcessGroup] = None,
                 loss_fn: Optional[Callable[..., torch.Tensor]] = None):
        self.group = group
        if loss_fn is None:
            from .chamfer import chamfer_loss
            loss_fn = chamfer_loss
        self.loss_fn = loss_fn

    def __call__(self, p1: torch.Tensor, p2: torch.Tensor, w1: float = 1.0, w2: float = 1.0):
        loss_local = self.loss_fn(p1, p2, w1, w2)
        total = loss_local.detach().sum()
        work = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            work = dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return loss_local, PendingScalar(total, work)
