"""Batch-sharded Chamfer loss for one process per GPU.

Every cloud pair is independent, so the batch is split contiguously over the ranks and each rank runs the
single-GPU path on its shard; point gradients stay rank-local. The only exchange the path has is the scalar the
reference logs and back-propagates from -- ``pc_metric(syn, ref).sum()`` (src/models/few_shot.py:110,117) -- which
is all-reduced (SUM). Two implementations:

* ``ShardedChamferLoss``  -- autograd path; the sum is all-reduced over NCCL, launched asynchronously right after the
  forward and waited for only when the value is read, so it overlaps the backward kernel.
* ``ShardedChamferStep``  -- the fused loss step (forward + backward in one launch chain) with the all-reduce done
  INSIDE the loss kernel's epilogue through NVLink peer memory (``fpsg_chamfer_multi_loss_fwd_bwd_dist``,
  include/fpsg_b200.h): every rank's table is mapped into every process with torch's symmetric memory, the epilogue
  writes the rank's partial sum into all tables (one 8-byte store per peer) and adds up what the others wrote. No
  host-launched collective, no extra kernel: one collective per step costs the NVLink write latency.
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


class PeerScalarExchange:
    """The peer-mapped tables of the in-kernel scalar all-reduce (``fpsg_peer_exchange_t``). Collective constructor:
    every rank of ``group`` must create it at the same point."""

    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self._lib = _lib
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f"at most {_lib.MAX_PEERS} ranks (one NVLink domain), got {self.world}")
        self.table = symm_mem.empty(_lib.PEER_RING * _lib.MAX_PEERS, dtype=torch.int64, device=device)
        self.table.zero_()
        self._hdl = symm_mem.rendezvous(self.table, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every table is zero before any kernel writes into it
        self.global_out = torch.zeros(_lib.PEER_RING, dtype=torch.float32, device=device)
        self.seq = 0

    def next(self, lag: int = 0):
        """The descriptor of the next step (sequence numbers advance identically on all ranks)."""
        self.seq += 1
        px = self._lib.PeerExchange()
        for r in range(self.world):
            px.table[r] = self.ptrs[r]
        px.global_out = self.global_out.data_ptr()
        px.rank, px.world, px.seq, px.lag = self.rank, self.world, self.seq, int(lag)
        return px

    def collect(self, seq: int, stream: int) -> None:
        """Enqueue the collection of step ``seq`` (the flush after the last step of a lag-1 sequence)."""
        px = self._lib.PeerExchange()
        for r in range(self.world):
            px.table[r] = self.ptrs[r]
        px.global_out = self.global_out.data_ptr()
        px.rank, px.world, px.seq, px.lag = self.rank, self.world, int(seq), 0
        self._lib.check(self._lib.lib().fpsg_peer_exchange_collect(px, stream))

    def global_sum(self, seq: int) -> torch.Tensor:
        """0-d view of the reduced value of step ``seq`` (valid once the step that collected it has run; the ring keeps
        the last ``PEER_RING`` steps)."""
        return self.global_out[seq % self._lib.PEER_RING]


class ShardedChamferStep:
    """One training-loss step on this rank's shard with the scalar all-reduce inside the kernel.

    ``step = ShardedChamferStep(B_local, N, M, device)``; ``loss, local_sum, gp1, seq = step(p1, p2)`` launches sweep +
    finish (per-pair losses, their local sum, the gradient w.r.t. p1 -- the upstream gradient is the constant 1 of
    ``.sum()`` / ``ttl_loss.backward()``, reference src/models/few_shot.py:110, src/trainNetwork.py:144) and the exchange;
    ``step.global_sum(seq)`` is the all-reduced sum of step ``seq`` (with ``lag=L`` it is produced L steps later; L <= PEER_RING/4)."""

    def __init__(self, B: int, N: int, M: int, device: torch.device, group: Optional[dist.ProcessGroup] = None, lag: int = 0):
        from . import _lib
        self._lib = _lib
        self.B, self.N, self.M, self.device, self.lag = B, N, M, device, lag
        self.exchange = PeerScalarExchange(device, group)
        L = _lib.lib()
        self.nbytes = int(L.fpsg_chamfer_fwd_workspace_bytes(B, N, M))
        self.ws = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=device)
        self.idx = torch.empty(B * (N + M), dtype=torch.int64, device=device)
        self.out = torch.empty(2 + B, dtype=torch.float32, device=device)  # totals[2], loss[B]
        self.gp1 = torch.empty(B, N, 3, dtype=torch.float32, device=device)
        self.seg = (_lib.ChamferSegment * 1)()

    def __call__(self, p1: torch.Tensor, p2: torch.Tensor, w1: float = 1.0, w2: float = 1.0):
        _lib = self._lib
        if not (p1.is_cuda and p2.is_cuda and p1.dtype == torch.float32 and p2.dtype == torch.float32 and p1.is_contiguous() and p2.is_contiguous()
                and tuple(p1.shape) == (self.B, self.N, 3) and tuple(p2.shape) == (self.B, self.M, 3)):
            raise RuntimeError(f"expected contiguous float32 CUDA tensors of shape ({self.B},{self.N},3) and ({self.B},{self.M},3)")
        sg = self.seg
        sg[0].p1, sg[0].p2, sg[0].gp1, sg[0].gp2, sg[0].batch, sg[0].weight = p1.data_ptr(), p2.data_ptr(), self.gp1.data_ptr(), None, self.B, 1.0
        px = self.exchange.next(self.lag)
        ip = self.idx.data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().fpsg_chamfer_multi_loss_fwd_bwd_dist(sg, 1, self.N, self.M, float(w1), float(w2), None, self.out.data_ptr() + 8,
                                                                   self.out.data_ptr(), ip, ip + 8 * self.B * self.N, self.ws.data_ptr(), self.nbytes,
                                                                   px, stream))
        return self.out[2:], self.out[0], self.gp1, px.seq

    def flush(self) -> int:
        """lag >= 1: collect the sums of the last ``lag`` steps (nothing follows them that would). Returns the sequence
        number of the last step."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for s in range(max(1, self.exchange.seq - max(self.lag, 1) + 1), self.exchange.seq + 1):
            self.exchange.collect(s, stream)
        return self.exchange.seq

    def global_sum(self, seq: int) -> torch.Tensor:
        return self.exchange.global_sum(seq)
