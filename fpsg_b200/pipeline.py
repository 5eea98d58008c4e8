"""Host-buffer streaming Chamfer step (include/fpsg_b200.h, fpsg_chamfer_pipeline_*): the shape of the reference's
training loop around the loss -- ``to_cuda(sample)``, ``model.loss(sample)``, ``ttl_loss.backward()``, ``.item()``
(reference src/trainNetwork.py:141-148) -- with the H2D copies of the next batch overlapped with the kernels of
the current one, and (every slot being a lane with its own stream and workspace) the latency-bound tail of a batch
overlapped with the distance sweep of the next. The runtime (streams, events, slots) is native code in libfpsg_b200.so; this class owns the two arenas
(a device tensor and a pinned host tensor) and turns results into tensors."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class ChamferPipeline:
    """``pipe = ChamferPipeline(B, N, M, device=..., depth=4)``; per batch ``pipe.submit(p1_cpu, p2_cpu)`` then, possibly
    after further submits, ``loss_sum, loss, gp1 = pipe.result()``: the sum over the batch of
    ``w1*mean(dist1) + w2*mean(dist2)`` (a Python float), the per-pair losses (CPU tensor (B,)), and the gradient of the
    sum w.r.t. p1 as a CUDA tensor (B,N,3) that aliases the pipeline's arena (valid until ``depth`` further submits)."""

    def __init__(self, B: int, N: int, M: int, device=None, w1: float = 1.0, w2: float = 1.0, depth: int = 4):
        L = _lib.lib()
        self.B, self.N, self.M, self.depth = int(B), int(N), int(M), int(depth)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ChamferPipeline needs a CUDA device (fpsg_b200 has no CPU path)")
        dbytes, pbytes = ctypes.c_size_t(), ctypes.c_size_t()
        _lib.check(L.fpsg_chamfer_pipeline_arena_bytes(self.B, self.N, self.M, self.depth, ctypes.byref(dbytes), ctypes.byref(pbytes)))
        self._arena = torch.empty(dbytes.value, dtype=torch.uint8, device=self.device)
        self._pinned = torch.empty(pbytes.value, dtype=torch.uint8).pin_memory()
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.fpsg_chamfer_pipeline_create(self.B, self.N, self.M, float(w1), float(w2), self.depth, self._arena.data_ptr(), dbytes.value,
                                                      self._pinned.data_ptr(), pbytes.value, ctypes.byref(self._h)))
        self._keep = []   # host tensors of the batches in flight
        self._loss = torch.empty(self.B, dtype=torch.float32)
        self._sum = ctypes.c_float()
        self._gp = ctypes.c_void_p()
        self._base = self._arena.data_ptr()
        self._gp_views = {}

    def set_exchange(self, exchange) -> None:
        """Multi-GPU: every batch takes part in the in-kernel scalar all-reduce (``fpsg_b200.sharded.PeerScalarExchange``,
        lag = depth <= 4). All ranks must submit the same number of batches. Call before the first submit."""
        tables = (ctypes.c_void_p * len(exchange.ptrs))(*exchange.ptrs)
        _lib.check(_lib.lib().fpsg_chamfer_pipeline_set_exchange(self._h, tables, exchange.rank, exchange.world))
        self._exchange = exchange

    def previous_global_sum(self) -> float:
        """After ``result()`` of batch n (n > depth): the loss sum of batch n - depth over ALL ranks."""
        v = ctypes.c_float()
        _lib.check(_lib.lib().fpsg_chamfer_pipeline_previous_global_sum(self._h, ctypes.byref(v)))
        return v.value

    def submit(self, p1_host: torch.Tensor, p2_host: torch.Tensor) -> None:
        for name, t, n in (("p1", p1_host, self.N), ("p2", p2_host, self.M)):
            if t.device.type != "cpu" or t.dtype != torch.float32 or tuple(t.shape) != (self.B, n, 3) or not t.is_contiguous():
                raise RuntimeError(f"{name} must be a contiguous float32 CPU tensor of shape ({self.B}, {n}, 3), got {t.dtype} {tuple(t.shape)} on {t.device}")
        _lib.check(_lib.lib().fpsg_chamfer_pipeline_submit(self._h, p1_host.data_ptr(), p2_host.data_ptr()))
        self._keep.append((p1_host, p2_host))

    def submit_device(self, p1: torch.Tensor, p2: torch.Tensor) -> None:
        """The same step for tensors that are already on the device (e.g. the decoder's output): no copies; the lane's kernels
        are ordered behind the current stream. ``p1`` / ``p2`` must not be modified until the batch's ``result()``."""
        for name, t, n in (("p1", p1, self.N), ("p2", p2, self.M)):
            if t.device != self.device or t.dtype != torch.float32 or tuple(t.shape) != (self.B, n, 3) or not t.is_contiguous():
                raise RuntimeError(f"{name} must be a contiguous float32 tensor of shape ({self.B}, {n}, 3) on {self.device}, "
                                   f"got {t.dtype} {tuple(t.shape)} on {t.device}")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().fpsg_chamfer_pipeline_submit_device(self._h, p1.data_ptr(), p2.data_ptr(), stream))
        self._keep.append((p1, p2))

    def result(self):
        _lib.check(_lib.lib().fpsg_chamfer_pipeline_result(self._h, ctypes.byref(self._sum), self._loss.data_ptr(), ctypes.byref(self._gp), None, None))
        self._keep.pop(0)
        off = self._gp.value - self._base
        gp1 = self._gp_views.get(off)
        if gp1 is None:  # one view per slot, built once: the call is on the per-step host path
            gp1 = self._gp_views[off] = self._arena[off:off + self.B * self.N * 12].view(torch.float32).view(self.B, self.N, 3)
        return self._sum.value, self._loss, gp1

    @property
    def in_flight(self) -> int:
        return len(self._keep)

    def close(self) -> None:
        if self._h:
            _lib.lib().fpsg_chamfer_pipeline_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
