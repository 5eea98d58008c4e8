"""``knn(x, k)`` of the DGCNN encoder on B200 (reference src/dgcnn/model.py:13-20), same signature and result
contract: x (B,C,N) float32 CUDA -> LongTensor (B,N,k), nearest first, self included. Non-differentiable, like the
reference's ``topk(...)[1]``."""
from __future__ import annotations

import torch

from . import _lib
from .chamfer import _device_guard, _stream_ptr


_MAX_CLOUDS_PER_CALL = 65535


def knn(x: torch.Tensor, k: int) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"x must be a torch.Tensor, got {type(x).__name__}")
    if not x.is_cuda:
        raise RuntimeError("x must be a CUDA tensor (fpsg_b200 has no CPU fallback)")
    if x.dtype != torch.float32:
        raise RuntimeError(f"x must be float32, got {x.dtype}")
    if x.dim() != 3:
        raise RuntimeError(f"x must have shape (B, C, N), got {tuple(x.shape)}")
    B, C, N = x.shape
    k = int(k)
    if not (1 <= k <= N):
        # torch.topk raises RuntimeError("selected index k out of range") here
        raise RuntimeError(f"selected index k out of range: k={k}, N={N}")
    if k > _lib.KNN_MAX_K:
        raise RuntimeError(f"k={k} exceeds the supported maximum {_lib.KNN_MAX_K}")
    x = x.detach().contiguous()
    idx = torch.empty((B, N, k), dtype=torch.int64, device=x.device)
    if B == 0:
        return idx
    L = _lib.lib()
    with _device_guard(x.device):
        # one call of the library takes at most 65535 clouds (a grid dimension); larger batches go in chunks, as the reference's
        # torch ops accept any batch size
        for b0 in range(0, B, _MAX_CLOUDS_PER_CALL):
            nb = min(_MAX_CLOUDS_PER_CALL, B - b0)
            nbytes = L.fpsg_knn_workspace_bytes(nb, C, N, k)
            ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=x.device)
            rc = L.fpsg_knn_fwd(x.data_ptr() + 4 * b0 * C * N, nb, C, N, k, idx.data_ptr() + 8 * b0 * N * k, ws.data_ptr(), nbytes,
                                _stream_ptr(x.device))
            _lib.check(rc)
    return idx
