"""``get_graph_feature(x, k=20, idx=None)`` of the DGCNN encoder on B200 (reference src/dgcnn/model.py:23-42), same
signature and result: x (B,C,N) float32 CUDA -> (B, 2C, N, k) with ``cat(neighbour - centre, centre)`` along the
channel axis. ``idx`` defaults to ``fpsg_b200.knn(x, k)``. Differentiable w.r.t. x (the reference's autograd path
through the flat gather, ``repeat``, ``cat`` and ``permute``); idx carries no gradient."""
from __future__ import annotations

import torch

from . import _lib
from .knn import knn


from .chamfer import _stream_ptr as _stream

# one call of the library takes at most 65535 clouds (a grid dimension); the reference takes any batch size: larger batches go in chunks
_MAX_CLOUDS_PER_CALL = 65535


class _GraphFeature(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        B, C, N = x.shape
        k = idx.shape[2]
        out = torch.empty((B, 2 * C, N, k), dtype=torch.float32, device=x.device)
        if B > 0:
            with torch.cuda.device(x.device):
                _lib.check(_lib.lib().fpsg_edge_feature_fwd(x.data_ptr(), idx.data_ptr(), B, C, N, k, out.data_ptr(), _stream(x.device)))
        ctx.save_for_backward(idx)
        ctx.shape = (B, C, N, k)
        return out

    @staticmethod
    def backward(ctx, gout):
        (idx,) = ctx.saved_tensors
        B, C, N, k = ctx.shape
        gx = torch.empty((B, C, N), dtype=torch.float32, device=gout.device)
        if B > 0:
            gout = gout.to(torch.float32).contiguous()
            L = _lib.lib()
            with torch.cuda.device(gout.device):
                nbytes = L.fpsg_edge_feature_bwd_workspace_bytes(B, C, N, k)
                ws = torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=gout.device)
                _lib.check(L.fpsg_edge_feature_bwd(gout.data_ptr(), idx.data_ptr(), B, C, N, k, gx.data_ptr(), ws.data_ptr(), nbytes,
                                                   _stream(gout.device)))
        return gx, None


def get_graph_feature(x: torch.Tensor, k: int = 20, idx: torch.Tensor | None = None) -> torch.Tensor:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("x must be a CUDA tensor (fpsg_b200 has no CPU fallback)")
    if x.dtype != torch.float32:
        raise RuntimeError(f"x must be float32, got {x.dtype}")
    B, C = x.shape[0], x.shape[1]
    x = x.reshape(B, C, -1).contiguous()  # model.py:24-25 (`x.view(batch_size, -1, num_points)`)
    N = x.shape[2]
    if idx is None:
        idx = knn(x, k=k)
    else:
        if idx.dtype != torch.int64 or idx.dim() != 3 or idx.shape[0] != B or idx.shape[1] != N:
            raise RuntimeError(f"idx must be int64 of shape (B, N, k), got {idx.dtype} {tuple(idx.shape)}")
        idx = idx.to(x.device).contiguous()
        # a caller-supplied index is range-checked (the reference's gather raises an indexing error; the kernels do not check):
        # one host synchronisation, on this path only -- DGCNNfeat.forward never passes idx
        if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= N):
            raise IndexError(f"idx values must lie in [0, {N}), got [{int(idx.min())}, {int(idx.max())}]")
    if B > _MAX_CLOUDS_PER_CALL:
        m = _MAX_CLOUDS_PER_CALL
        return torch.cat([_GraphFeature.apply(x[s:s + m], idx[s:s + m]) for s in range(0, B, m)], dim=0)
    return _GraphFeature.apply(x, idx)
