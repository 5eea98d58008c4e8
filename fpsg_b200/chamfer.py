"""Chamfer / sided distance on B200: the host-side mirror of ``kaolin.metrics.pointcloud`` as the reference
uses it (reference src/models/few_shot.py:13,57,110,117,167), calling the C ABI of libfpsg_b200.so.

Three public call surfaces, all ``torch.autograd`` aware:

* ``chamfer_distance(p1, p2) -> (dist1, dist2, idx1, idx2)``  -- the north-star primitive.
* ``chamfer_loss(p1, p2, w1=1., w2=1.) -> (B,)``               -- what Kaolin's ``chamfer_distance`` returns
  (``w1*mean(dist1) + w2*mean(dist2)``, squared L2); fused forward reduction and uniform-gradient backward.
* ``sided_distance(p1, p2) -> (dist, idx)``                    -- Kaolin's one-directional primitive.
* ``chamfer_loss_multi(pairs, weights) -> (total, sums, loss)`` -- several calls (the query-set and support-set
  losses of one training step, src/models/few_shot.py:109-122) and their weighted sum in one launch chain.

Error behaviour follows Kaolin's TORCH_CHECKs: anything but CUDA float32 tensors of shape (B,N,3)/(B,M,3) raises;
non-contiguous inputs are made contiguous. There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib


def _check_clouds(p1: torch.Tensor, p2: torch.Tensor):
    for name, t in (("p1", p1), ("p2", p2)):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor (fpsg_b200 has no CPU fallback)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name} must be float32, got {t.dtype}")
        if t.dim() != 3 or t.shape[2] != 3:
            raise RuntimeError(f"{name} must have shape (B, N, 3), got {tuple(t.shape)}")
    if p1.shape[0] != p2.shape[0]:
        raise RuntimeError(f"batch sizes differ: {p1.shape[0]} vs {p2.shape[0]}")
    if p1.device != p2.device:
        raise RuntimeError(f"p1 and p2 are on different devices: {p1.device} vs {p2.device}")
    if p1.shape[1] < 1 or p2.shape[1] < 1:
        raise RuntimeError(f"empty point cloud: N={p1.shape[1]}, M={p2.shape[1]}")
    return p1.contiguous(), p2.contiguous()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream_ptr(device) -> int:
    """The current stream's cudaStream_t for `device` (the raw accessor is ~10x cheaper than building a Stream)."""
    if _raw_stream is not None:
        return _raw_stream(device.index if device.index is not None else torch.cuda.current_device())
    return torch.cuda.current_stream(device).cuda_stream


class _device_guard:
    """torch.cuda.device(dev) only when dev is not already current (the common case costs nothing)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if torch.cuda.current_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


_WS_BYTES = {}  # (B, N, M) -> workspace size; the query is pure arithmetic, cache it
_DIAG = {"keep": False, "last": None}  # tests: keep the last workspace so that its diagnostic counter can be read


def _workspace(B: int, N: int, M: int, device) -> tuple[torch.Tensor, int]:
    key = (B, N, M, device.index)
    nbytes = _WS_BYTES.get(key)
    if nbytes is None:
        nbytes = int(_lib.lib().fpsg_chamfer_fwd_workspace_bytes(B, N, M))
        _WS_BYTES[key] = nbytes
    # a fresh caching-allocator block per call: safe under multi-stream use, and the allocator makes it cheap
    ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
    if _DIAG["keep"]:
        _DIAG["last"] = (ws, B, N, M)
    return ws, nbytes


def diag_keep_workspaces(on: bool = True) -> None:
    """Test hook: keep a reference to the workspace of the most recent Chamfer forward (see ``diag_mismatches``)."""
    _DIAG["keep"] = bool(on)
    _DIAG["last"] = None


def diag_mismatches() -> int:
    """The diagnostic counter of the most recent forward's workspace (include/fpsg_b200.h, fpsg_chamfer_ws_diag_offset):
    arg-min recoveries that found no candidate bit-equal to the merged minimum. Must be 0. Synchronises."""
    if _DIAG["last"] is None:
        raise RuntimeError("no workspace kept: call diag_keep_workspaces() before the forward")
    ws, B, N, M = _DIAG["last"]
    off = int(_lib.lib().fpsg_chamfer_ws_diag_offset(B, N, M))
    return int(ws[off:off + 4].view(torch.int32).item())


def _forward(p1: torch.Tensor, p2: torch.Tensor):
    """Runs K1 and returns (dist1, dist2, idx1, idx2)."""
    L = _lib.lib()
    B, N, _ = p1.shape
    M = p2.shape[1]
    dev = p1.device
    idx = torch.empty((B * (N + M),), dtype=torch.int64, device=dev)  # one allocation for both index outputs
    idx1 = idx[:B * N].view(B, N)
    idx2 = idx[B * N:].view(B, M)
    dist1 = torch.empty((B, N), dtype=torch.float32, device=dev)
    dist2 = torch.empty((B, M), dtype=torch.float32, device=dev)
    if B == 0:
        return dist1, dist2, idx1, idx2
    with _device_guard(dev):
        ws, nbytes = _workspace(B, N, M, dev)
        _lib.check(L.fpsg_chamfer_fwd(p1.data_ptr(), p2.data_ptr(), B, N, M, dist1.data_ptr(), dist2.data_ptr(),
                                      idx1.data_ptr(), idx2.data_ptr(), ws.data_ptr(), nbytes, _stream_ptr(dev)))
    return dist1, dist2, idx1, idx2


def _backward(p1, p2, idx1, idx2, g1, g2, need1: bool, need2: bool):
    L = _lib.lib()
    B, N, _ = p1.shape
    M = p2.shape[1]
    dev = p1.device
    gp1 = torch.empty_like(p1) if need1 else None
    gp2 = torch.empty_like(p2) if need2 else None
    if B == 0 or not (need1 or need2):
        return gp1, gp2
    g1 = g1.to(torch.float32).contiguous()
    g2 = g2.to(torch.float32).contiguous()
    with _device_guard(dev):
        rc = L.fpsg_chamfer_bwd(p1.data_ptr(), p2.data_ptr(), idx1.data_ptr(), idx2.data_ptr(), g1.data_ptr(),
                                g2.data_ptr(), B, N, M, gp1.data_ptr() if need1 else None,
                                gp2.data_ptr() if need2 else None, None, 0, _stream_ptr(dev))
        _lib.check(rc)
    return gp1, gp2


class _ChamferDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p1, p2):
        p1, p2 = _check_clouds(p1, p2)
        dist1, dist2, idx1, idx2 = _forward(p1, p2)
        ctx.save_for_backward(p1, p2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, g1, g2, _gi1, _gi2):
        p1, p2, idx1, idx2 = ctx.saved_tensors
        gp1, gp2 = _backward(p1, p2, idx1, idx2, g1, g2, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return gp1, gp2


class _SidedDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p1, p2):
        p1, p2 = _check_clouds(p1, p2)
        dist1, _dist2, idx1, idx2 = _forward(p1, p2)
        ctx.save_for_backward(p1, p2, idx1, idx2)
        ctx.mark_non_differentiable(idx1)
        return dist1, idx1

    @staticmethod
    def backward(ctx, g1, _gi1):
        p1, p2, idx1, idx2 = ctx.saved_tensors
        g2 = torch.zeros(idx2.shape, dtype=torch.float32, device=p1.device)  # the other direction carries no gradient
        gp1, gp2 = _backward(p1, p2, idx1, idx2, g1, g2, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return gp1, gp2


class _ChamferLoss(torch.autograd.Function):
    """The reference's call shape (``pc_metric(a, b).sum().backward()``): kept lean -- one index buffer, no views, the
    pointers of idx1 / idx2 are offsets into it -- because at B=32, N=2048 the device needs ~64 us and every host
    microsecond shows."""

    @staticmethod
    def forward(ctx, p1, p2, w1, w2):
        p1, p2 = _check_clouds(p1, p2)
        B, N = p1.shape[0], p1.shape[1]
        M = p2.shape[1]
        dev = p1.device
        idx = torch.empty((B * (N + M),), dtype=torch.int64, device=dev)  # idx1 (B,N) then idx2 (B,M)
        loss = torch.empty((B,), dtype=torch.float32, device=dev)
        if B > 0:
            with _device_guard(dev):
                ws, nbytes = _workspace(B, N, M, dev)
                ip = idx.data_ptr()
                _lib.check(_lib.lib().fpsg_chamfer_loss_fwd(p1.data_ptr(), p2.data_ptr(), B, N, M, w1, w2, loss.data_ptr(), None, None,
                                                            ip, ip + 8 * B * N, ws.data_ptr(), nbytes, _stream_ptr(dev)))
        ctx.save_for_backward(p1, p2, idx)
        ctx.w = (w1, w2)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        p1, p2, idx = ctx.saved_tensors
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        B, N = p1.shape[0], p1.shape[1]
        M = p2.shape[1]
        dev = p1.device
        gp1 = torch.empty_like(p1) if need1 else None
        gp2 = torch.empty_like(p2) if need2 else None
        if B > 0 and (need1 or need2):
            if gloss.dtype != torch.float32 or not gloss.is_contiguous():
                gloss = gloss.to(torch.float32).contiguous()
            with _device_guard(dev):
                ip = idx.data_ptr()
                _lib.check(_lib.lib().fpsg_chamfer_loss_bwd(p1.data_ptr(), p2.data_ptr(), ip, ip + 8 * B * N, gloss.data_ptr(),
                                                            ctx.w[0], ctx.w[1], B, N, M, gp1.data_ptr() if need1 else None,
                                                            gp2.data_ptr() if need2 else None, None, 0, _stream_ptr(dev)))
        return gp1, gp2, None, None


class _ChamferLossMulti(torch.autograd.Function):
    """forward(w1, w2, weights, p1_0, p2_0, p1_1, p2_1, ...) -> (total (), sums (nseg,), loss (Btot,))"""

    @staticmethod
    def forward(ctx, w1, w2, weights, *clouds):
        nseg = len(clouds) // 2
        pairs = [_check_clouds(clouds[2 * s], clouds[2 * s + 1]) for s in range(nseg)]
        N, M = pairs[0][0].shape[1], pairs[0][1].shape[1]
        dev = pairs[0][0].device
        for a, b in pairs:
            if a.shape[1] != N or b.shape[1] != M:
                raise RuntimeError(f"all segments must share N and M: got {(a.shape[1], b.shape[1])} vs {(N, M)}")
            if a.device != dev:
                raise RuntimeError("all segments must live on the same device")
        Btot = sum(a.shape[0] for a, _ in pairs)
        L = _lib.lib()
        segs = (_lib.ChamferSegment * nseg)()
        for s, (a, b) in enumerate(pairs):
            segs[s].p1, segs[s].p2, segs[s].gp1, segs[s].gp2 = a.data_ptr(), b.data_ptr(), None, None
            segs[s].batch, segs[s].weight = a.shape[0], float(weights[s])
        idx = torch.empty((Btot * (N + M),), dtype=torch.int64, device=dev)
        out = torch.empty((1 + nseg + Btot,), dtype=torch.float32, device=dev)  # totals, then the per-pair losses
        with _device_guard(dev):
            ws, nbytes = _workspace(Btot, N, M, dev) if Btot > 0 else (None, 0)
            _lib.check(L.fpsg_chamfer_multi_loss_fwd(segs, nseg, N, M, float(w1), float(w2), out[1 + nseg:].data_ptr(),
                                                     out.data_ptr(), idx.data_ptr(), idx[Btot * N:].data_ptr(),
                                                     ws.data_ptr() if ws is not None else None, nbytes, _stream_ptr(dev)))
        ctx.save_for_backward(idx, *[t for pr in pairs for t in pr])
        ctx.meta = (float(w1), float(w2), tuple(float(w) for w in weights), N, M, Btot)
        total, sums, loss = out[0], out[1:1 + nseg], out[1 + nseg:]
        ctx.mark_non_differentiable(sums, loss)
        return total, sums, loss

    @staticmethod
    def backward(ctx, gtotal, _gsums, _gloss):
        idx, *clouds = ctx.saved_tensors
        w1, w2, weights, N, M, Btot = ctx.meta
        nseg = len(clouds) // 2
        dev = idx.device
        L = _lib.lib()
        segs = (_lib.ChamferSegment * nseg)()
        grads = [None] * (2 * nseg)
        for s in range(nseg):
            a, b = clouds[2 * s], clouds[2 * s + 1]
            if ctx.needs_input_grad[3 + 2 * s]:
                grads[2 * s] = torch.empty_like(a)
            if ctx.needs_input_grad[4 + 2 * s]:
                grads[2 * s + 1] = torch.empty_like(b)
            segs[s].p1, segs[s].p2 = a.data_ptr(), b.data_ptr()
            segs[s].gp1 = grads[2 * s].data_ptr() if grads[2 * s] is not None else None
            segs[s].gp2 = grads[2 * s + 1].data_ptr() if grads[2 * s + 1] is not None else None
            segs[s].batch, segs[s].weight = a.shape[0], weights[s]
        if Btot > 0 and any(g is not None for g in grads):
            gtotal = gtotal.to(torch.float32).contiguous()
            with _device_guard(dev):
                _lib.check(L.fpsg_chamfer_multi_loss_bwd(segs, nseg, idx.data_ptr(), idx[Btot * N:].data_ptr(), gtotal.data_ptr(),
                                                         w1, w2, N, M, _stream_ptr(dev)))
        return (None, None, None, *grads)


def chamfer_loss_multi(pairs, weights=None, w1: float = 1.0, w2: float = 1.0):
    """Several Chamfer loss calls and the reduction the reference applies on top, in one launch chain.

    ``pairs`` = [(p1_0, p2_0), (p1_1, p2_1), ...] (each as for ``chamfer_loss``; all with the same N and M),
    ``weights`` = one factor per pair of tensors (default 1). Returns ``(total, sums, loss)``:
    ``sums[s] = chamfer_loss(p1_s, p2_s).sum()``, ``total = sum_s weights[s] * sums[s]`` (differentiable w.r.t. the
    clouds; ``sums`` and the per-pair ``loss`` (Btot,) are for logging and carry no gradient).

    This is src/models/few_shot.py:109-122 -- ``loss_rec_q = pc_metric(syn_q, ref_q).sum()``, ``loss_rec_s = ...``,
    ``query_factor * loss_rec_q + support_factor * loss_rec_s`` -- as one forward and one backward call."""
    pairs = list(pairs)
    if not 1 <= len(pairs) <= _lib.MAX_SEGMENTS:
        raise RuntimeError(f"need 1..{_lib.MAX_SEGMENTS} pairs of clouds, got {len(pairs)}")
    weights = [1.0] * len(pairs) if weights is None else [float(w) for w in weights]
    if len(weights) != len(pairs):
        raise RuntimeError("one weight per pair of clouds is required")
    flat = [t for pr in pairs for t in pr]
    return _ChamferLossMulti.apply(float(w1), float(w2), tuple(weights), *flat)


def chamfer_loss_and_grad(p1: torch.Tensor, p2: torch.Tensor, w1: float = 1.0, w2: float = 1.0, gloss: torch.Tensor | None = None,
                          need_gp1: bool = True, need_gp2: bool = False, return_indices: bool = False):
    """Forward AND backward of ``chamfer_loss(p1, p2, w1, w2)`` in one fused launch chain, for a loss whose upstream
    gradient is known up front -- the reference's training step, ``pc_metric(a, b).sum()`` followed by
    ``ttl_loss.backward()`` (src/models/few_shot.py:110,117, src/trainNetwork.py:144), where that gradient is a constant.

    ``gloss`` (B,) is the upstream gradient of the per-pair loss (``None`` = ones, the ``.sum()`` case). Returns
    ``(loss (B,), gp1 or None, gp2 or None)`` (plus ``idx1, idx2`` with ``return_indices``). Not recorded by autograd:
    feed ``gp1`` to ``p1.backward(gp1)`` (or use ``chamfer_loss`` when the upstream gradient is not known in advance)."""
    p1, p2 = _check_clouds(p1.detach(), p2.detach())
    B, N = p1.shape[0], p1.shape[1]
    M = p2.shape[1]
    dev = p1.device
    idx = torch.empty((B * (N + M),), dtype=torch.int64, device=dev)
    loss = torch.empty((B,), dtype=torch.float32, device=dev)
    gp1 = torch.empty_like(p1) if need_gp1 else None
    gp2 = torch.empty_like(p2) if need_gp2 else None
    if gloss is not None:
        if gloss.shape != (B,) or gloss.device != dev:
            raise RuntimeError(f"gloss must have shape ({B},) on {dev}, got {tuple(gloss.shape)} on {gloss.device}")
        gloss = gloss.detach().to(torch.float32).contiguous()
    if B > 0:
        with _device_guard(dev):
            ws, nbytes = _workspace(B, N, M, dev)
            ip = idx.data_ptr()
            _lib.check(_lib.lib().fpsg_chamfer_loss_fwd_bwd(
                p1.data_ptr(), p2.data_ptr(), B, N, M, float(w1), float(w2), gloss.data_ptr() if gloss is not None else None,
                loss.data_ptr(), None, None, ip, ip + 8 * B * N, gp1.data_ptr() if need_gp1 else None,
                gp2.data_ptr() if need_gp2 else None, ws.data_ptr(), nbytes, _stream_ptr(dev)))
    if return_indices:
        return loss, gp1, gp2, idx[:B * N].view(B, N), idx[B * N:].view(B, M)
    return loss, gp1, gp2


def chamfer_loss_multi_and_grad(pairs, weights=None, w1: float = 1.0, w2: float = 1.0, gtotal: torch.Tensor | None = None,
                                need_grad=None):
    """``chamfer_loss_multi`` forward and backward in one fused launch chain (see ``chamfer_loss_and_grad``):
    ``gtotal`` = device scalar, the upstream gradient of the weighted total (``None`` = 1, ``ttl_loss.backward()``);
    ``need_grad`` = one ``(need_gp1, need_gp2)`` per pair of clouds (default ``(True, False)``: the reference's ground
    truth never needs a gradient). Returns ``(total, sums, loss, grads)`` with ``grads[s] = (gp1 or None, gp2 or None)``."""
    pairs = [_check_clouds(a.detach(), b.detach()) for a, b in pairs]
    nseg = len(pairs)
    if not 1 <= nseg <= _lib.MAX_SEGMENTS:
        raise RuntimeError(f"need 1..{_lib.MAX_SEGMENTS} pairs of clouds, got {nseg}")
    weights = [1.0] * nseg if weights is None else [float(w) for w in weights]
    need_grad = [(True, False)] * nseg if need_grad is None else list(need_grad)
    if len(weights) != nseg or len(need_grad) != nseg:
        raise RuntimeError("one weight and one need_grad entry per pair of clouds is required")
    N, M = pairs[0][0].shape[1], pairs[0][1].shape[1]
    dev = pairs[0][0].device
    for a, b in pairs:
        if a.shape[1] != N or b.shape[1] != M:
            raise RuntimeError(f"all segments must share N and M: got {(a.shape[1], b.shape[1])} vs {(N, M)}")
        if a.device != dev:
            raise RuntimeError("all segments must live on the same device")
    Btot = sum(a.shape[0] for a, _ in pairs)
    segs = (_lib.ChamferSegment * nseg)()
    grads = []
    for s, (a, b) in enumerate(pairs):
        g1 = torch.empty_like(a) if need_grad[s][0] else None
        g2 = torch.empty_like(b) if need_grad[s][1] else None
        grads.append((g1, g2))
        segs[s].p1, segs[s].p2 = a.data_ptr(), b.data_ptr()
        segs[s].gp1 = g1.data_ptr() if g1 is not None else None
        segs[s].gp2 = g2.data_ptr() if g2 is not None else None
        segs[s].batch, segs[s].weight = a.shape[0], weights[s]
    idx = torch.empty((Btot * (N + M),), dtype=torch.int64, device=dev)
    out = torch.empty((1 + nseg + Btot,), dtype=torch.float32, device=dev)
    if gtotal is not None:
        gtotal = gtotal.detach().to(torch.float32).reshape(()).contiguous()
    with _device_guard(dev):
        ws, nbytes = _workspace(Btot, N, M, dev) if Btot > 0 else (None, 0)
        _lib.check(_lib.lib().fpsg_chamfer_multi_loss_fwd_bwd(
            segs, nseg, N, M, float(w1), float(w2), gtotal.data_ptr() if gtotal is not None else None, out[1 + nseg:].data_ptr(),
            out.data_ptr(), idx.data_ptr(), idx[Btot * N:].data_ptr(), ws.data_ptr() if ws is not None else None, nbytes, _stream_ptr(dev)))
    return out[0], out[1:1 + nseg], out[1 + nseg:], grads


# One call of the library takes at most 65535 pairs (a grid dimension). Kaolin accepts any batch size, so larger batches go in
# chunks: pairs are independent, and torch.cat keeps the pieces differentiable.
_MAX_PAIRS_PER_CALL = 65535


def _in_chunks(fn, p1: torch.Tensor, p2: torch.Tensor, *args):
    B = p1.shape[0] if p1.dim() == 3 else 0
    if B <= _MAX_PAIRS_PER_CALL or p2.dim() != 3 or p2.shape[0] != B:
        return fn(p1, p2, *args)  # (shape errors are reported by the call itself)
    outs = [fn(p1[s:s + _MAX_PAIRS_PER_CALL], p2[s:s + _MAX_PAIRS_PER_CALL], *args) for s in range(0, B, _MAX_PAIRS_PER_CALL)]
    if isinstance(outs[0], tuple):
        return tuple(torch.cat(o, dim=0) for o in zip(*outs))
    return torch.cat(outs, dim=0)


def chamfer_distance(p1: torch.Tensor, p2: torch.Tensor):
    """(dist1 (B,N), dist2 (B,M), idx1 (B,N) int64, idx2 (B,M) int64): squared-L2 nearest-neighbour distances in
    both directions and their arg-mins (lowest index on exact ties). Differentiable w.r.t. p1 and p2 through
    dist1/dist2; the index outputs are non-differentiable."""
    return _in_chunks(_ChamferDistance.apply, p1, p2)


def sided_distance(p1: torch.Tensor, p2: torch.Tensor):
    """Kaolin's ``sided_distance``: (dist (B,N), idx (B,N) int64) from every p1 point to its nearest p2 point."""
    return _in_chunks(_SidedDistance.apply, p1, p2)


def chamfer_loss(p1: torch.Tensor, p2: torch.Tensor, w1: float = 1.0, w2: float = 1.0) -> torch.Tensor:
    """(B,) = w1 * mean_i dist1 + w2 * mean_j dist2 -- the tensor ``kaolin.metrics.pointcloud.chamfer_distance``
    returns and the reference sums (src/models/few_shot.py:110,117,167)."""
    return _in_chunks(_ChamferLoss.apply, p1, p2, float(w1), float(w2))
