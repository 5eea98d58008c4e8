"""Shared input generators and parity gates for the tests."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_pair(B: int, N: int, M: int, seed: int = 1234, kind: str = "decoder"):
    """Synthetic inputs of SURVEY.md 8(d): p1 = tanh(0.5*randn) (decoder-like), p2 = unit-ball normalised randn.
    kind='uniform' gives both clouds U[-1,1]^3."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if kind == "uniform":
        p1 = torch.rand(B, N, 3, generator=g) * 2 - 1
        p2 = torch.rand(B, M, 3, generator=g) * 2 - 1
    else:
        p1 = torch.tanh(0.5 * torch.randn(B, N, 3, generator=g))
        p2 = torch.randn(B, M, 3, generator=g)
        if M > 0:
            p2 = p2 - p2.mean(1, keepdim=True)
            p2 = p2 / p2.norm(dim=2).amax(1).clamp_min(1e-12)[:, None, None]
    return p1.contiguous(), p2.contiguous()


def make_cloud_cn(B: int, C: int, N: int, seed: int = 1234):
    """kNN input (B,C,N): C==3 -> unit-ball cloud transposed (config c3), else randn features."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    if C == 3:
        p = torch.randn(B, N, 3, generator=g)
        p = p - p.mean(1, keepdim=True)
        p = p / p.norm(dim=2).amax(1).clamp_min(1e-12)[:, None, None]
        return p.transpose(1, 2).contiguous()
    return torch.randn(B, C, N, generator=g)


def assert_close_rel(actual, expected, rel=1e-5, what="", scale_frac=1e-3):
    """|a-e| <= rel * max(|e|, scale) elementwise, scale = max|expected| * scale_frac guards exact zeros /
    cancellation in sums of signed terms (gradients). rel=1e-5 is the north-star tolerance for distances and
    gradients. scale_frac=1.0 makes the test norm-wise (used only where the two sides round the upstream
    gradient differently, so cancellation legitimately amplifies the per-element relative error)."""
    a = np.asarray(actual, dtype=np.float64)
    e = np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    if a.size == 0:
        return
    scale = max(float(np.abs(e).max()) * scale_frac, 1e-30)
    err = np.abs(a - e) / np.maximum(np.abs(e), scale)
    worst = float(err.max())
    assert worst <= rel, f"{what}: max relative error {worst:.3e} > {rel:.1e} at {np.unravel_index(err.argmax(), err.shape)}"


def assert_knn_equivalent(idx, ref_idx, pd32, what=""):
    """kNN gate: every row holds k distinct valid indices, and the sequence of reference FP32
    pairwise values at the chosen indices is identical to the reference's -- i.e. the two results may differ only
    by a permutation / substitution among candidates whose reference distances are exactly equal in FP32."""
    idx = np.asarray(idx)
    ref_idx = np.asarray(ref_idx)
    assert idx.shape == ref_idx.shape, f"{what}: shape {idx.shape} vs {ref_idx.shape}"
    assert idx.dtype == np.int64
    N = pd32.shape[-1]
    assert idx.min() >= 0 and idx.max() < N, f"{what}: index out of range"
    srt = np.sort(idx, axis=-1)
    assert (srt[..., 1:] != srt[..., :-1]).all(), f"{what}: duplicate neighbour in a row"
    va = np.take_along_axis(pd32, idx, 2)
    vb = np.take_along_axis(pd32, ref_idx, 2)
    bad = int((va != vb).sum())
    assert bad == 0, f"{what}: {bad} positions differ beyond exact FP32 ties"
