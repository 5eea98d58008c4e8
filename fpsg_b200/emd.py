"""Sinkhorn EMD on B200: the reference's evaluation metric ``emd_wrapper(pc1, pc2)`` (reference src/models/utils.py:9-13,
``neuralnet_pytorch.metrics.emd_loss(pc1, pc2, reduce='sum', sinkhorn=True)``; used at src/models/few_shot.py:168 under
``torch.no_grad()``), calling ``fpsg_emd_sinkhorn_fwd`` of libfpsg_b200.so.

PARITY UNPINNED: neither neuralnet_pytorch nor the geomloss package it delegates to (``geomloss.SamplesLoss()``) is in the
reference tree or in this image. What is implemented is geomloss' published default as recalled: the debiased Sinkhorn
divergence, p = 2 (cost |x-y|^2 / 2), blur = 0.05, scaling = 0.5, uniform weights, epsilon-scaling from diameter^2 down to
blur^2 (``epsilon_schedule`` below), symmetrised updates. Forward only (the reference evaluates it without gradients)."""
from __future__ import annotations

import ctypes
import math

import torch

from . import _lib
from .chamfer import _check_clouds, _device_guard, _stream_ptr


def epsilon_schedule(diameter: float, blur: float = 0.05, scaling: float = 0.5, p: int = 2):
    """[diameter^p] + [exp(e) for e in arange(p log diameter, p log blur, p log scaling)] + [blur^p]  (geomloss)."""
    out = [diameter ** p]
    e, stop, step = p * math.log(diameter), p * math.log(blur), p * math.log(scaling)
    while e > stop:
        out.append(math.exp(e))
        e += step
    out.append(blur ** p)
    return out


def max_diameter(p1: torch.Tensor, p2: torch.Tensor) -> float:
    """Length of the diagonal of the bounding box of all points of both tensors (geomloss ``max_diameter``). Synchronises."""
    lo = torch.minimum(p1.reshape(-1, 3).amin(0), p2.reshape(-1, 3).amin(0))
    hi = torch.maximum(p1.reshape(-1, 3).amax(0), p2.reshape(-1, 3).amax(0))
    return float((hi - lo).norm())


def emd_sinkhorn(p1: torch.Tensor, p2: torch.Tensor, blur: float = 0.05, scaling: float = 0.5, diameter: float | None = None,
                 debias: bool = True, eps_schedule=None, return_potentials: bool = False):
    """(B,) Sinkhorn divergence between the clouds of every pair; ``p1`` (B,N,3), ``p2`` (B,M,3) float32 CUDA.

    ``eps_schedule`` (a list of regularisation strengths, one symmetrised update each) overrides ``blur`` / ``scaling`` /
    ``diameter``; ``debias=False`` returns the dual value of the plain entropic transport cost OT_eps(p1, p2)."""
    p1, p2 = _check_clouds(p1.detach(), p2.detach())
    B, N, M = p1.shape[0], p1.shape[1], p2.shape[1]
    dev = p1.device
    out = torch.empty((B,), dtype=torch.float32, device=dev)
    f = torch.empty((B, N), dtype=torch.float32, device=dev) if return_potentials else None
    g = torch.empty((B, M), dtype=torch.float32, device=dev) if return_potentials else None
    if B == 0:
        return (out, f, g) if return_potentials else out
    if eps_schedule is None:
        if diameter is None:
            diameter = max_diameter(p1, p2)
        if not diameter > 0.0:
            diameter = blur  # all points coincide: one step at blur^2
        eps_schedule = epsilon_schedule(max(diameter, blur), blur, scaling)
    eps = (ctypes.c_float * len(eps_schedule))(*[float(e) for e in eps_schedule])
    L = _lib.lib()
    with _device_guard(dev):
        nbytes = int(L.fpsg_emd_sinkhorn_workspace_bytes(B, N, M))
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        _lib.check(L.fpsg_emd_sinkhorn_fwd(p1.data_ptr(), p2.data_ptr(), B, N, M, eps, len(eps_schedule), 1 if debias else 0, out.data_ptr(),
                                           f.data_ptr() if f is not None else None, g.data_ptr() if g is not None else None,
                                           ws.data_ptr(), nbytes, _stream_ptr(dev)))
    return (out, f, g) if return_potentials else out
