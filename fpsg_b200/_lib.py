"""ctypes loader for libfpsg_b200.so (C ABI in include/fpsg_b200.h).

There is deliberately no fallback of any kind: if the shared object is missing or a call fails, the caller gets
a RuntimeError. Build it with ``python fpsg_b200/build.py`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os

from .build import LIB

ABI_VERSION = 1
KNN_MAX_K = 128

_f32p = ctypes.c_void_p  # device pointers travel as integers
_i64p = ctypes.c_void_p
_lib = None


MAX_SEGMENTS = 8


class FpsgError(RuntimeError):
    pass


class ChamferSegment(ctypes.Structure):
    """``fpsg_chamfer_segment_t`` of include/fpsg_b200.h."""
    _fields_ = [("p1", ctypes.c_void_p), ("p2", ctypes.c_void_p), ("gp1", ctypes.c_void_p), ("gp2", ctypes.c_void_p),
                ("batch", ctypes.c_int), ("weight", ctypes.c_float)]


MAX_PEERS = 8
PEER_RING = 32


class PeerExchange(ctypes.Structure):
    """``fpsg_peer_exchange_t`` of include/fpsg_b200.h."""
    _fields_ = [("table", ctypes.c_void_p * MAX_PEERS), ("global_out", ctypes.c_void_p), ("rank", ctypes.c_int), ("world", ctypes.c_int),
                ("seq", ctypes.c_uint), ("lag", ctypes.c_int)]


def _declare(L: ctypes.CDLL) -> None:
    c_int, c_size_t, c_float, vp = ctypes.c_int, ctypes.c_size_t, ctypes.c_float, ctypes.c_void_p
    L.fpsg_abi_version.restype = c_int
    L.fpsg_abi_version.argtypes = []
    L.fpsg_last_error.restype = ctypes.c_char_p
    L.fpsg_last_error.argtypes = []
    L.fpsg_chamfer_fwd_workspace_bytes.restype = c_size_t
    L.fpsg_chamfer_fwd_workspace_bytes.argtypes = [c_int, c_int, c_int]
    L.fpsg_chamfer_fwd.restype = c_int
    L.fpsg_chamfer_fwd.argtypes = [vp, vp, c_int, c_int, c_int, vp, vp, vp, vp, vp, c_size_t, vp]
    L.fpsg_chamfer_bwd_workspace_bytes.restype = c_size_t
    L.fpsg_chamfer_bwd_workspace_bytes.argtypes = [c_int, c_int, c_int]
    L.fpsg_chamfer_bwd.restype = c_int
    L.fpsg_chamfer_bwd.argtypes = [vp, vp, vp, vp, vp, vp, c_int, c_int, c_int, vp, vp, vp, c_size_t, vp]
    L.fpsg_chamfer_loss_fwd.restype = c_int
    L.fpsg_chamfer_loss_fwd.argtypes = [vp, vp, c_int, c_int, c_int, c_float, c_float, vp, vp, vp, vp, vp, vp, c_size_t, vp]
    L.fpsg_chamfer_loss_bwd.restype = c_int
    L.fpsg_chamfer_loss_bwd.argtypes = [vp, vp, vp, vp, vp, c_float, c_float, c_int, c_int, c_int, vp, vp, vp, c_size_t, vp]
    segp = ctypes.POINTER(ChamferSegment)
    L.fpsg_chamfer_multi_loss_fwd.restype = c_int
    L.fpsg_chamfer_multi_loss_fwd.argtypes = [segp, c_int, c_int, c_int, c_float, c_float, vp, vp, vp, vp, vp, c_size_t, vp]
    L.fpsg_chamfer_multi_loss_bwd.restype = c_int
    L.fpsg_chamfer_multi_loss_bwd.argtypes = [segp, c_int, vp, vp, vp, c_float, c_float, c_int, c_int, vp]
    L.fpsg_chamfer_loss_fwd_bwd.restype = c_int
    L.fpsg_chamfer_loss_fwd_bwd.argtypes = [vp, vp, c_int, c_int, c_int, c_float, c_float, vp, vp, vp, vp, vp, vp, vp, vp, vp, c_size_t, vp]
    L.fpsg_chamfer_multi_loss_fwd_bwd.restype = c_int
    L.fpsg_chamfer_multi_loss_fwd_bwd.argtypes = [segp, c_int, c_int, c_int, c_float, c_float, vp, vp, vp, vp, vp, vp, c_size_t, vp]
    L.fpsg_chamfer_multi_loss_fwd_bwd_dist.restype = c_int
    L.fpsg_chamfer_multi_loss_fwd_bwd_dist.argtypes = [segp, c_int, c_int, c_int, c_float, c_float, vp, vp, vp, vp, vp, vp, c_size_t,
                                                       ctypes.POINTER(PeerExchange), vp]
    L.fpsg_peer_exchange_collect.restype = c_int
    L.fpsg_peer_exchange_collect.argtypes = [ctypes.POINTER(PeerExchange), vp]
    szp, vpp = ctypes.POINTER(c_size_t), ctypes.POINTER(vp)
    L.fpsg_chamfer_pipeline_arena_bytes.restype = c_int
    L.fpsg_chamfer_pipeline_arena_bytes.argtypes = [c_int, c_int, c_int, c_int, szp, szp]
    L.fpsg_chamfer_pipeline_create.restype = c_int
    L.fpsg_chamfer_pipeline_create.argtypes = [c_int, c_int, c_int, c_float, c_float, c_int, vp, c_size_t, vp, c_size_t, vpp]
    L.fpsg_chamfer_pipeline_submit.restype = c_int
    L.fpsg_chamfer_pipeline_submit.argtypes = [vp, vp, vp]
    L.fpsg_chamfer_pipeline_submit_device.restype = c_int
    L.fpsg_chamfer_pipeline_submit_device.argtypes = [vp, vp, vp, vp]
    L.fpsg_chamfer_pipeline_result.restype = c_int
    L.fpsg_chamfer_pipeline_result.argtypes = [vp, ctypes.POINTER(c_float), vp, vpp, vpp, vpp]
    L.fpsg_chamfer_pipeline_set_exchange.restype = c_int
    L.fpsg_chamfer_pipeline_set_exchange.argtypes = [vp, ctypes.POINTER(vp), c_int, c_int]
    L.fpsg_chamfer_pipeline_previous_global_sum.restype = c_int
    L.fpsg_chamfer_pipeline_previous_global_sum.argtypes = [vp, ctypes.POINTER(c_float)]
    L.fpsg_chamfer_pipeline_compute_stream.restype = vp
    L.fpsg_chamfer_pipeline_compute_stream.argtypes = [vp]
    L.fpsg_chamfer_pipeline_destroy.restype = c_int
    L.fpsg_chamfer_pipeline_destroy.argtypes = [vp]
    L.fpsg_chamfer_ws_diag_offset.restype = c_size_t
    L.fpsg_chamfer_ws_diag_offset.argtypes = [c_int, c_int, c_int]
    L.fpsg_emd_sinkhorn_workspace_bytes.restype = c_size_t
    L.fpsg_emd_sinkhorn_workspace_bytes.argtypes = [c_int, c_int, c_int]
    L.fpsg_emd_sinkhorn_fwd.restype = c_int
    L.fpsg_emd_sinkhorn_fwd.argtypes = [vp, vp, c_int, c_int, c_int, ctypes.POINTER(c_float), c_int, c_int, vp, vp, vp, vp, c_size_t, vp]
    L.fpsg_knn_workspace_bytes.restype = c_size_t
    L.fpsg_knn_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    L.fpsg_knn_fwd.restype = c_int
    L.fpsg_knn_fwd.argtypes = [vp, c_int, c_int, c_int, c_int, vp, vp, c_size_t, vp]
    L.fpsg_edge_feature_fwd.restype = c_int
    L.fpsg_edge_feature_fwd.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, vp]
    L.fpsg_edge_feature_bwd_workspace_bytes.restype = c_size_t
    L.fpsg_edge_feature_bwd_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    L.fpsg_edge_feature_bwd.restype = c_int
    L.fpsg_edge_feature_bwd.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, vp, c_size_t, vp]
    L.fpsg_normalize_clouds.restype = c_int
    L.fpsg_normalize_clouds.argtypes = [vp, c_int, c_int, vp, vp]
    L.fpsg_ply_read_ascii_host.restype = c_int
    L.fpsg_ply_read_ascii_host.argtypes = [ctypes.c_char_p, c_int, vp, ctypes.POINTER(c_int)]
    L.fpsg_chamfer_sweep_only.restype = c_int
    L.fpsg_chamfer_sweep_only.argtypes = [vp, vp, c_int, c_int, c_int, vp, c_size_t, vp]
    L.fpsg_probe_fp32_peak.restype = c_int
    L.fpsg_probe_fp32_peak.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), vp]


_path = LIB


def use_library(path: str) -> None:
    """Development hook (bench_micro/): load another build of the library -- the tuning variant with its ablation switches and
    timeline stamps -- instead of the product one. Must be called before the first use."""
    global _path
    if _lib is not None:
        raise FpsgError("use_library() must be called before the library is first used")
    _path = path


def lib() -> ctypes.CDLL:
    """The loaded library. Raises if it has not been built: the product never runs without its CUDA code."""
    global _lib
    if _lib is None:
        if not os.path.exists(_path):
            raise FpsgError(f"{_path} is missing: build it with `python fpsg_b200/build.py` "
                            "(fpsg_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(_path)
        _declare(L)
        got = L.fpsg_abi_version()
        if got != ABI_VERSION:
            raise FpsgError(f"libfpsg_b200.so ABI version {got}, host code expects {ABI_VERSION}: rebuild")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().fpsg_last_error().decode("utf-8", "replace")
        raise FpsgError(f"fpsg_b200 error {rc}: {msg}")
