"""ctypes binding of ``oracle/libfpsg_oracle.so`` (built from ``fpsg_oracle.c`` by ``oracle/Makefile``).

TEST INFRASTRUCTURE, not product code: the checker for tests/, smoke() and bench.py's CPU-baseline legs.
Every function here mirrors one C entry point; the C file cites the reference lines it restates.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfpsg_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc) if the shared object is missing or older than its source."""
    src = os.path.join(_HERE, "fpsg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libfpsg_oracle.so"], check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_abi_version.restype = ctypes.c_int
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_sided_distance_f32.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _i64p]
        L.oracle_sided_distance_f64.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p, _i64p, _f64p]
        L.oracle_chamfer_fwd_f32.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_float, _f32p, _f32p, _i64p, _i64p, _f32p]
        L.oracle_chamfer_bwd_f32.argtypes = [_f32p, _f32p, _i64p, _i64p, _f32p, _f32p, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, _f32p, _f32p]
        L.oracle_knn_f32.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i64p, _f32p]
        L.oracle_knn_pd_f64.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f64p]
        L.oracle_normalize_clouds_f32.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _f32p]
        for name in ("oracle_sided_distance_f32", "oracle_sided_distance_f64", "oracle_chamfer_fwd_f32",
                     "oracle_chamfer_bwd_f32", "oracle_knn_f32", "oracle_knn_pd_f64", "oracle_set_threads"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _p(a, t):
    return a.ctypes.data_as(t)


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _check_pair(p1, p2):
    p1, p2 = _f32(p1), _f32(p2)
    if p1.ndim != 3 or p2.ndim != 3 or p1.shape[2] != 3 or p2.shape[2] != 3 or p1.shape[0] != p2.shape[0]:
        raise ValueError(f"expected (B,N,3) and (B,M,3), got {p1.shape} and {p2.shape}")
    if p1.shape[1] > 0 and p2.shape[1] == 0:
        raise ValueError("empty candidate cloud")
    return p1, p2


def sided_distance_f32(p1, p2):
    """(dist (B,N) f32, idx (B,N) i64): nearest p2 point of every p1 point, lowest index on ties."""
    p1, p2 = _check_pair(p1, p2)
    B, N, _ = p1.shape
    M = p2.shape[1]
    dist = np.empty((B, N), np.float32)
    idx = np.empty((B, N), np.int64)
    if B * N:
        lib().oracle_sided_distance_f32(_p(p1, _f32p), _p(p2, _f32p), B, N, M, _p(dist, _f32p), _p(idx, _i64p))
    return dist, idx


def sided_distance_f64(p1, p2):
    """Float64 adjudicator: (dist f64, idx i64, gap f64 = second-best minus best)."""
    p1, p2 = _check_pair(p1, p2)
    B, N, _ = p1.shape
    M = p2.shape[1]
    dist = np.empty((B, N), np.float64)
    idx = np.empty((B, N), np.int64)
    gap = np.empty((B, N), np.float64)
    if B * N:
        lib().oracle_sided_distance_f64(_p(p1, _f32p), _p(p2, _f32p), B, N, M, _p(dist, _f64p), _p(idx, _i64p),
                                        _p(gap, _f64p))
    return dist, idx, gap


def chamfer_fwd(p1, p2, w1: float = 1.0, w2: float = 1.0):
    """(dist1, dist2, idx1, idx2, loss (B,)) -- loss = w1*mean(dist1) + w2*mean(dist2) per pair."""
    p1, p2 = _check_pair(p1, p2)
    B, N, _ = p1.shape
    M = p2.shape[1]
    d1 = np.empty((B, N), np.float32)
    d2 = np.empty((B, M), np.float32)
    i1 = np.empty((B, N), np.int64)
    i2 = np.empty((B, M), np.int64)
    loss = np.empty((B,), np.float32)
    if B and N and M:
        lib().oracle_chamfer_fwd_f32(_p(p1, _f32p), _p(p2, _f32p), B, N, M, w1, w2, _p(d1, _f32p), _p(d2, _f32p),
                                     _p(i1, _i64p), _p(i2, _i64p), _p(loss, _f32p))
    return d1, d2, i1, i2, loss


def chamfer_bwd(p1, p2, idx1, idx2, g1, g2, need_p1: bool = True, need_p2: bool = True):
    """(grad_p1 | None, grad_p2 | None) for upstream gradients g1 (B,N) of dist1 and g2 (B,M) of dist2."""
    p1, p2 = _check_pair(p1, p2)
    B, N, _ = p1.shape
    M = p2.shape[1]
    idx1, idx2, g1, g2 = _i64(idx1), _i64(idx2), _f32(g1), _f32(g2)
    assert idx1.shape == (B, N) and idx2.shape == (B, M) and g1.shape == (B, N) and g2.shape == (B, M)
    gp1 = np.zeros((B, N, 3), np.float32) if need_p1 else None
    gp2 = np.zeros((B, M, 3), np.float32) if need_p2 else None
    if B and N and M:
        lib().oracle_chamfer_bwd_f32(_p(p1, _f32p), _p(p2, _f32p), _p(idx1, _i64p), _p(idx2, _i64p), _p(g1, _f32p),
                                     _p(g2, _f32p), B, N, M, _p(gp1, _f32p) if need_p1 else None,
                                     _p(gp2, _f32p) if need_p2 else None)
    return gp1, gp2


def knn_f32(x, k: int, return_pd: bool = False):
    """idx (B,N,k) i64 of reference ``knn(x, k)`` (src/dgcnn/model.py:13-20) for x (B,C,N)."""
    x = _f32(x)
    if x.ndim != 3:
        raise ValueError(f"expected (B,C,N), got {x.shape}")
    B, C, N = x.shape
    if not (1 <= k <= N):
        raise ValueError(f"k={k} out of range for N={N}")
    idx = np.empty((B, N, k), np.int64)
    pd = np.empty((B, N, N), np.float32) if return_pd else None
    if B:
        lib().oracle_knn_f32(_p(x, _f32p), B, C, N, k, _p(idx, _i64p), _p(pd, _f32p) if return_pd else None)
    return (idx, pd) if return_pd else idx


def knn_pd_f64(x):
    """Float64 difference-form matrix -(|x_i - x_j|^2), (B,N,N): the near-tie adjudicator for knn."""
    x = _f32(x)
    B, C, N = x.shape
    pd = np.empty((B, N, N), np.float64)
    if B * N:
        lib().oracle_knn_pd_f64(_p(x, _f32p), B, C, N, _p(pd, _f64p))
    return pd


def normalize_clouds(x):
    """(B,N,3) float32 -> every cloud centred on its mean point and divided by its largest point norm, in numpy's
    float32 arithmetic (reference src/datasets/modelnet.py:66-69)."""
    x = _f32(x)
    assert x.ndim == 3 and x.shape[2] == 3
    out = np.empty_like(x)
    if x.shape[0] and x.shape[1]:
        lib().oracle_normalize_clouds_f32(_p(x, _f32p), x.shape[0], x.shape[1], _p(out, _f32p))
    return out


def ply_reader_py(file_path, n_verts=2048):
    """ply_reader of the reference (src/datasets/modelnet.py:15-29), restated line for line in Python: the checker
    of the C parser (the real function's outputs are in tests/golden/pipeline_ply.npz)."""
    with open(file_path, "r") as f:
        while True:
            cur = f.readline().strip()
            if cur == "end_header":
                break
            cur = cur.split(" ")
            if len(cur) > 2 and cur[1] == "vertex":
                n_verts = min(int(cur[2]), n_verts)
        vertices = [[float(s) for s in f.readline().strip().split(" ")] for _ in range(n_verts)]
    return np.asarray(vertices, dtype=np.float32)
