"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/fpsg_b200.h declares,
sizes workspaces without a GPU, and the host code refuses to run without CUDA tensors (no fallback)."""
import ctypes
import sys

import pytest
import torch

import __graft_entry__ as entry
from fpsg_b200 import _lib


def test_library_exports_every_declared_symbol():
    names = entry.declared_symbols()
    assert {"fpsg_chamfer_fwd", "fpsg_chamfer_bwd", "fpsg_chamfer_loss_fwd", "fpsg_chamfer_loss_bwd", "fpsg_knn_fwd",
            "fpsg_last_error", "fpsg_abi_version"} <= set(names)
    raw = ctypes.CDLL(_lib.LIB)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
    assert _lib.lib().fpsg_abi_version() == _lib.ABI_VERSION


def test_workspace_queries_need_no_gpu():
    L = _lib.lib()
    assert L.fpsg_chamfer_fwd_workspace_bytes(0, 10, 10) == 0
    small = L.fpsg_chamfer_fwd_workspace_bytes(2, 100, 100)
    big = L.fpsg_chamfer_fwd_workspace_bytes(32, 2048, 2048)
    assert 0 < small < big < 64 * 2 ** 20
    assert L.fpsg_knn_workspace_bytes(32, 3, 1024, 20) >= 32 * 1024 * 4
    assert L.fpsg_chamfer_bwd_workspace_bytes(32, 2048, 2048) == 0


def test_invalid_arguments_are_reported_not_crashed():
    L = _lib.lib()
    rc = L.fpsg_knn_fwd(None, 1, 3, 8, 9, None, None, 0, None)  # k > N is rejected before any CUDA call
    assert rc == -1 and b"k=9" in L.fpsg_last_error()
    rc = L.fpsg_chamfer_fwd(None, None, 1, 0, 4, None, None, None, None, None, 0, None)
    assert rc == -1
    with pytest.raises(_lib.FpsgError):
        _lib.check(rc)


def test_multi_segment_argument_checks_need_no_gpu():
    L = _lib.lib()
    segs = (_lib.ChamferSegment * 2)()
    assert L.fpsg_chamfer_multi_loss_fwd(segs, 0, 8, 8, 1.0, 1.0, None, None, None, None, None, 0, None) == -1
    assert L.fpsg_chamfer_multi_loss_fwd(segs, _lib.MAX_SEGMENTS + 1, 8, 8, 1.0, 1.0, None, None, None, None, None, 0, None) == -1
    segs[0].batch = -1
    assert L.fpsg_chamfer_multi_loss_fwd(segs, 2, 8, 8, 1.0, 1.0, None, None, None, None, None, 0, None) == -1
    assert b"batch" in L.fpsg_last_error()
    segs[0].batch = 0
    assert L.fpsg_chamfer_multi_loss_bwd(segs, 2, None, None, None, 1.0, 1.0, 8, 8, None) == -1  # null gtotal
    import ctypes as C
    assert C.sizeof(_lib.ChamferSegment) == 40  # 4 pointers + int + float, as the C struct


def test_host_code_refuses_cpu_tensors():
    import fpsg_b200
    with pytest.raises(RuntimeError, match="CUDA"):
        fpsg_b200.chamfer_distance(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        fpsg_b200.chamfer_loss(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        fpsg_b200.knn(torch.zeros(1, 3, 8), 2)


def test_kaolin_shim_installs_and_exposes_reference_surface():
    import fpsg_b200.shims as shims
    saved = {k: sys.modules.get(k) for k in ("kaolin", "kaolin.metrics", "kaolin.metrics.pointcloud")}
    try:
        shims.install(force=True, patch_dgcnn=False)
        from kaolin.metrics.pointcloud import chamfer_distance, sided_distance  # the reference's import line
        import inspect
        sig = inspect.signature(chamfer_distance)
        assert list(sig.parameters)[:4] == ["p1", "p2", "w1", "w2"]
        assert callable(sided_distance)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_product_never_imports_the_oracle():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "fpsg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # comments may CITE the oracle (which function a kernel is checked against); nothing may import, include,
                # load or link it
                assert "import oracle" not in text and "from oracle" not in text and "libfpsg_oracle" not in text, f
                import re
                assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', text), f
                assert not re.search(r'(dlopen|CDLL|LoadLibrary)\([^)]*oracle', text), f
