"""Builds fpsg_b200/libfpsg_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery:
the library is a plain C-ABI shared object, see include/fpsg_b200.h)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfpsg_b200.so")
LIB_TUNING = os.path.join(HERE, "libfpsg_b200_tuning.so")  # development build: FPSG_* ablation switches + timeline stamps
SOURCES = ["api.cu", "chamfer_fwd.cu", "chamfer_finish.cu", "chamfer_bwd.cu", "pipeline.cu", "emd_sinkhorn.cu", "knn.cu", "knn_tc.cu", "edge_feature.cu", "pointcloud_io.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-cudart", "static", "-diag-suppress", "550"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libfpsg_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "fpsg_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, tuning: bool = False) -> str:
    """tuning=True builds the development variant (-DFPSG_TUNING: the FPSG_* ablation switches and the timeline stamps used by
    bench_micro/) as a separate file; the product library never contains them."""
    out = LIB_TUNING if tuning else LIB
    if not force and not tuning and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DFPSG_TUNING"] if tuning else []) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build_library(force=True, verbose="-v" in sys.argv, tuning="--tuning" in sys.argv))
