"""Measurement only: the sweep kernel with the EXPANDED form |a|^2 + |b|^2 - 2 a.b in its inner loop (1 FADD + 3 FFMA per evaluation
instead of 3 FADD + FMUL + 2 FFMA) -- NOT the reference's rounding, results are not usable -- to see what a filter kernel built on it
could gain inside the real kernel (bookkeeping, column records, prefetch included). Builds: nvcc -DFPSG_TUNING
-DFPSG_SWEEP_EXPANDED_PROBE=<CTAs per SM> -> fpsg_b200/libfpsg_b200_probe<n>.so. Usage: expanded_probe.py <lib.so>"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpsg_b200 import _lib
if len(sys.argv) > 1:
    _lib.use_library(os.path.abspath(sys.argv[1]))
from tests.helpers import make_pair
L = _lib.lib()
t = ctypes.c_double(); mhz = ctypes.c_double()
_lib.check(L.fpsg_probe_fp32_peak(ctypes.byref(t), ctypes.byref(mhz), None))
out = {"lib": os.path.basename(sys.argv[1]) if len(sys.argv) > 1 else "product", "fp32_peak_tflops": round(t.value, 2)}
for (B, N, M) in [(32, 2048, 2048), (148, 2048, 2048), (32, 4096, 4096)]:
    sets = [tuple(x.cuda() for x in make_pair(B, N, M, seed=s)) for s in range(max(2, int(200e6 / (B * (N + M) * 12))))]
    nb = L.fpsg_chamfer_fwd_workspace_bytes(B, N, M); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def f(i):
        a, b = sets[i % len(sets)]
        _lib.check(L.fpsg_chamfer_sweep_only(a.data_ptr(), b.data_ptr(), B, N, M, ws.data_ptr(), nb, st))
    for i in range(10): f(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): f(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    out[f"B{B}_N{N}_sweep_us"] = round(us, 2)
    del sets
print(json.dumps(out))
