"""Sweep kernel alone over a few batch sizes (development aid): how much of the c1 gap is partition granularity."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpsg_b200 import _lib
from tests.helpers import make_pair
L = _lib.lib()
t = ctypes.c_double(); mhz = ctypes.c_double()
_lib.check(L.fpsg_probe_fp32_peak(ctypes.byref(t), ctypes.byref(mhz), None))
peak = t.value
shapes = [(32, 2048, 2048), (37, 2048, 2048), (64, 2048, 2048), (74, 2048, 2048), (148, 2048, 2048), (32, 4096, 4096)]
if len(sys.argv) > 1: shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (B, N, M) in shapes:
    sets = [tuple(x.cuda() for x in make_pair(B, N, M, seed=s)) for s in range(max(2, int(200e6 / (B * (N + M) * 12))))]
    nb = L.fpsg_chamfer_fwd_workspace_bytes(B, N, M); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def f(i):
        a, b = sets[i % len(sets)]
        _lib.check(L.fpsg_chamfer_sweep_only(a.data_ptr(), b.data_ptr(), B, N, M, ws.data_ptr(), nb, st))
    for i in range(10): f(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    it = 200
    e0.record()
    for i in range(it): f(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / it * 1e3
    tf = 16.0 * B * N * M / us / 1e6
    print(json.dumps({"B": B, "N": N, "M": M, "sweep_us": round(us, 2), "tflops": round(tf, 2), "frac": round(tf / peak, 3)}))
