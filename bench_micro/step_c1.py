"""c1 step anatomy through the C ABI (development aid): sweep alone, forward (sweep + finalize), backward alone, the
chained three-kernel step and the fused step (sweep + finish kernel), over the same ring of HBM-resident inputs
bench.py uses. Usage: step_c1.py [B N M]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpsg_b200 import _lib, build
if os.environ.get("FPSG_TUNING_LIB"):   # the development build, so that the FPSG_* ablation switches take effect
    _lib.use_library(build.LIB_TUNING)
from bench import make_inputs

L = _lib.lib()
args = [int(v) for v in sys.argv[1:4]] if len(sys.argv) >= 4 else [32, 2048, 2048]
B, N, M = args
dev = torch.device("cuda", 0)
sets = max(4, min(96, int(160e6 / (B * (N + M) * 12))))
ring = [make_inputs(B, N, M, seed=1235 + s, device=dev) for s in range(sets)]
ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in ring]
i1 = torch.empty(B, N, dtype=torch.int64, device=dev); i2 = torch.empty(B, M, dtype=torch.int64, device=dev)
loss = torch.empty(B, device=dev); gl = torch.ones(B, device=dev); gp1 = torch.empty(B, N, 3, device=dev)
nb = L.fpsg_chamfer_fwd_workspace_bytes(B, N, M); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
o = (loss.data_ptr(), i1.data_ptr(), i2.data_ptr(), ws.data_ptr(), gl.data_ptr(), gp1.data_ptr())

def sweep(i):
    a, b = ptrs[i % sets]
    _lib.check(L.fpsg_chamfer_sweep_only(a, b, B, N, M, o[3], nb, st))
def fwd(i):
    a, b = ptrs[i % sets]
    _lib.check(L.fpsg_chamfer_loss_fwd(a, b, B, N, M, 1.0, 1.0, o[0], None, None, o[1], o[2], o[3], nb, st))
def bwd(i):
    a, b = ptrs[i % sets]
    _lib.check(L.fpsg_chamfer_loss_bwd(a, b, o[1], o[2], o[4], 1.0, 1.0, B, N, M, o[5], None, None, 0, st))
def step(i):
    fwd(i); bwd(i)
def fused(i):
    a, b = ptrs[i % sets]
    _lib.check(L.fpsg_chamfer_loss_fwd_bwd(a, b, B, N, M, 1.0, 1.0, None, o[0], None, None, o[1], o[2], o[5], None, o[3], nb, st))

def t(fn, it):
    for i in range(20): fn(i)
    torch.cuda.synchronize()
    best = []
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(it): fn(20 + i)
        e1.record(); torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / it * 1e3)
    return round(min(best), 2)

it = 400 if B * N * M < 5e9 else 20
fwd(0); torch.cuda.synchronize()
print(json.dumps({"B": B, "N": N, "M": M, "sweep_us": t(sweep, it), "fwd_us": t(fwd, it), "bwd_us": t(bwd, it), "step_us": t(step, it),
                  "fused_step_us": t(fused, it), "env": {k: v for k, v in os.environ.items() if k.startswith("FPSG_")}}))
