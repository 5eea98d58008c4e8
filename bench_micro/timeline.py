"""Timeline of one fused Chamfer step (development aid, tuning build): %globaltimer stamps written by the sweep and finish
kernels (FPSG_TRACE in the sources), reported relative to the first sweep CTA's start. Usage: timeline.py [B N M] [--three]
Prints one JSON object: per kernel and phase the min / median / max over CTAs, in microseconds."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fpsg_b200 import _lib, build
if not os.path.exists(build.LIB_TUNING):
    build.build_library(force=True, tuning=True)
_lib.use_library(build.LIB_TUNING)
from bench import make_inputs
import ctypes

L = _lib.lib()
nums = [int(v) for v in sys.argv[1:] if v.isdigit()]
B, N, M = nums[:3] if len(nums) >= 3 else (32, 2048, 2048)
three = "--three" in sys.argv
dev = torch.device("cuda", 0)
KC, KP = 4096, 8
trace = torch.zeros(4 * KC * KP, dtype=torch.int64, device=dev)
L.fpsg_tune_set_trace.argtypes = [ctypes.c_void_p]
L.fpsg_tune_set_trace(trace.data_ptr())
sets = max(4, min(96, int(160e6 / (B * (N + M) * 12))))
ring = [make_inputs(B, N, M, seed=1235 + s, device=dev) for s in range(sets)]
idx = torch.empty(B * (N + M), dtype=torch.int64, device=dev)
loss = torch.empty(B, device=dev); gl = torch.ones(B, device=dev); gp1 = torch.empty(B, N, 3, device=dev)
nb = L.fpsg_chamfer_fwd_workspace_bytes(B, N, M); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
ip = idx.data_ptr()

def step(i):
    a, b = ring[i % sets]
    if three:
        _lib.check(L.fpsg_chamfer_loss_fwd(a.data_ptr(), b.data_ptr(), B, N, M, 1.0, 1.0, loss.data_ptr(), None, None, ip, ip + 8 * B * N, ws.data_ptr(), nb, st))
        _lib.check(L.fpsg_chamfer_loss_bwd(a.data_ptr(), b.data_ptr(), ip, ip + 8 * B * N, gl.data_ptr(), 1.0, 1.0, B, N, M, gp1.data_ptr(), None, None, 0, st))
    else:
        _lib.check(L.fpsg_chamfer_loss_fwd_bwd(a.data_ptr(), b.data_ptr(), B, N, M, 1.0, 1.0, None, loss.data_ptr(), None, None, ip, ip + 8 * B * N,
                                               gp1.data_ptr(), None, ws.data_ptr(), nb, st))

for i in range(50):
    step(i)
torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.int64).reshape(4, KC, KP)
names = {0: ["start", "pair_published", "run_done"], 1: ["resident", "pair_swept", "merged+staged", "argmin_done", "cluster_barrier2", "end"]}
t0 = t[0][:, 0][t[0][:, 0] > 0].min()
out = {"B": B, "N": N, "M": M, "mode": "three-kernel" if three else "fused"}
for k, ph in names.items():
    for p, name in enumerate(ph):
        v = t[k][:, p]
        v = v[v > 0]
        if v.size:
            r = (v - t0) / 1e3
            out[f"k{k}.{name}"] = {"min": round(float(r.min()), 2), "p50": round(float(np.median(r)), 2), "max": round(float(r.max()), 2), "n": int(v.size)}
# per pair: when was the pair swept, when did its finish cluster end
if not three:
    cl = int(os.environ.get("FPSG_FINISH_CL", 8))
    sw = t[1][:B * cl, 1].reshape(B, cl); en = t[1][:B * cl, 5].reshape(B, cl)
    out["pair_swept_us"] = [round(float((sw[b].max() - t0) / 1e3), 1) for b in range(B)]
    out["pair_finished_us"] = [round(float((en[b].max() - t0) / 1e3), 1) for b in range(B)]
# scheduling priority: per SM, the sweep CTAs in the order they finished, as (blockIdx, warp slot of warp 0, end time)
loc = t[0][:, 3]
grid = int((t[0][:, 0] > 0).sum())
by_sm = {}
for g in range(grid):
    by_sm.setdefault(int(loc[g] >> 32), []).append((round(float((t[0][g, 2] - t0) / 1e3), 1), g, int(loc[g] & 0xffffffff)))
out["sm_examples"] = {str(sm): sorted(v) for sm, v in list(sorted(by_sm.items()))[:6]}
rank_by_wid = []
for sm, v in by_sm.items():
    v = sorted(v)  # by end time
    wids = [x[2] for x in v]
    rank_by_wid.append([sorted(wids, reverse=True).index(wd) for wd in wids])  # 0 = highest warp slot
import collections
out["finish_order_vs_warp_slot_rank"] = collections.Counter(tuple(r) for r in rank_by_wid).most_common(5)
print(json.dumps(out))
