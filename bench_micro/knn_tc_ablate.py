"""Ablation of the tensor-core filter kernel (tuning build, FPSG_TC_DBG bits: 1 = epilogue without TMEM loads / arithmetic,
2 = no MMA issue, 4 = no TMA loads of the candidate tiles): prep + filter time per call. Usage: knn_tc_ablate.py B C N k"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpsg_b200 import _lib, build
_lib.use_library(build.LIB_TUNING)
from tests.helpers import make_cloud_cn
L = _lib.lib()
B, C, N, k = (int(v) for v in sys.argv[1:5])
vp = ctypes.c_void_p
L.fpsg_tune_knn_tc_scores.restype = ctypes.c_int
L.fpsg_tune_knn_tc_scores.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_size_t, vp]
x = make_cloud_cn(B, C, N).cuda()
nb = L.fpsg_knn_workspace_bytes(B, C, N, k)
ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
res = {"B": B, "C": C, "N": N, "k": k}
for dbg in (0, 1, 2, 3, 4, 5, 6, 7):
    os.environ["FPSG_TC_DBG"] = str(dbg)
    for _ in range(3):
        _lib.check(L.fpsg_tune_knn_tc_scores(x.data_ptr(), B, C, N, k, None, ws.data_ptr(), nb, st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.fpsg_tune_knn_tc_scores(x.data_ptr(), B, C, N, k, None, ws.data_ptr(), nb, st)
    e1.record()
    torch.cuda.synchronize()
    res[f"dbg{dbg}_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 2)
print(json.dumps(res))

# timeline of CTA 0 (dbg bit 8), with and without the work
import numpy as np
for dbg in (8, 15):
    os.environ["FPSG_TC_DBG"] = str(dbg)
    tr = torch.zeros(3 * 512, dtype=torch.int64, device="cuda")
    _lib.check(L.fpsg_tune_knn_tc_scores(x.data_ptr(), B, C, N, k, tr.data_ptr(), ws.data_ptr(), nb, st))
    torch.cuda.synchronize()
    t = tr.cpu().numpy().reshape(3, 512)
    t0 = t[t > 0].min()
    rel = np.where(t > 0, t - t0, -1)
    out = {"dbg": dbg, "producer_stage_issue_ns": rel[0, :60].tolist(), "producer_aempty_ns": rel[0, 400:404].tolist(),
           "mma_tempty_acquired_tfull_commit_ns": rel[1, :40].tolist(), "mma_afull_ns": rel[1, 400:404].tolist(),
           "epi_tfull_acquired_tempty_arrived_ns": rel[2, :40].tolist()}
    print(json.dumps(out))
