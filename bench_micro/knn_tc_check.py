"""Development check of the tensor-core kNN pre-filter (knn_tc.cu) through the tuning build:
 1. the raw accumulator values s_ij = x_i . x_j - xx_j / 2 of pass 1 against float64 numpy, relative to the bound E the kernel uses;
 2. the candidate masks: they must contain the oracle's top k of every query; survivors per query;
 3. the product call fpsg_knn_fwd against the oracle (bit-exact), and its time.
Usage: knn_tc_check.py [B C N k]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fpsg_b200 import _lib, build
_lib.use_library(build.LIB_TUNING)
from oracle import cpu_oracle as O
from tests.helpers import make_cloud_cn

L = _lib.lib()
B, C, N, k = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (2, 64, 1024, 20)
vp = ctypes.c_void_p
L.fpsg_tune_knn_tc_scores.restype = ctypes.c_int
L.fpsg_tune_knn_tc_scores.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_size_t, vp]
L.fpsg_tune_knn_tc_mask_offset.restype = ctypes.c_size_t
L.fpsg_tune_knn_tc_mask_offset.argtypes = [ctypes.c_int] * 4

x = make_cloud_cn(B, C, N, seed=B + C + N)
xd = x.cuda()
nb = L.fpsg_knn_workspace_bytes(B, C, N, k)
ws = torch.zeros(nb, dtype=torch.uint8, device="cuda")
res = {"B": B, "C": C, "N": N, "k": k, "ws_bytes": nb}
st = torch.cuda.current_stream().cuda_stream

if B * N * N <= 64 << 20:
    scores = torch.full((B, N, N), float("nan"), device="cuda")
    _lib.check(L.fpsg_tune_knn_tc_scores(xd.data_ptr(), B, C, N, k, scores.data_ptr(), ws.data_ptr(), nb, st))
    torch.cuda.synchronize()
    s = scores.cpu().numpy().astype(np.float64)
    xn = x.numpy().astype(np.float64)
    xx32 = np.zeros((B, N), np.float32)
    for c in range(C):
        xx32 = (xx32 + (x.numpy()[:, c, :] * x.numpy()[:, c, :]).astype(np.float32)).astype(np.float32)
    want = np.einsum("bci,bcj->bij", xn, xn) - xx32.astype(np.float64)[:, None, :] / 2
    mx = xx32.max(axis=1)
    sc = np.array([2.0 ** (7 - np.frexp(np.sqrt(np.float32(m)))[1]) if m > 0 else 1.0 for m in mx])  # the operand scale of the kernel
    s = s / (sc * sc)[:, None, None]
    err = np.abs(s - want)
    E = 1.25 * 2.0 ** -10 * np.sqrt(xx32.astype(np.float64) * mx[:, None]) + 2.0 ** -16 * mx[:, None]
    res["scores_nan"] = int(np.isnan(s).sum())
    res["max_abs_err"] = float(np.nanmax(err))
    res["max_err_over_E"] = float(np.nanmax(err / E[:, :, None]))
    W = (N + 31) // 32
    moff = L.fpsg_tune_knn_tc_mask_offset(B, C, N, k)
    masks = ws[moff:moff + B * N * W * 4].view(torch.int32).cpu().numpy().view(np.uint32).reshape(B, N, W)
    bits = np.unpackbits(masks.view(np.uint8).reshape(B, N, W, 4), axis=-1, bitorder="little").reshape(B, N, W * 32)[:, :, :N]
    ref = O.knn_f32(x.numpy(), k)
    hit = np.take_along_axis(bits, ref, axis=2)
    res["topk_missing_from_masks"] = int((hit == 0).sum())
    cnt = bits.sum(axis=2)
    res["survivors_mean"] = float(cnt.mean())
    res["survivors_max"] = int(cnt.max())
    res["survivors_min"] = int(cnt.min())

idx = torch.empty(B, N, k, dtype=torch.int64, device="cuda")
_lib.check(L.fpsg_knn_fwd(xd.data_ptr(), B, C, N, k, idx.data_ptr(), ws.data_ptr(), nb, st))
torch.cuda.synchronize()
ref = O.knn_f32(x.numpy(), k)
got = idx.cpu().numpy()
res["knn_mismatches"] = int((got != ref).sum())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    L.fpsg_knn_fwd(xd.data_ptr(), B, C, N, k, idx.data_ptr(), ws.data_ptr(), nb, st)
e0.record()
for _ in range(20):
    L.fpsg_knn_fwd(xd.data_ptr(), B, C, N, k, idx.data_ptr(), ws.data_ptr(), nb, st)
e1.record()
torch.cuda.synchronize()
res["us_per_call"] = round(e0.elapsed_time(e1) / 20 * 1e3, 2)
print(json.dumps(res))
