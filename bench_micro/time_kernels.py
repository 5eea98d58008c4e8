"""Quick device-side timing of the library's kernels through the C ABI (development aid, not the bench)."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpsg_b200 import _lib
from tests.helpers import make_pair, make_cloud_cn

def time_fn(fn, iters=50, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

L = _lib.lib()
t = ctypes.c_double(); mhz = ctypes.c_double()
_lib.check(L.fpsg_probe_fp32_peak(ctypes.byref(t), ctypes.byref(mhz), None))
print(json.dumps({"fp32_peak_tflops": t.value, "sm_mhz": mhz.value}))
peak = t.value
CH = [(32, 2048, 2048), (64, 4096, 4096), (64, 8192, 8192), (16, 16384, 16384), (1184, 2048, 2048), (5, 2048, 2048)]
KN = [(32, 3, 1024, 20), (32, 64, 1024, 20), (32, 128, 1024, 20)]
if "--c1" in sys.argv: CH, KN = CH[:1], KN[:1]
for (B, N, M) in CH:
    p1, p2 = make_pair(B, N, M); p1 = p1.cuda(); p2 = p2.cuda()
    d1 = torch.empty(B, N, device="cuda"); d2 = torch.empty(B, M, device="cuda")
    i1 = torch.empty(B, N, dtype=torch.int64, device="cuda"); i2 = torch.empty(B, M, dtype=torch.int64, device="cuda")
    loss = torch.empty(B, device="cuda"); gl = torch.ones(B, device="cuda"); gp1 = torch.empty_like(p1)
    nb = L.fpsg_chamfer_fwd_workspace_bytes(B, N, M); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    fwd = lambda: _lib.check(L.fpsg_chamfer_loss_fwd(p1.data_ptr(), p2.data_ptr(), B, N, M, 1.0, 1.0, loss.data_ptr(), d1.data_ptr(), d2.data_ptr(), i1.data_ptr(), i2.data_ptr(), ws.data_ptr(), nb, s))
    bwd = lambda: _lib.check(L.fpsg_chamfer_loss_bwd(p1.data_ptr(), p2.data_ptr(), i1.data_ptr(), i2.data_ptr(), gl.data_ptr(), 1.0, 1.0, B, N, M, gp1.data_ptr(), None, None, 0, s))
    it = 50 if B * N * M < 5e9 else 5
    tf = time_fn(fwd, it, 5); tb = time_fn(bwd, it, 5)
    flops = 16.0 * B * N * M
    print(json.dumps({"B": B, "N": N, "M": M, "fwd_us": round(tf, 2), "bwd_us": round(tb, 2), "fwd_tflops": round(flops / tf / 1e6, 2),
                      "frac_of_probe_peak": round(flops / tf / 1e6 / peak, 3), "pairs_per_s_fwd_bwd": round(B / (tf + tb) * 1e6), "ws_MB": round(nb / 1e6, 2)}))
for (B, C, N, k) in KN:
    x = make_cloud_cn(B, C, N).cuda(); idx = torch.empty(B, N, k, dtype=torch.int64, device="cuda")
    nb = L.fpsg_knn_workspace_bytes(B, C, N, k); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    f = lambda: _lib.check(L.fpsg_knn_fwd(x.data_ptr(), B, C, N, k, idx.data_ptr(), ws.data_ptr(), nb, s))
    def torch_knn():  # the reference's tensor expressions (src/dgcnn/model.py:14-18) on the same GPU, for scale
        inner = -2 * torch.matmul(x.transpose(2, 1), x)
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        return (-xx - inner - xx.transpose(2, 1)).topk(k=k, dim=-1)[1]
    print(json.dumps({"knn": [B, C, N, k], "us": round(time_fn(f, 20, 3), 2), "torch_reference_ops_us": round(time_fn(torch_knn, 10, 2), 2)}))

import fpsg_b200
for (B, C, N, k) in ([(32, 64, 1024, 20)] if "--c1" in sys.argv else [(32, 3, 1024, 20), (32, 64, 1024, 20), (32, 128, 1024, 20)]):
    x = make_cloud_cn(B, C, N).cuda(); idx = fpsg_b200.knn(x, k)
    out = torch.empty(B, 2 * C, N, k, device="cuda"); gx = torch.empty(B, C, N, device="cuda")
    nb = L.fpsg_edge_feature_bwd_workspace_bytes(B, C, N, k); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    f = lambda: _lib.check(L.fpsg_edge_feature_fwd(x.data_ptr(), idx.data_ptr(), B, C, N, k, out.data_ptr(), s))
    g = lambda: _lib.check(L.fpsg_edge_feature_bwd(out.data_ptr(), idx.data_ptr(), B, C, N, k, gx.data_ptr(), ws.data_ptr(), nb, s))
    tf, tb = time_fn(f, 20, 3), time_fn(g, 20, 3)
    nbytes = 8.0 * C * N * k * B
    def torch_ops():  # the reference's tensor expressions (src/dgcnn/model.py:31-41) on the same GPU, for scale
        base = torch.arange(0, B, device=x.device).view(-1, 1, 1) * N
        flat = (idx + base).view(-1)
        xt = x.transpose(2, 1).contiguous()
        feat = xt.view(B * N, -1)[flat, :].view(B, N, k, C)
        ctr = xt.view(B, N, 1, C).repeat(1, 1, k, 1)
        return torch.cat((feat - ctr, ctr), dim=3).permute(0, 3, 1, 2).contiguous()
    tt = time_fn(torch_ops, 10, 2)
    print(json.dumps({"edge_feature": [B, C, N, k], "fwd_us": round(tf, 1), "fwd_GBs": round(nbytes / tf / 1e3), "bwd_us": round(tb, 1), "bwd_GBs": round(nbytes / tb / 1e3), "torch_reference_ops_fwd_us": round(tt, 1)}))
