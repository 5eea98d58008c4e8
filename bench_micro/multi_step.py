"""One c2 episode (5 query + 32 support pairs, N=M=2048): the two calls the reference issues vs. one multi-segment
call, device time per forward+backward through the C ABI (development aid)."""
import sys, os, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpsg_b200 import _lib
from tests.helpers import make_pair
L = _lib.lib()
N = M = 2048
dev = "cuda"
def bufs(B):
    p1, p2 = make_pair(B, N, M, seed=B); p1 = p1.cuda(); p2 = p2.cuda()
    return dict(B=B, p1=p1, p2=p2, i1=torch.empty(B, N, dtype=torch.int64, device=dev), i2=torch.empty(B, M, dtype=torch.int64, device=dev),
                loss=torch.empty(B, device=dev), gl=torch.ones(B, device=dev), gp1=torch.empty_like(p1),
                ws=torch.empty(L.fpsg_chamfer_fwd_workspace_bytes(B, N, M), dtype=torch.uint8, device=dev))
q, s_ = bufs(5), bufs(32)
st = torch.cuda.current_stream().cuda_stream
def single(d):
    _lib.check(L.fpsg_chamfer_loss_fwd(d["p1"].data_ptr(), d["p2"].data_ptr(), d["B"], N, M, 1.0, 1.0, d["loss"].data_ptr(), None, None, d["i1"].data_ptr(), d["i2"].data_ptr(), d["ws"].data_ptr(), d["ws"].numel(), st))
    _lib.check(L.fpsg_chamfer_loss_bwd(d["p1"].data_ptr(), d["p2"].data_ptr(), d["i1"].data_ptr(), d["i2"].data_ptr(), d["gl"].data_ptr(), 1.0, 1.0, d["B"], N, M, d["gp1"].data_ptr(), None, None, 0, st))
def two_calls():
    single(q); single(s_)
B = 37
segs = (_lib.ChamferSegment * 2)()
for k, d in enumerate((q, s_)):
    segs[k].p1, segs[k].p2, segs[k].gp1, segs[k].gp2, segs[k].batch, segs[k].weight = d["p1"].data_ptr(), d["p2"].data_ptr(), d["gp1"].data_ptr(), None, d["B"], 1.0
i1 = torch.empty(B, N, dtype=torch.int64, device=dev); i2 = torch.empty(B, M, dtype=torch.int64, device=dev)
loss = torch.empty(B, device=dev); totals = torch.empty(3, device=dev); g = torch.ones((), device=dev)
ws = torch.empty(L.fpsg_chamfer_fwd_workspace_bytes(B, N, M), dtype=torch.uint8, device=dev)
def multi():
    _lib.check(L.fpsg_chamfer_multi_loss_fwd(segs, 2, N, M, 1.0, 1.0, loss.data_ptr(), totals.data_ptr(), i1.data_ptr(), i2.data_ptr(), ws.data_ptr(), ws.numel(), st))
    _lib.check(L.fpsg_chamfer_multi_loss_bwd(segs, 2, i1.data_ptr(), i2.data_ptr(), g.data_ptr(), 1.0, 1.0, N, M, st))
def t(fn, it=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
print(json.dumps({"episode": "5 query + 32 support pairs, N=M=2048", "two_calls_us": round(t(two_calls), 2), "multi_segment_us": round(t(multi), 2)}))
