"""kNN at c3 (B=32, C=3, N=1024, k=20), a few calls: the target of an ncu capture and a CUDA-event timing (development aid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpsg_b200 import _lib
L = _lib.lib()
B, C, N, k = 32, 3, int(os.environ.get("KNN_N", "1024")), 20
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(5)
x = torch.randn(B, C, N, generator=g).to(dev)
idx = torch.empty(B, N, k, dtype=torch.int64, device=dev)
nb = L.fpsg_knn_workspace_bytes(B, C, N, k)
ws = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
sp = torch.cuda.current_stream(dev).cuda_stream
def call():
    _lib.check(L.fpsg_knn_fwd(x.data_ptr(), B, C, N, k, idx.data_ptr(), ws.data_ptr(), nb, sp))
for _ in range(5):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    e0.record()
    for _ in range(50):
        call()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 50 * 1e3)
print(json.dumps({"N": N, "us_per_call": round(min(ts), 2), "checksum": int(idx.sum().item())}))
