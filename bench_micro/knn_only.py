"""Runs the kNN C ABI a few times on the c3 shape (development aid for ncu captures and quick timings)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpsg_b200 import _lib
from tests.helpers import make_cloud_cn
L = _lib.lib()
shapes = [(32, 3, 1024, 20)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (B, C, N, k) in shapes:
    x = make_cloud_cn(B, C, N).cuda(); idx = torch.empty(B, N, k, dtype=torch.int64, device="cuda")
    nb = L.fpsg_knn_workspace_bytes(B, C, N, k); ws = torch.empty(max(nb, 1), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    f = lambda: _lib.check(L.fpsg_knn_fwd(x.data_ptr(), B, C, N, k, idx.data_ptr(), ws.data_ptr(), nb, s))
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"knn": [B, C, N, k], "us": round(e0.elapsed_time(e1) / 20 * 1e3, 2)}))
