"""EdgeConv backward alone on one shape (development aid for ncu)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fpsg_b200
from fpsg_b200 import _lib
from tests.helpers import make_cloud_cn
L = _lib.lib()
B, C, N, k = (32, 64, 1024, 20) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(","))
x = make_cloud_cn(B, C, N).cuda(); idx = fpsg_b200.knn(x, k)
out = torch.randn(B, 2 * C, N, k, device="cuda"); gx = torch.empty(B, C, N, device="cuda")
nb = L.fpsg_edge_feature_bwd_workspace_bytes(B, C, N, k); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
g = lambda: _lib.check(L.fpsg_edge_feature_bwd(out.data_ptr(), idx.data_ptr(), B, C, N, k, gx.data_ptr(), ws.data_ptr(), nb, s))
for _ in range(3): g()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): g()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"edge_bwd": [B, C, N, k], "us": round(e0.elapsed_time(e1) / 10 * 1e3, 1)}))
