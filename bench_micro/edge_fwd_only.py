"""EdgeConv forward alone on one shape (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fpsg_b200
from fpsg_b200 import _lib
from tests.helpers import make_cloud_cn
L = _lib.lib()
B, C, N, k = (32, 64, 1024, 20) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(","))
x = make_cloud_cn(B, C, N).cuda(); idx = fpsg_b200.knn(x, k)
out = torch.empty(B, 2 * C, N, k, device="cuda")
s = torch.cuda.current_stream().cuda_stream
g = lambda: _lib.check(L.fpsg_edge_feature_fwd(x.data_ptr(), idx.data_ptr(), B, C, N, k, out.data_ptr(), s))
for _ in range(3): g()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): g()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
print(json.dumps({"edge_fwd": [B, C, N, k], "us": round(us, 1), "tb_per_s_written": round(out.numel() * 4 / us / 1e6, 2)}))
