"""Input pipeline timings (development aid): C parser vs the reference's per-line Python reader (restated), and the
batched GPU normalisation vs numpy per item."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fpsg_b200
from oracle import cpu_oracle as O
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cloud_b.ply")
def best(fn, n=20):
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return min(t) * 1e6
print(json.dumps({"ply_2048_vertices": {"c_parser_us": round(best(lambda: fpsg_b200.ply_reader(path)), 1),
                                        "python_reader_us": round(best(lambda: O.ply_reader_py(path), 5), 1)}}))
B, N = 37, 2048
x = torch.randn(B, N, 3) * 3 + 1
xc = x.cuda()
for _ in range(5): fpsg_b200.normalize_point_clouds(xc)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): fpsg_b200.normalize_point_clouds(xc)
e1.record(); torch.cuda.synchronize()
xn = x.numpy()
def numpy_items():
    for b in range(B):
        p = xn[b]; p = p - np.expand_dims(np.mean(p, axis=0), 0); p = p / np.max(np.sqrt(np.sum(p ** 2, axis=1)), 0)
print(json.dumps({"normalise_37x2048": {"gpu_kernel_us": round(e0.elapsed_time(e1) / 50 * 1e3, 1), "numpy_per_item_total_us": round(best(numpy_items, 5), 1)}}))
