"""Where the end-to-end step's host time goes (development aid): cProfile of the shim path."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fpsg_b200.shims
fpsg_b200.shims.install(force=True)
from kaolin.metrics.pointcloud import chamfer_distance
from bench import make_inputs
dev = torch.device("cuda", 0)
B, N, M = 32, 2048, 2048
host = [make_inputs(B, N, M, seed=s, pin=True) for s in range(4)]
d1 = torch.empty(B, N, 3, device=dev); d2 = torch.empty(B, M, 3, device=dev)
loss_host = torch.empty((), dtype=torch.float32).pin_memory()
def step(i, sync=True):
    h1, h2 = host[i % 4]
    d1.copy_(h1, non_blocking=True); d2.copy_(h2, non_blocking=True)
    a = d1.requires_grad_(True)
    val = chamfer_distance(a, d2).sum()
    val.backward()
    loss_host.copy_(val.detach(), non_blocking=True)
    if sync: torch.cuda.current_stream().synchronize()
    a.grad = None; d1.requires_grad_(False)
for i in range(20): step(i)
for sync in (True, False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(300): step(i, sync)
    torch.cuda.synchronize(); print("sync" if sync else "nosync", (time.perf_counter() - t0) / 300 * 1e6, "us/step")
pr = cProfile.Profile(); pr.enable()
for i in range(300): step(i)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
