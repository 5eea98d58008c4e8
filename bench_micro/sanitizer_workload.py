import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fpsg_b200
from tests.helpers import make_pair, make_cloud_cn
for (B, N, M) in [(1, 1, 1), (2, 1, 70), (2, 70, 1), (3, 31, 33), (2, 513, 129), (1, 2049, 2047), (5, 2048, 2048), (1, 4100, 300), (2, 300, 4100)]:
    p1, p2 = make_pair(B, N, M, seed=N + M)
    a = p1.cuda().requires_grad_(True); b = p2.cuda().requires_grad_(True)
    d1, d2, i1, i2 = fpsg_b200.chamfer_distance(a, b)
    (d1.sum() + d2.sum()).backward()
    l = fpsg_b200.chamfer_loss(a, b, 0.5, 2.0); l.sum().backward()
    d, i = fpsg_b200.sided_distance(a, b); d.sum().backward()
    fpsg_b200.chamfer_loss_and_grad(a, b, 0.5, 2.0, need_gp2=True)   # fused finish kernel (cluster + DSMEM)
    fpsg_b200.chamfer_loss_and_grad(a, b)
for (B, C, N, k) in [(2, 3, 100, 20), (1, 3, 33, 33), (2, 3, 257, 64), (1, 3, 300, 128), (2, 64, 300, 20), (1, 128, 200, 20), (2, 5, 77, 7), (1, 3, 1500, 1), (1, 70, 100, 40)]:
    x = make_cloud_cn(B, C, N, seed=C + N)
    fpsg_b200.knn(x.cuda(), k)
# threshold-selection kernels: ragged N, both fine-group sizes, the overflow fallbacks (identical points), k = 1 and 32
for (B, C, N, k) in [(2, 3, 1024, 20), (1, 3, 1000, 32), (1, 3, 2500, 20), (1, 3, 65, 1), (2, 64, 1024, 20), (1, 16, 130, 32), (1, 8, 2048, 5)]:
    x = make_cloud_cn(B, C, N, seed=C + N)
    fpsg_b200.knn(x.cuda(), k)
    fpsg_b200.knn(x[:, :, :1].expand(B, C, N).contiguous().cuda(), k)
    fpsg_b200.get_graph_feature(x.cuda().requires_grad_(True), k=min(k, 8)).sum().backward()
# multi-segment loss and the input pipeline
segs = []
for s_, B in enumerate((3, 0, 2)):
    p1, p2 = make_pair(B, 257, 300, seed=s_)
    segs.append((p1.cuda().requires_grad_(True), p2.cuda().requires_grad_(True)))
total, sums, loss = fpsg_b200.chamfer_loss_multi(segs, [0.5, 1.0, 2.0]); total.backward()
fpsg_b200.chamfer_loss_multi_and_grad(segs, [0.5, 1.0, 2.0], need_grad=[(True, True), (True, False), (False, True)])
hub = torch.zeros(1, 700, 3).cuda()   # all p1 points identical: one bucket takes every edge (overflow list)
fpsg_b200.chamfer_loss_and_grad(hub, segs[0][1][:1, :, :].contiguous(), need_gp2=True)
# list kernels with several candidate tiles (the lists live in shared memory between tiles)
for (B, C, N, k) in [(1, 64, 700, 100), (1, 128, 300, 128), (1, 64, 600, 50)]:
    fpsg_b200.knn(make_cloud_cn(B, C, N, seed=C + N).cuda(), k)
for (B, N) in [(1, 1), (3, 2049), (2, 5000)]:
    fpsg_b200.normalize_point_clouds(torch.randn(B, N, 3).cuda())
torch.cuda.synchronize()
print("sanitizer workload done")
