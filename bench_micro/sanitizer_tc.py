"""compute-sanitizer workload for this round's additions: the tensor-core kNN path (every eligible shape class, degenerate and
ragged clouds), the pipeline lanes with host and device submits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fpsg_b200
from tests.helpers import make_pair, make_cloud_cn
for (B, C, N, k) in [(2, 64, 256, 20), (1, 128, 300, 20), (1, 16, 511, 1), (1, 72, 700, 32), (2, 64, 1024, 20), (1, 32, 1100, 20), (1, 24, 2049, 5)]:
    x = make_cloud_cn(B, C, N, seed=C + N)
    fpsg_b200.knn(x.cuda(), k)
    fpsg_b200.knn(x[:, :, :1].expand(B, C, N).contiguous().cuda(), k)        # all points identical: streaming re-scoring
    fpsg_b200.knn((x * 1e6).cuda(), k)
B, N, M = 3, 300, 257
pipe = fpsg_b200.ChamferPipeline(B, N, M, depth=3)
batches = [make_pair(B, N, M, seed=s) for s in range(7)]
nxt = done = 0
while done < len(batches):
    while nxt < len(batches) and pipe.in_flight < 3:
        if nxt % 2:
            pipe.submit(*batches[nxt])
        else:
            pipe.submit_device(batches[nxt][0].cuda(), batches[nxt][1].cuda())
        nxt += 1
    pipe.result()
    done += 1
pipe.close()
torch.cuda.synchronize()
print("sanitizer workload (tensor-core kNN, lanes) done")
