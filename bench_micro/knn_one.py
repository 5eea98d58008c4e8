"""Runs knn for one (B, C, N, k) a few times (profiling target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fpsg_b200
from tests.helpers import make_cloud_cn
B, C, N, k = (int(v) for v in sys.argv[1:5])
x = make_cloud_cn(B, C, N).cuda()
for _ in range(5): fpsg_b200.knn(x, k)
torch.cuda.synchronize()
