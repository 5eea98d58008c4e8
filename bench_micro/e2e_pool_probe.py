"""Does PyTorch's stream pool (32 + 32 streams created on the first torch.cuda.Stream()) disturb the pipeline's lanes? e2e rate of
the native pipeline (depth 4, host submits) before and after the pool exists, several repetitions."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fpsg_b200
from bench import make_inputs
B, N, M = 32, 2048, 2048
dev = torch.device("cuda", 0)
host = [make_inputs(B, N, M, seed=777 + s, pin=True) for s in range(8)]


def rate(depth=4, n=600):
    pipe = fpsg_b200.ChamferPipeline(B, N, M, device=dev, depth=depth)
    def loop(k):
        sub = 0
        for i in range(k):
            while sub < k and pipe.in_flight < depth:
                pipe.submit(*host[sub % 8]); sub += 1
            pipe.result()
    loop(20); torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(n); dt = time.perf_counter() - t0
    pipe.close()
    return round(B * n / dt / 1e3, 1)


res = {"CUDA_DEVICE_MAX_CONNECTIONS": os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), "before_pool": [rate() for _ in range(4)]}
streams = [torch.cuda.Stream(device=dev) for _ in range(3)]   # initialises torch's stream pools
x = torch.zeros(1, device=dev)
for s in streams:
    with torch.cuda.stream(s):
        x = x + 1
torch.cuda.synchronize()
res["after_pool"] = [rate() for _ in range(6)]
print(json.dumps(res))
