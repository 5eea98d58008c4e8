"""e2e rate of the native host-buffer pipeline for several depths (= lanes), host and device submits (development aid)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fpsg_b200
from fpsg_b200 import _lib, build
if os.environ.get("FPSG_TUNING_LIB"):   # the development build, so that the FPSG_* switches take effect
    _lib.use_library(build.LIB_TUNING)
from bench import make_inputs
B, N, M = 32, 2048, 2048
dev = torch.device("cuda", 0)
host = [make_inputs(B, N, M, seed=777 + s, pin=True) for s in range(8)]
devb = [(a.to(dev), b.to(dev)) for a, b in host]
res = {}
for depth in (4, 6):
    for mode in ("host", "device"):
        pipe = fpsg_b200.ChamferPipeline(B, N, M, device=dev, depth=depth)
        def loop(n):
            sub = 0
            for i in range(n):
                while sub < n and pipe.in_flight < depth:
                    if mode == "host":
                        pipe.submit(*host[sub % 8])
                    else:
                        pipe.submit_device(*devb[sub % 8])
                    sub += 1
                pipe.result()
        loop(200)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loop(1000)
        dt = time.perf_counter() - t0
        res[f"depth{depth}_{mode}_kpairs_per_s"] = round(B * 1000 / dt / 1e3, 1)
        pipe.close()
print(json.dumps(res))
