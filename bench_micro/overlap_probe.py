"""Cross-step overlap probe (development aid): the fused c1 loss step (sweep + finish) issued round-robin on S streams,
each stream with its OWN workspace and outputs, so consecutive steps of independent batches may overlap on the device
(the finish kernel of step i is latency-bound and leaves most issue slots idle; the sweep of step i+1 can use them).
Within a stream the usual full dependency holds, so a workspace is never reused before its last reader has finished.
Usage: overlap_probe.py [B N M]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fpsg_b200 import _lib, build
if os.environ.get("FPSG_TUNING_LIB"):   # the development build, so that the FPSG_* switches take effect
    _lib.use_library(build.LIB_TUNING)
from bench import make_inputs

L = _lib.lib()
args = [int(v) for v in sys.argv[1:4]] if len(sys.argv) >= 4 else [32, 2048, 2048]
B, N, M = args
dev = torch.device("cuda", 0)
sets = max(4, min(96, int(160e6 / (B * (N + M) * 12))))
ring = [make_inputs(B, N, M, seed=1235 + s, device=dev) for s in range(sets)]
ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in ring]
nb = L.fpsg_chamfer_fwd_workspace_bytes(B, N, M)
MAXS = 8


class Slot:
    def __init__(self, prio=0):
        self.i1 = torch.empty(B, N, dtype=torch.int64, device=dev)
        self.i2 = torch.empty(B, M, dtype=torch.int64, device=dev)
        self.loss = torch.empty(B, device=dev)
        self.gp1 = torch.empty(B, N, 3, device=dev)
        self.ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        self.stream = torch.cuda.Stream(device=dev, priority=prio)
        self.sp = self.stream.cuda_stream

    def step(self, i):
        a, b = ptrs[i % sets]
        _lib.check(L.fpsg_chamfer_loss_fwd_bwd(a, b, B, N, M, 1.0, 1.0, None, self.loss.data_ptr(), None, None, self.i1.data_ptr(), self.i2.data_ptr(),
                                               self.gp1.data_ptr(), None, self.ws.data_ptr(), nb, self.sp))


slots = [Slot() for _ in range(MAXS)]
main = torch.cuda.current_stream(dev)


def run(ns, it):
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in slots[:ns]:
        s.stream.wait_event(e0)
    for i in range(it):
        slots[i % ns].step(i)
    for s in slots[:ns]:
        ev = torch.cuda.Event()
        ev.record(s.stream)
        main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


res = {"B": B, "N": N, "M": M, "env": {k: v for k, v in os.environ.items() if k.startswith("FPSG_")}}
it = 600 if B * N * M < 5e9 else 30
for ns in (1, 2, 3, 4, 6, 8):
    run(ns, 40)
    res[f"streams{ns}_us_per_step"] = round(min(run(ns, it) for _ in range(3)), 2)

# correctness of the overlapped issue: every slot's outputs of its LAST step equal a serial run of the same batch
ref = Slot()
ok = True
last = {}
for i in range(it):
    last[i % 3] = i
torch.cuda.synchronize()
for i in range(it):
    slots[i % 3].step(i)
torch.cuda.synchronize()
for k, i in last.items():
    ref.step(i)
    torch.cuda.synchronize()
    s = slots[k]
    ok &= bool(torch.equal(ref.i1, s.i1) and torch.equal(ref.i2, s.i2) and torch.equal(ref.loss, s.loss) and torch.equal(ref.gp1, s.gp1))
res["overlapped_outputs_equal_serial"] = ok
print(json.dumps(res))
