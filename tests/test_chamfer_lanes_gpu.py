"""GPU test of the lane pattern (bench.py `value`): independent loss steps issued round-robin over several
streams, each with its own workspace and outputs, overlap on the device -- and give exactly the results of a serial run
(bit-identical losses, indices and gradients; one batch also against the CPU oracle). The C ABI is stateless, so this is a
property of the caller's stream / workspace discipline that the library must support: no hidden shared state between calls."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_close_rel, make_pair

pytestmark = pytest.mark.gpu


class _Lane:
    def __init__(self, L, B, N, M, dev, own_stream=True):
        self.nb = int(L.fpsg_chamfer_fwd_workspace_bytes(B, N, M))
        self.ws = torch.empty(self.nb, dtype=torch.uint8, device=dev)
        self.i1 = torch.empty(B, N, dtype=torch.int64, device=dev)
        self.i2 = torch.empty(B, M, dtype=torch.int64, device=dev)
        self.loss = torch.empty(B, device=dev)
        self.gp1 = torch.empty(B, N, 3, device=dev)
        self.stream = torch.cuda.Stream(device=dev) if own_stream else torch.cuda.current_stream(dev)


def _step(_lib, L, lane, p1, p2, B, N, M):
    _lib.check(L.fpsg_chamfer_loss_fwd_bwd(p1.data_ptr(), p2.data_ptr(), B, N, M, 1.0, 1.0, None, lane.loss.data_ptr(), None, None,
                                           lane.i1.data_ptr(), lane.i2.data_ptr(), lane.gp1.data_ptr(), None, lane.ws.data_ptr(), lane.nb,
                                           lane.stream.cuda_stream))


@pytest.mark.parametrize("B,N,M,lanes", [(32, 2048, 2048, 3), (5, 700, 513, 2), (37, 1024, 2048, 4)])
def test_steps_over_lanes_equal_serial_steps(B, N, M, lanes):
    from fpsg_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    nsteps = 4 * lanes + 1
    batches = [tuple(t.to(dev) for t in make_pair(B, N, M, seed=900 + s)) for s in range(nsteps)]
    torch.cuda.synchronize(dev)
    ls = [_Lane(L, B, N, M, dev) for _ in range(lanes)]
    got = {}
    for s, (p1, p2) in enumerate(batches):
        ln = ls[s % lanes]
        _step(_lib, L, ln, p1, p2, B, N, M)
        if s >= nsteps - lanes:  # the last step of every lane: its outputs survive until the end
            got[s] = ln
    torch.cuda.synchronize(dev)
    ref = _Lane(L, B, N, M, dev, own_stream=False)
    for s, ln in got.items():
        _step(_lib, L, ref, batches[s][0], batches[s][1], B, N, M)
        torch.cuda.synchronize(dev)
        assert torch.equal(ref.i1, ln.i1) and torch.equal(ref.i2, ln.i2), f"step {s}: indices differ between the overlapped and the serial run"
        assert torch.equal(ref.loss, ln.loss) and torch.equal(ref.gp1, ln.gp1), f"step {s}: loss / gradient differ"
    # one of them against the oracle as well
    from oracle import cpu_oracle as O
    s = nsteps - 1
    p1, p2 = batches[s][0].cpu().numpy(), batches[s][1].cpu().numpy()
    od1, od2, oi1, oi2, oloss = O.chamfer_fwd(p1, p2)
    assert (got[s].i1.cpu().numpy() == oi1).all() and (got[s].i2.cpu().numpy() == oi2).all()
    assert_close_rel(got[s].loss.cpu().numpy(), oloss, 1e-6, "per-pair loss")
    g1 = np.full((B, N), np.float32(1.0) / np.float32(N), np.float32)
    g2 = np.full((B, M), np.float32(1.0) / np.float32(M), np.float32)
    ogp1, _ = O.chamfer_bwd(p1, p2, oi1, oi2, g1, g2, True, False)
    assert_close_rel(got[s].gp1.cpu().numpy(), ogp1, 1e-5, "grad_p1")
