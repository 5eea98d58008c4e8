"""GPU tests of the multi-segment Chamfer loss: the query-set and support-set calls of the
reference's training loss (src/models/few_shot.py:109-122) and their weighted sum as one forward / one backward.
Gate: per-pair losses and gradients bit-identical to the single-call path (which the other tests pin to the oracle),
totals within 1e-6 relative of the composed float64 sums, and a direct comparison with the CPU oracle."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_close_rel, make_pair

pytestmark = pytest.mark.gpu


def _segments(spec, N, M, seed=7, grad_p2=False):
    out = []
    for s, B in enumerate(spec):
        p1, p2 = make_pair(B, N, M, seed=seed + 31 * s)
        out.append((p1.cuda().requires_grad_(True), p2.cuda().requires_grad_(grad_p2)))
    return out


@pytest.mark.parametrize("spec,N,M,weights", [
    ((5, 32), 2048, 2048, (1.0, 1.0)),          # one episode of config c2: 5 query + 32 support pairs
    ((5, 32), 2048, 2048, (0.75, 0.3)),
    ((3, 0, 4), 257, 300, (2.0, 5.0, 0.5)),     # an empty segment in the middle, ragged sizes
    ((1,), 70, 33, (1.5,)),
    ((1, 2, 1, 3, 1, 2, 1, 2), 100, 90, tuple(0.1 * (i + 1) for i in range(8))),
])
def test_multi_matches_separate_calls(spec, N, M, weights):
    import fpsg_b200
    segs = _segments(spec, N, M)
    total, sums, loss = fpsg_b200.chamfer_loss_multi(segs, weights)
    total.backward()
    got = [a.grad.clone() for a, _ in segs]
    assert total.shape == () and sums.shape == (len(spec),) and loss.shape == (sum(spec),)
    assert not sums.requires_grad and not loss.requires_grad
    # the same thing as the reference issues it: one call per segment, .sum(), weighted sum
    ref_losses, ref_total = [], 0.0
    for (a, b), w in zip(segs, weights):
        a.grad = None
        l = fpsg_b200.chamfer_loss(a, b)
        (w * l.sum()).backward()
        ref_losses.append(l.detach())
        ref_total += w * float(l.detach().double().sum())
    ref_loss = torch.cat(ref_losses)
    assert torch.equal(loss, ref_loss), "per-pair losses must be bit-identical to the single-call path"
    for s, ((a, _), g) in enumerate(zip(segs, got)):
        assert torch.equal(a.grad, g), f"segment {s}: gradient differs from the single-call path"
    off = 0
    for s, B in enumerate(spec):
        want = float(ref_loss[off:off + B].double().sum())
        assert abs(float(sums[s]) - want) <= 1e-6 * max(abs(want), 1e-30), f"sum of segment {s}"
        off += B
    assert abs(float(total) - ref_total) <= 2e-6 * max(abs(ref_total), 1e-30)


def test_multi_vs_oracle_with_p2_gradients():
    from oracle import cpu_oracle as O
    import fpsg_b200
    spec, N, M, weights = (2, 3), 200, 150, (0.5, 2.0)
    segs = _segments(spec, N, M, seed=11, grad_p2=True)
    total, sums, loss = fpsg_b200.chamfer_loss_multi(segs, weights, w1=1.0, w2=1.0)
    (3.0 * total).backward()
    want_total = 0.0
    for (a, b), w, s in zip(segs, weights, range(len(spec))):
        p1, p2 = a.detach().cpu().numpy(), b.detach().cpu().numpy()
        d1, d2, i1, i2, _ = O.chamfer_fwd(p1, p2)
        B = p1.shape[0]
        per_pair = d1.astype(np.float64).mean(1) + d2.astype(np.float64).mean(1)
        want_total += w * per_pair.sum()
        g1 = np.full((B, N), 3.0 * w / N, dtype=np.float32)
        g2 = np.full((B, M), 3.0 * w / M, dtype=np.float32)
        gp1, gp2 = O.chamfer_bwd(p1, p2, i1, i2, g1, g2)
        assert_close_rel(a.grad.cpu().numpy(), gp1, 1e-5, f"gp1 of segment {s}")
        assert_close_rel(b.grad.cpu().numpy(), gp2, 1e-5, f"gp2 of segment {s}")
    assert abs(float(total) - want_total) <= 1e-5 * abs(want_total)


def test_multi_error_behaviour_and_empty_batch():
    import fpsg_b200
    a, b = make_pair(2, 64, 64)
    with pytest.raises(RuntimeError):
        fpsg_b200.chamfer_loss_multi([(a.cuda(), b.cuda()), (a.cuda()[:, :32], b.cuda())])  # N differs between segments
    with pytest.raises(RuntimeError):
        fpsg_b200.chamfer_loss_multi([(a.cuda(), b.cuda())] * 9)
    with pytest.raises(RuntimeError):
        fpsg_b200.chamfer_loss_multi([(a, b)])  # CPU tensors: no fallback
    e = torch.empty(0, 64, 3, device="cuda")
    total, sums, loss = fpsg_b200.chamfer_loss_multi([(e, e)])
    assert float(total) == 0.0 and float(sums[0]) == 0.0 and loss.numel() == 0
