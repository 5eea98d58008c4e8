"""GPU parity tests for the fused loss forward + backward (`fpsg_chamfer_loss_fwd_bwd`, the cluster "finish" kernel):
against the CPU oracle, and against the separate forward / backward calls whose results it must reproduce."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_close_rel, make_pair

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import cpu_oracle
    return cpu_oracle


@pytest.fixture(autouse=True)
def _diag():
    from fpsg_b200 import chamfer
    chamfer.diag_keep_workspaces(True)
    yield
    chamfer.diag_keep_workspaces(False)


def _oracle_step(p1, p2, w1, w2, gloss):
    O = _oracle()
    B, N, M = p1.shape[0], p1.shape[1], p2.shape[1]
    od1, od2, oi1, oi2, _ = O.chamfer_fwd(p1.numpy(), p2.numpy())
    loss = (np.float64(w1) * od1.astype(np.float64).mean(1) + np.float64(w2) * od2.astype(np.float64).mean(1))
    gl = np.ones(B, np.float32) if gloss is None else gloss.numpy().astype(np.float32)
    g1 = np.repeat(((gl * np.float32(w1)) / np.float32(N))[:, None], N, 1).astype(np.float32)
    g2 = np.repeat(((gl * np.float32(w2)) / np.float32(M))[:, None], M, 1).astype(np.float32)
    gp1, gp2 = O.chamfer_bwd(p1.numpy(), p2.numpy(), oi1, oi2, g1, g2)
    return loss, oi1, oi2, gp1, gp2


SHAPES = [(1, 1, 1), (2, 1, 70), (2, 70, 1), (3, 31, 33), (2, 257, 64), (1, 513, 129), (2, 2047, 2049), (37, 100, 90), (5, 2048, 2048),
          (32, 2048, 2048), (1, 4096, 4096), (2, 4095, 300), (2, 300, 4096),
          (1, 4097, 500), (1, 5000, 5000)]  # the last two take the three-kernel chain behind the same entry point


@pytest.mark.parametrize("B,N,M", SHAPES)
def test_fused_step_vs_oracle(B, N, M):
    import fpsg_b200
    from fpsg_b200 import chamfer
    p1, p2 = make_pair(B, N, M, seed=4321 + N + M)
    g = torch.Generator().manual_seed(B)
    gloss = torch.randn(B, generator=g) if (N + M) % 2 else None
    w1, w2 = (1.0, 1.0) if B % 2 else (0.75, 1.5)
    loss, gp1, gp2, i1, i2 = fpsg_b200.chamfer_loss_and_grad(p1.cuda(), p2.cuda(), w1, w2, gloss.cuda() if gloss is not None else None,
                                                             need_gp1=True, need_gp2=True, return_indices=True)
    torch.cuda.synchronize()
    assert chamfer.diag_mismatches() == 0
    oloss, oi1, oi2, ogp1, ogp2 = _oracle_step(p1, p2, w1, w2, gloss)
    assert (i1.cpu().numpy() == oi1).all() and (i2.cpu().numpy() == oi2).all()
    assert_close_rel(loss.cpu().numpy(), oloss, 1e-6, "loss")
    assert_close_rel(gp1.cpu().numpy(), ogp1, 1e-5, "grad_p1")
    assert_close_rel(gp2.cpu().numpy(), ogp2, 1e-5, "grad_p2")


@pytest.mark.parametrize("B,N,M", [(4, 600, 500), (32, 2048, 2048), (3, 4096, 1000)])
def test_fused_step_equals_separate_calls(B, N, M):
    import fpsg_b200
    p1, p2 = make_pair(B, N, M, seed=99 + N)
    a = p1.cuda().requires_grad_(True)
    b = p2.cuda().requires_grad_(True)
    ref = fpsg_b200.chamfer_loss(a, b, 1.0, 0.5)
    ref.sum().backward()
    loss, gp1, gp2 = fpsg_b200.chamfer_loss_and_grad(p1.cuda(), p2.cuda(), 1.0, 0.5, need_gp2=True)
    assert_close_rel(loss.cpu().numpy(), ref.detach().cpu().numpy(), 2e-7, "loss vs separate calls")
    # same edge terms, double accumulation in a different (arbitrary) order: equal to the last bit in practice
    assert_close_rel(gp1.cpu().numpy(), a.grad.cpu().numpy(), 1e-6, "grad_p1 vs separate calls")
    assert_close_rel(gp2.cpu().numpy(), b.grad.cpu().numpy(), 1e-6, "grad_p2 vs separate calls")


def test_fused_step_hub_contention():
    """Clustered p1 (every p2 point is nearest to one of a handful of p1 points): buckets overflow into the spill lists."""
    import fpsg_b200
    from fpsg_b200 import chamfer
    g = torch.Generator().manual_seed(11)
    p2 = torch.randn(3, 2048, 3, generator=g)
    p1 = torch.zeros(3, 2048, 3)
    p1[0] = 1e-3 * torch.randn(2048, 3, generator=g)          # a tight blob: its hull points become hubs
    p1[1, :] = torch.tensor([0.1, -0.2, 0.3])                 # all identical: one hub of 2048 edges (lowest index wins)
    p1[2] = torch.randn(2048, 3, generator=g)
    p1[2, 5:] = p1[2, :1]                                     # six distinct points
    loss, gp1, gp2, i1, i2 = fpsg_b200.chamfer_loss_and_grad(p1.cuda(), p2.cuda(), need_gp2=True, return_indices=True)
    torch.cuda.synchronize()
    assert chamfer.diag_mismatches() == 0
    oloss, oi1, oi2, ogp1, ogp2 = _oracle_step(p1, p2, 1.0, 1.0, None)
    assert (i1.cpu().numpy() == oi1).all() and (i2.cpu().numpy() == oi2).all()
    assert_close_rel(loss.cpu().numpy(), oloss, 1e-6, "loss")
    assert_close_rel(gp1.cpu().numpy(), ogp1, 1e-5, "grad_p1")
    assert_close_rel(gp2.cpu().numpy(), ogp2, 1e-5, "grad_p2")


def test_fused_multi_segment_equals_autograd_path():
    import fpsg_b200
    q1, q2 = make_pair(5, 2048, 2048, seed=1)
    s1, s2 = make_pair(32, 2048, 2048, seed=2)
    a, b = q1.cuda().requires_grad_(True), s1.cuda().requires_grad_(True)
    total, sums, per_pair = fpsg_b200.chamfer_loss_multi([(a, q2.cuda()), (b, s2.cuda())], [1.0, 0.25])
    total.backward()
    t2, su2, pp2, grads = fpsg_b200.chamfer_loss_multi_and_grad([(q1.cuda(), q2.cuda()), (s1.cuda(), s2.cuda())], [1.0, 0.25])
    assert_close_rel(pp2.cpu().numpy(), per_pair.detach().cpu().numpy(), 2e-7, "per-pair loss")
    assert_close_rel(su2.cpu().numpy(), sums.detach().cpu().numpy(), 1e-6, "segment sums")
    assert abs(float(t2) - float(total)) <= 1e-6 * abs(float(total))
    assert grads[0][1] is None and grads[1][1] is None
    assert_close_rel(grads[0][0].cpu().numpy(), a.grad.cpu().numpy(), 1e-6, "query grad")
    assert_close_rel(grads[1][0].cpu().numpy(), b.grad.cpu().numpy(), 1e-6, "support grad")


def test_fused_step_nonfinite_inputs_follow_the_strict_scan():
    """NaN / Inf coordinates: a NaN distance is skipped, a point without any finite distance gets +INF and index 0
    (include/fpsg_b200.h) -- the oracle's strict-'<' scan from +INF does the same."""
    import fpsg_b200
    O = _oracle()
    p1, p2 = make_pair(2, 300, 260, seed=3)
    p1[0, 7] = float("nan")
    p1[1, 100, 1] = float("inf")
    p2[0, 33, 2] = float("nan")
    p2[1, 0] = float("-inf")
    od1, od2, oi1, oi2, _ = O.chamfer_fwd(p1.numpy(), p2.numpy())
    d1, d2, i1, i2 = fpsg_b200.chamfer_distance(p1.cuda(), p2.cuda())
    assert (i1.cpu().numpy() == oi1).all() and (i2.cpu().numpy() == oi2).all()
    np.testing.assert_array_equal(d1.cpu().numpy(), od1)
    np.testing.assert_array_equal(d2.cpu().numpy(), od2)
    _, _, _, f1, f2 = fpsg_b200.chamfer_loss_and_grad(p1.cuda(), p2.cuda(), return_indices=True)
    assert (f1.cpu().numpy() == oi1).all() and (f2.cpu().numpy() == oi2).all()
