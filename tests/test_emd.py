"""Sinkhorn EMD (reference src/models/utils.py:9-13, src/models/few_shot.py:168). PARITY UNPINNED: the
reference's arithmetic lives in neuralnet_pytorch / geomloss, both absent. CPU tests pin the float64 oracle's definition to
an independent witness (exact optimal assignment by scipy as eps -> 0); GPU tests compare the kernel with the oracle."""
import numpy as np
import pytest
import torch

from tests.helpers import make_pair


def test_epsilon_schedule_matches_the_geomloss_formula():
    from fpsg_b200.emd import epsilon_schedule
    s = epsilon_schedule(2.0, 0.05, 0.5)
    want = [4.0] + [float(np.exp(e)) for e in np.arange(2 * np.log(2.0), 2 * np.log(0.05), 2 * np.log(0.5))] + [0.0025]
    assert np.allclose(s, want, rtol=1e-12) and s[0] == 4.0 and s[-1] == 0.05 ** 2
    assert all(a >= b for a, b in zip(s[1:], s[2:]))


def test_oracle_plain_ot_tends_to_the_optimal_assignment_cost():
    """Independent witness of the definition: for small eps the entropic dual value of two equal-size uniform clouds
    approaches the optimal assignment cost mean_i |x_i - y_sigma(i)|^2 / 2 (scipy's Hungarian solver shares no code)."""
    from scipy.optimize import linear_sum_assignment
    from oracle import emd_oracle as E
    rng = np.random.default_rng(0)
    x, y = rng.normal(size=(40, 3)) * 0.3, rng.normal(size=(40, 3)) * 0.3 + 0.1
    C = 0.5 * ((x[:, None] - y[None]) ** 2).sum(-1)
    r, c = linear_sum_assignment(C)
    exact = C[r, c].mean()
    sched = [1.0 * 0.7 ** t for t in range(40)] + [1e-4] * 400
    val, f, g = E.sinkhorn_divergence(x, y, sched, debias=False)
    assert abs(val - exact) <= 0.02 * exact + 1e-4 * np.log(40)  # entropic bias is O(eps log n)
    # dual feasibility up to the entropic slack: f_i + g_j <= C_ij + small
    assert (f[:, None] + g[None, :] - C).max() <= 5e-3


def test_oracle_divergence_is_zero_on_identical_clouds_and_positive_otherwise():
    from fpsg_b200.emd import epsilon_schedule
    from oracle import emd_oracle as E
    p1, p2 = make_pair(2, 64, 64, seed=3)
    sched = epsilon_schedule(2.0)
    same = E.emd_sinkhorn(p1.numpy(), p1.numpy(), sched)
    diff = E.emd_sinkhorn(p1.numpy(), p2.numpy(), sched)
    assert np.abs(same).max() <= 1e-12 and (diff > 1e-4).all()


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,M,debias", [(1, 1, 1, True), (2, 7, 130, True), (3, 257, 300, True), (2, 512, 512, False), (1, 2500, 700, True), (4, 2048, 2048, True)])
def test_gpu_emd_matches_the_float64_oracle(B, N, M, debias):
    import fpsg_b200
    from fpsg_b200.emd import epsilon_schedule
    from oracle import emd_oracle as E
    p1, p2 = make_pair(B, N, M, seed=77 + N)
    sched = epsilon_schedule(2.0, 0.05, 0.5)
    got, f, g = fpsg_b200.emd_sinkhorn(p1.cuda(), p2.cuda(), eps_schedule=sched, debias=debias, return_potentials=True)
    nb = min(B, 2)  # the float64 oracle builds four dense cost matrices per pair
    want = np.array([E.sinkhorn_divergence(p1[b].numpy(), p2[b].numpy(), sched, debias) for b in range(nb)], dtype=object)
    for b in range(nb):
        val, bi, aj = want[b]
        # potentials: 1e-5 relative to their scale (FP32 log-sum-exp of up to 2500 terms per row, ten steps)
        assert np.abs(f[b].cpu().numpy() - bi).max() <= 1e-5 * max(np.abs(bi).max(), 1e-3)
        assert np.abs(g[b].cpu().numpy() - aj).max() <= 1e-5 * max(np.abs(aj).max(), 1e-3)
        # the debiased value is a difference of nearly equal means: tolerance relative to the potentials' scale
        tol = 2e-5 * max(np.abs(bi).mean(), np.abs(aj).mean(), 1e-3) + 1e-4 * abs(val)
        assert abs(float(got[b]) - val) <= tol, (float(got[b]), val, tol)


@pytest.mark.gpu
def test_gpu_emd_default_schedule_and_edge_cases():
    import fpsg_b200
    from fpsg_b200.emd import epsilon_schedule, max_diameter
    from oracle import emd_oracle as E
    p1, p2 = make_pair(2, 300, 200, seed=5)
    a, b = p1.cuda(), p2.cuda()
    got = fpsg_b200.emd_sinkhorn(a, b)   # diameter from the data, blur 0.05, scaling 0.5: geomloss' defaults
    sched = epsilon_schedule(max_diameter(a, b))
    want = E.emd_sinkhorn(p1.numpy(), p2.numpy(), sched)
    assert np.abs(got.cpu().numpy() - want).max() <= 1e-4 * np.abs(want).max() + 2e-6
    same = fpsg_b200.emd_sinkhorn(a, a.clone())
    assert float(same.abs().max()) <= 2e-6          # identical clouds: zero divergence
    pt = torch.zeros(1, 50, 3).cuda()
    assert float(fpsg_b200.emd_sinkhorn(pt, pt).abs().max()) <= 1e-6   # all points coincide (diameter 0)
    assert fpsg_b200.emd_sinkhorn(a[:0], b[:0]).shape == (0,)
    with pytest.raises(RuntimeError):
        fpsg_b200.emd_sinkhorn(p1, p2)          # CPU tensors: no CPU path


@pytest.mark.gpu
def test_emd_wrapper_signature_through_the_neuralnet_pytorch_shim():
    """The reference's call, verbatim: `emd_loss(pc1, pc2, reduce='sum', sinkhorn=True)` then `.sum()`
    (src/models/utils.py:12-13, src/models/few_shot.py:168)."""
    import sys
    import fpsg_b200.shims
    saved = {n: m for n, m in sys.modules.items() if n.split(".")[0] == "neuralnet_pytorch"}
    for n in saved:
        del sys.modules[n]
    try:
        fpsg_b200.shims.install_neuralnet_pytorch(force=True)
        from neuralnet_pytorch.metrics import emd_loss
        p1, p2 = make_pair(3, 128, 128, seed=9)
        out = emd_loss(p1.cuda(), p2.cuda(), reduce='sum', sinkhorn=True)
        assert out.shape == (3,) and torch.isfinite(out).all() and float(out.sum()) > 0
        with pytest.raises(NotImplementedError):
            emd_loss(p1.cuda(), p2.cuda(), sinkhorn=False)
    finally:
        for n in list(sys.modules):
            if n.split(".")[0] == "neuralnet_pytorch":
                del sys.modules[n]
        sys.modules.update(saved)
