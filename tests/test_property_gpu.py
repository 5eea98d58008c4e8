"""Randomised parity tests (hypothesis): arbitrary small shapes and adversarial value distributions through the CUDA
path vs the CPU oracle. Complements the fixed edge cases of test_chamfer_gpu.py / test_knn_gpu.py: ragged sizes around
every tile boundary, exact ties (integer grids, duplicated points), large magnitudes."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from tests.helpers import assert_close_rel

pytestmark = pytest.mark.gpu

KINDS = ["normal", "grid", "dup", "far"]


def _cloud(rng, B, N, kind):
    if kind == "grid":      # integer grid: many exactly equal distances
        return rng.integers(-3, 4, size=(B, N, 3)).astype(np.float32)
    if kind == "dup":       # a handful of distinct points repeated
        base = rng.standard_normal((B, max(1, N // 7 + 1), 3)).astype(np.float32)
        return np.take_along_axis(base, rng.integers(0, base.shape[1], size=(B, N, 1)).repeat(3, axis=2), axis=1)
    x = rng.standard_normal((B, N, 3)).astype(np.float32)
    return x * 50.0 + 100.0 if kind == "far" else x


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(B=st.integers(1, 4), N=st.integers(1, 700), M=st.integers(1, 700), kind1=st.sampled_from(KINDS),
       kind2=st.sampled_from(KINDS), seed=st.integers(0, 2 ** 20))
def test_chamfer_random_shapes_match_oracle(B, N, M, kind1, kind2, seed):
    from oracle import cpu_oracle as O
    import fpsg_b200
    rng = np.random.default_rng(seed)
    p1, p2 = _cloud(rng, B, N, kind1), _cloud(rng, B, M, kind2)
    a = torch.from_numpy(p1).cuda().requires_grad_(True)
    b = torch.from_numpy(p2).cuda().requires_grad_(True)
    d1, d2, i1, i2 = fpsg_b200.chamfer_distance(a, b)
    g1 = torch.from_numpy(rng.standard_normal((B, N)).astype(np.float32)).cuda()
    g2 = torch.from_numpy(rng.standard_normal((B, M)).astype(np.float32)).cuda()
    ((d1 * g1).sum() + (d2 * g2).sum()).backward()
    od1, od2, oi1, oi2, oloss = O.chamfer_fwd(p1, p2)
    assert (i1.cpu().numpy() == oi1).all() and (i2.cpu().numpy() == oi2).all()
    assert (d1.detach().cpu().numpy() == od1).all() and (d2.detach().cpu().numpy() == od2).all()
    ogp1, ogp2 = O.chamfer_bwd(p1, p2, oi1, oi2, g1.cpu().numpy(), g2.cpu().numpy())
    assert_close_rel(a.grad.cpu().numpy(), ogp1, 1e-5, "grad_p1")
    assert_close_rel(b.grad.cpu().numpy(), ogp2, 1e-5, "grad_p2")
    loss = fpsg_b200.chamfer_loss(a.detach(), b.detach())
    assert_close_rel(loss.cpu().numpy(), oloss, 1e-6, "loss")


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(B=st.integers(1, 3), C=st.sampled_from([1, 2, 3, 3, 3, 5, 16, 64]), N=st.integers(1, 600), kfrac=st.floats(0.0, 1.0),
       kind=st.sampled_from(["normal", "grid", "dup"]), seed=st.integers(0, 2 ** 20))
def test_knn_random_shapes_match_oracle(B, C, N, kfrac, kind, seed):
    from oracle import cpu_oracle as O
    import fpsg_b200
    rng = np.random.default_rng(seed)
    k = max(1, min(N, 128, int(round(kfrac * min(N, 40))) or 1))
    if kind == "grid":
        x = rng.integers(-2, 3, size=(B, C, N)).astype(np.float32)
    elif kind == "dup":
        base = rng.standard_normal((B, C, max(1, N // 5 + 1))).astype(np.float32)
        x = np.take_along_axis(base, rng.integers(0, base.shape[2], size=(B, 1, N)).repeat(C, axis=1), axis=2)
    else:
        x = rng.standard_normal((B, C, N)).astype(np.float32)
    got = fpsg_b200.knn(torch.from_numpy(np.ascontiguousarray(x)).cuda(), k).cpu().numpy()
    ref = O.knn_f32(np.ascontiguousarray(x), k)
    assert (got == ref).all(), f"B={B} C={C} N={N} k={k} kind={kind}: {(got != ref).sum()} positions differ"


TC_KINDS = ["normal", "grid", "dup", "offset", "scaled", "relu", "clusters", "lowrank"]


def _features(rng, B, C, N, kind):
    """Feature clouds (B,C,N) that stress the tensor-core pre-filter's error bound (knn_tc.cu): exact ties (grid / dup), norms much
    larger than distances (offset), extreme magnitudes (scaled: the per-cloud FP16 scaling), post-activation features (relu),
    tight clusters far apart (near-ties inside, huge gaps between), features that live in a few directions (lowrank)."""
    x = rng.standard_normal((B, C, N)).astype(np.float32)
    if kind == "grid":
        return rng.integers(-2, 3, size=(B, C, N)).astype(np.float32)
    if kind == "dup":
        base = rng.standard_normal((B, C, max(1, N // 5 + 1))).astype(np.float32)
        return np.take_along_axis(base, rng.integers(0, base.shape[2], size=(B, 1, N)).repeat(C, axis=1), axis=2)
    if kind == "offset":
        return (x * np.float32(rng.choice([0.01, 0.3, 1.0])) + np.float32(rng.choice([5.0, 40.0, 300.0]))).astype(np.float32)
    if kind == "scaled":
        return (x * np.float32(10.0 ** rng.integers(-12, 13))).astype(np.float32)
    if kind == "relu":
        return np.where(x > 0, x, np.float32(0.2) * x).astype(np.float32) * np.float32(3.0)
    if kind == "clusters":
        centres = (rng.standard_normal((B, C, 6)) * 20).astype(np.float32)
        pick = rng.integers(0, 6, size=(B, 1, N)).repeat(C, axis=1)
        return (np.take_along_axis(centres, pick, axis=2) + x * np.float32(1e-3)).astype(np.float32)
    if kind == "lowrank":
        basis = rng.standard_normal((B, C, 3)).astype(np.float32)
        coef = rng.standard_normal((B, 3, N)).astype(np.float32)
        return np.einsum("bcr,brn->bcn", basis, coef).astype(np.float32)
    return x


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(B=st.integers(1, 3), C=st.sampled_from([16, 24, 32, 64, 64, 72, 128]), N=st.integers(256, 1300), k=st.integers(1, 32),
       kind=st.sampled_from(TC_KINDS), seed=st.integers(0, 2 ** 20))
def test_knn_tensor_core_path_random_inputs_match_oracle(B, C, N, k, kind, seed):
    """Shapes inside the tensor-core path's range, value distributions chosen against its error bound: bit-exact or bust."""
    from oracle import cpu_oracle as O
    import fpsg_b200
    x = np.ascontiguousarray(_features(np.random.default_rng(seed), B, C, N, kind))
    got = fpsg_b200.knn(torch.from_numpy(x).cuda(), k).cpu().numpy()
    ref = O.knn_f32(x, k)
    assert (got == ref).all(), f"B={B} C={C} N={N} k={k} kind={kind} seed={seed}: {(got != ref).sum()} positions differ"
