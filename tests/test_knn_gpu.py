"""GPU parity tests for knn(x, k): CUDA library vs the CPU oracle (bit-exact, including tie order) and vs the
golden outputs of the real reference function (exact up to permutations among exactly-equal FP32 values)."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, assert_knn_equivalent, make_cloud_cn

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import cpu_oracle
    return cpu_oracle


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "knn_*.npz"))), ids=os.path.basename)
def test_knn_matches_reference_golden(path):
    import fpsg_b200
    O = _oracle()
    z = np.load(path)
    x, k, ref = z["x"], int(z["k"]), z["idx"]
    idx = fpsg_b200.knn(torch.tensor(x).cuda(), k).cpu().numpy()
    _, pd = O.knn_f32(x, k, return_pd=True)
    assert_knn_equivalent(idx, ref, pd, os.path.basename(path))
    assert (idx == O.knn_f32(x, k)).all(), "bit-exact agreement with the oracle (tie order included)"


@pytest.mark.parametrize("B,C,N,k", [(2, 3, 100, 20), (3, 3, 1024, 20), (2, 3, 1000, 1), (1, 3, 33, 33), (2, 3, 257, 64),
                                     (1, 3, 300, 128), (2, 64, 300, 20), (2, 64, 1024, 20), (1, 128, 520, 20),
                                     (2, 5, 77, 7), (1, 3, 9000, 20), (1, 70, 100, 40), (1, 16, 150, 100), (1, 400, 64, 5),
                                     (1, 64, 2100, 20), (2, 1, 50, 3), (40, 8, 2048, 20)])
def test_knn_bit_exact_vs_oracle(B, C, N, k):
    import fpsg_b200
    O = _oracle()
    x = make_cloud_cn(B, C, N, seed=B + C + N)
    idx = fpsg_b200.knn(x.cuda(), k)
    assert idx.dtype == torch.int64 and idx.shape == (B, N, k)
    ref = O.knn_f32(x.numpy(), k)
    got = idx.cpu().numpy()
    assert (got == ref).all(), f"{(got != ref).sum()} of {ref.size} positions differ"


@pytest.mark.parametrize("B,C,N,k", [
    (1, 64, 1024, 100), (1, 128, 600, 128), (2, 32, 2100, 70),   # knn_kernel<4,false>, several candidate tiles (lists live in shared memory between tiles)
    (1, 64, 1024, 50), (1, 128, 600, 64),                         # knn_blocked_kernel<2>, several tiles
    (1, 3, 5000, 100), (1, 3, 3000, 40),                          # knn_kernel<4,true> / <2,true>
    (1, 64, 2500, 20),                                            # tensor-core path (knn_tc.cu), several 256-candidate blocks and tiles
    (1, 20, 2500, 20),                                            # generic C outside the tensor-core range, beyond the scratch-row kernel: knn_gemm_kernel
    (2, 20, 900, 20),                                             # ... inside it: knn_rows_kernel + knn_rows_select_kernel
    (1, 3, 9000, 32), (1, 450, 200, 20),                          # knn_bulk_kernel<true> / <false> (C too large for the register-tiled kernels)
])
def test_knn_every_dispatch_path_multi_tile(B, C, N, k):
    """One shape per kernel of knn.cu's dispatch, each large enough for several candidate tiles."""
    import fpsg_b200
    O = _oracle()
    x = make_cloud_cn(B, C, N, seed=7 * B + C + N + k)
    got = fpsg_b200.knn(x.cuda(), k).cpu().numpy()
    ref = O.knn_f32(x.numpy(), k)
    assert (got == ref).all(), f"{(got != ref).sum()} of {ref.size} positions differ"


@pytest.mark.parametrize("B,N,k", [(1, 8192, 20), (1, 4100, 32), (2, 2048, 20), (2, 1024, 32), (2, 1024, 1), (1, 897, 20), (1, 896, 20), (2, 512, 20),
                                   (2, 257, 7), (3, 64, 20), (2, 40, 32), (2, 33, 3)])
def test_knn_xyz_selection_kernel_shape_space(B, N, k):
    """knn_select_kernel (C = 3, k <= 32, N <= 8192) over its plan space: every fine-group size (8 .. 64 candidates), clouds with fewer
    fine groups than the 64 coarse ones (empty coarse groups), the two-octet specialisation (896 < N <= 1024) and its neighbours,
    k = 1 and k = 32 (the network's full width), clouds smaller than k groups (everything takes the streaming path)."""
    import fpsg_b200
    O = _oracle()
    x = make_cloud_cn(B, 3, N, seed=11 * B + N + k)
    got = fpsg_b200.knn(x.cuda(), k).cpu().numpy()
    ref = O.knn_f32(x.numpy(), k)
    assert (got == ref).all(), f"{(got != ref).sum()} of {ref.size} positions differ"


def test_knn_duplicates_and_ties():
    import fpsg_b200
    O = _oracle()
    g = torch.Generator().manual_seed(0)
    x = torch.randint(-2, 3, (2, 3, 200), generator=g).float()  # grid points: massive exact ties
    got = fpsg_b200.knn(x.cuda(), 20).cpu().numpy()
    assert (got == O.knn_f32(x.numpy(), 20)).all()


@pytest.mark.parametrize("C,N,k", [(3, 300, 20), (3, 1500, 32), (16, 300, 20), (64, 700, 7), (3, 2500, 20), (8, 2048, 32)])
def test_knn_degenerate_clouds_take_the_overflow_paths(C, N, k):
    """All points identical / a handful of distinct points: every candidate ties with the threshold, so the
    threshold-selection kernels must fall back to their streaming merge -- and still be bit-exact (lowest index
    first among equal values)."""
    import fpsg_b200
    O = _oracle()
    g = torch.Generator().manual_seed(N + C)
    same = torch.randn(1, C, 1, generator=g).expand(1, C, N).contiguous()
    few = torch.randn(1, C, 3, generator=g)[:, :, torch.randint(0, 3, (N,), generator=g)].contiguous()
    for x in (same, few):
        got = fpsg_b200.knn(x.cuda(), k).cpu().numpy()
        ref = O.knn_f32(x.numpy(), k)
        assert (got == ref).all(), f"{(got != ref).sum()} of {ref.size} positions differ"


def test_knn_vs_torch_reference_formula_on_gpu():
    """Same-device check against the reference expression (src/dgcnn/model.py:13-20) evaluated by torch on the
    GPU: neighbour values must agree to FP32 rounding of the expanded form."""
    import fpsg_b200
    from oracle.torch_formulation import knn_torch
    for C in (3, 64, 128):
        x = make_cloud_cn(4, C, 1024, seed=C).cuda()
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = knn_torch(x, 20)
        got = fpsg_b200.knn(x, 20)
        inner = -2 * torch.matmul(x.transpose(2, 1), x)
        xx = torch.sum(x ** 2, dim=1, keepdim=True)
        pd = -xx - inner - xx.transpose(2, 1)
        va = torch.gather(pd, 2, got)
        vb = torch.gather(pd, 2, ref)
        scale = float(xx.max())
        assert float((va - vb).abs().max()) <= 16 * np.finfo(np.float32).eps * scale
        assert float((got == ref).float().mean()) > 0.995


@pytest.mark.parametrize("C,N,k", [(3, 1024, 20), (3, 300, 5), (3, 2048, 32), (16, 300, 20)])
def test_knn_non_finite_inputs_terminate_with_indices_in_range(C, N, k):
    """NaN / Inf coordinates (include/fpsg_b200.h: the ranking of NaN scores is unspecified): no hang, every index in [0, N),
    and the clouds without a non-finite value are unaffected."""
    import fpsg_b200
    from oracle import cpu_oracle as O
    x = make_cloud_cn(4, C, N, seed=41)
    x[0, 1, 7] = float("nan")
    x[1, 0, 9] = float("inf")
    x[2, :, 11] = float("-inf")
    idx = fpsg_b200.knn(x.cuda(), k)
    torch.cuda.synchronize()
    assert int(idx.min()) >= 0 and int(idx.max()) < N
    assert np.array_equal(idx[3].cpu().numpy(), O.knn_f32(x[3:4].numpy(), k)[0])


def test_knn_error_behaviour():
    import fpsg_b200
    x = torch.zeros(2, 3, 16, device="cuda")
    with pytest.raises(RuntimeError):
        fpsg_b200.knn(x.cpu(), 4)
    with pytest.raises(RuntimeError):
        fpsg_b200.knn(x, 17)
    with pytest.raises(RuntimeError):
        fpsg_b200.knn(x, 0)
    with pytest.raises(RuntimeError):
        fpsg_b200.knn(x.double(), 4)
    assert fpsg_b200.knn(x[:0], 4).shape == (0, 16, 4)


def test_knn_batches_beyond_one_call_are_chunked(monkeypatch):
    """The library takes at most 65535 clouds per call; the wrapper chunks (checked with a small chunk size)."""
    import fpsg_b200
    import sys
    K = sys.modules["fpsg_b200.knn"]  # the module (the package attribute of that name is the function)
    O = _oracle()
    x = make_cloud_cn(7, 3, 60, seed=21)
    monkeypatch.setattr(K, "_MAX_CLOUDS_PER_CALL", 3)
    got = fpsg_b200.knn(x.cuda(), 5).cpu().numpy()
    assert (got == O.knn_f32(x.numpy(), 5)).all()
