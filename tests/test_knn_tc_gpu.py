"""GPU parity tests for the tensor-core kNN path (fpsg_b200/csrc/knn_tc.cu: tcgen05 pre-filter + exact FP32 re-scoring; generic C,
k <= 32, 256 <= N <= 4096): bit-exact against the CPU oracle (`knn`, reference src/dgcnn/model.py:13-20) on the shapes and the
input classes that stress the filter's error bound and its fall-backs."""
import numpy as np
import pytest
import torch

from tests.helpers import make_cloud_cn

pytestmark = pytest.mark.gpu


def _check(x, k):
    import fpsg_b200
    from oracle import cpu_oracle as O
    got = fpsg_b200.knn(x.cuda(), k).cpu().numpy()
    ref = O.knn_f32(x.numpy(), k)
    assert got.shape == ref.shape
    assert (got == ref).all(), f"{(got != ref).sum()} of {ref.size} positions differ"


@pytest.mark.parametrize("B,C,N,k", [
    (2, 64, 1024, 20), (1, 128, 1024, 20), (3, 64, 256, 20),      # DGCNN layers 2-4; the smallest eligible cloud
    (1, 16, 300, 5), (2, 24, 511, 1), (1, 72, 700, 20),            # ragged N, C not a multiple of 64, k = 1
    (1, 64, 1000, 32), (1, 128, 2049, 20), (1, 32, 4096, 32),      # k = 32; N just past a block; the largest eligible cloud
    (5, 64, 1100, 20), (37, 64, 512, 20),                          # more tiles than one CTA per SM takes in a round
])
def test_knn_tensor_core_path_bit_exact(B, C, N, k):
    _check(make_cloud_cn(B, C, N, seed=11 * B + C + N + k), k)


@pytest.mark.parametrize("scale", [1e-4, 1.0, 3e3, 1e8])
def test_knn_tensor_core_path_any_magnitude(scale):
    """The FP16 operands are scaled per cloud by a power of two: tiny and huge features rank exactly as well."""
    _check(make_cloud_cn(2, 64, 512, seed=5) * scale, 20)


def test_knn_tensor_core_path_outlier_and_offsets():
    x = make_cloud_cn(3, 64, 600, seed=9)
    x[0, :, 17] *= 1e4           # one point with a huge norm: it sets the scale, the others sink towards FP16's small values
    x[1] += 50.0                 # a common offset: norms much larger than distances (the band covers many candidates)
    x[2, :, 100:400] = x[2, :, 100:101]   # 300 copies of one point: massive exact ties, ordered by index
    _check(x, 20)


def test_knn_tensor_core_path_degenerate_clouds():
    z = torch.zeros(1, 32, 300)
    _check(z, 20)                                  # all-zero cloud: every candidate ties, indices 0..k-1
    c = torch.ones(1, 64, 260) * 0.37
    _check(c, 32)                                  # identical non-zero points
    g = make_cloud_cn(1, 16, 512, seed=3)
    g[:, 8:, :] = 0.0                              # half of the channels empty
    _check(g, 20)


def test_knn_tensor_core_path_non_finite_inputs_terminate():
    """NaN / Inf features: no hang, no out-of-range index (the re-scoring takes every candidate when the filter marks fewer than k)."""
    import fpsg_b200
    x = make_cloud_cn(2, 64, 512, seed=4)
    x[0, 3, 7] = float("nan")
    x[1, 5, 9] = float("inf")
    idx = fpsg_b200.knn(x.cuda(), 20)
    torch.cuda.synchronize()
    assert int(idx.min()) >= 0 and int(idx.max()) < 512


def test_knn_tensor_core_path_is_cuda_graph_capturable():
    """The tensor-core kNN call (memset node, four kernels, two TMA tensor maps passed by value) captured in a CUDA graph and
    replayed on new data in the same buffers gives the oracle's result: no allocation, no host synchronisation inside."""
    from fpsg_b200 import _lib
    from oracle import cpu_oracle as O
    L = _lib.lib()
    B, C, N, k = 2, 64, 700, 20
    x1, x2 = make_cloud_cn(B, C, N, seed=71), make_cloud_cn(B, C, N, seed=72) * 7.0
    dev = torch.device("cuda")
    x = x1.to(dev)
    idx = torch.empty(B, N, k, dtype=torch.int64, device=dev)
    nb = L.fpsg_knn_workspace_bytes(B, C, N, k)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)

    def enqueue(stream):
        _lib.check(L.fpsg_knn_fwd(x.data_ptr(), B, C, N, k, idx.data_ptr(), ws.data_ptr(), nb, stream))

    enqueue(torch.cuda.current_stream().cuda_stream)  # warm-up outside capture (function attributes, driver entry point)
    torch.cuda.synchronize()
    assert (idx.cpu().numpy() == O.knn_f32(x1.numpy(), k)).all()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        enqueue(torch.cuda.current_stream().cuda_stream)
    x.copy_(x2)
    idx.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert (idx.cpu().numpy() == O.knn_f32(x2.numpy(), k)).all(), "replayed graph differs from the oracle"
    graph.replay()  # and again: the workspace is reused without any re-initialisation by the host
    torch.cuda.synchronize()
    assert (idx.cpu().numpy() == O.knn_f32(x2.numpy(), k)).all()


@pytest.mark.parametrize("B,C,N,k", [(2, 64, 1024, 20), (1, 128, 300, 20), (3, 16, 511, 7), (1, 72, 1100, 32), (1, 24, 2049, 5)])
def test_knn_tensor_core_path_stays_inside_its_buffers(B, C, N, k):
    """Guard zones around the workspace and the output (compute-sanitizer is not available on this pool): the four kernels of the
    tensor-core path -- TMA stores into shared memory aside -- must not write a byte outside what the ABI hands them, and must not
    depend on what the workspace held before."""
    from fpsg_b200 import _lib
    from oracle import cpu_oracle as O
    L = _lib.lib()
    dev = torch.device("cuda")
    G = 4096
    x = make_cloud_cn(B, C, N, seed=3 * C + N).to(dev)
    nb = int(L.fpsg_knn_workspace_bytes(B, C, N, k))
    ws = torch.full((nb + 2 * G,), 0xA5, dtype=torch.uint8, device=dev)      # poisoned: stale workspace contents must not matter
    out = torch.full((B * N * k + 2 * G // 8,), -7, dtype=torch.int64, device=dev)
    ws_ptr, out_ptr = ws.data_ptr() + G, out.data_ptr() + G
    assert ws_ptr % 16 == 0
    for _ in range(2):
        _lib.check(L.fpsg_knn_fwd(x.data_ptr(), B, C, N, k, out_ptr, ws_ptr, nb, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert bool((ws[:G] == 0xA5).all()) and bool((ws[G + nb:] == 0xA5).all()), "write outside the workspace"
    assert bool((out[:G // 8] == -7).all()) and bool((out[G // 8 + B * N * k:] == -7).all()), "write outside the output"
    got = out[G // 8:G // 8 + B * N * k].view(B, N, k).cpu().numpy()
    assert (got == O.knn_f32(x.cpu().numpy(), k)).all()
