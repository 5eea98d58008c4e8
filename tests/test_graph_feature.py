"""EdgeConv graph feature: CPU check of the restatement against a direct numpy statement, and
GPU parity of the fused kernels (forward bit-exact, backward <= 1e-5 vs torch.autograd on the restatement)."""
import numpy as np
import pytest
import torch

from oracle import torch_formulation as TF
from tests.helpers import assert_close_rel, make_cloud_cn


def _numpy_feature(x, idx):
    B, C, N = x.shape
    pts = x.transpose(0, 2, 1)                                   # (B,N,C)
    nbr = np.stack([pts[b][idx[b]] for b in range(B)])           # (B,N,k,C)
    ctr = np.broadcast_to(pts[:, :, None, :], nbr.shape)
    return np.concatenate([nbr - ctr, ctr], axis=3).transpose(0, 3, 1, 2)


def test_restatement_matches_numpy_statement_cpu():
    x = make_cloud_cn(2, 5, 40, seed=1)
    idx = TF.knn_torch(x, 7)
    out = TF.get_graph_feature_torch(x, k=7, idx=idx)
    assert out.shape == (2, 10, 40, 7)
    assert np.array_equal(out.numpy(), _numpy_feature(x.numpy(), idx.numpy()))
    assert np.array_equal(TF.get_graph_feature_torch(x, k=7).numpy(), out.numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,N,k", [(2, 3, 100, 20), (3, 64, 257, 20), (1, 128, 96, 5), (2, 5, 33, 33), (1, 3, 1000, 1), (2, 7, 51, 3)])
def test_forward_bit_exact_and_backward_vs_autograd(B, C, N, k):
    import fpsg_b200
    x = make_cloud_cn(B, C, N, seed=B + C + N)
    idx = fpsg_b200.knn(x.cuda(), k)
    a = x.cuda().requires_grad_(True)
    out = fpsg_b200.get_graph_feature(a, k=k, idx=idx)
    assert out.shape == (B, 2 * C, N, k) and out.is_contiguous()
    assert np.array_equal(out.detach().cpu().numpy(), _numpy_feature(x.numpy(), idx.cpu().numpy()))
    g = torch.Generator().manual_seed(3)
    w = torch.randn(out.shape, generator=g)
    (out * w.cuda()).sum().backward()
    # the order-free statement of the same sum: autograd on the restatement in float64 (each gradient element adds
    # ~3k signed terms; FP32 autograd's own rounding is ~1e-5 of their magnitude, so it is only compared norm-wise)
    ref64 = x.double().requires_grad_(True)
    (TF.get_graph_feature_torch(ref64, k=k, idx=idx.cpu()) * w.double()).sum().backward()
    assert_close_rel(a.grad.cpu().numpy(), ref64.grad.numpy(), 1e-5, "grad_x vs float64 autograd")
    ref32 = x.clone().requires_grad_(True)
    (TF.get_graph_feature_torch(ref32, k=k, idx=idx.cpu()) * w).sum().backward()
    assert_close_rel(a.grad.cpu().numpy(), ref32.grad.numpy(), 1e-4, "grad_x vs float32 autograd", scale_frac=1.0)


@pytest.mark.gpu
def test_batches_beyond_one_call_are_chunked(monkeypatch):
    """The library takes at most 65535 clouds per call; the wrapper chunks and stays differentiable (checked with a small chunk size)."""
    import sys
    import fpsg_b200
    G = sys.modules["fpsg_b200.graph_feature"]
    x = make_cloud_cn(7, 6, 50, seed=13)
    w = torch.randn(7, 12, 50, 5, generator=torch.Generator().manual_seed(2)).cuda()

    def run():
        a = x.cuda().requires_grad_(True)
        out = fpsg_b200.get_graph_feature(a, k=5)
        (out * w).sum().backward()
        return out.detach().clone(), a.grad.clone()

    whole = run()
    monkeypatch.setattr(G, "_MAX_CLOUDS_PER_CALL", 3)
    parts = run()
    assert torch.equal(whole[0], parts[0]) and torch.allclose(whole[1], parts[1], rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_default_idx_path_and_dgcnn_patch():
    import sys
    import types
    import fpsg_b200
    import fpsg_b200.shims as shims
    x = make_cloud_cn(2, 3, 300, seed=9).cuda()
    out = fpsg_b200.get_graph_feature(x, k=20)
    ref = TF.get_graph_feature_torch(x.cpu(), k=20, idx=fpsg_b200.knn(x, 20).cpu())
    assert np.array_equal(out.cpu().numpy(), ref.numpy())
    mod = types.ModuleType("dgcnn.model")
    mod.knn = lambda x, k: None
    mod.get_graph_feature = lambda x, k=20, idx=None: None
    sys.modules["dgcnn.model"] = mod
    try:
        assert shims.patch_dgcnn_knn(graph_feature=True)
        assert mod.get_graph_feature is fpsg_b200.get_graph_feature and mod.knn is fpsg_b200.knn
    finally:
        sys.modules.pop("dgcnn.model", None)
    with pytest.raises(RuntimeError):
        fpsg_b200.get_graph_feature(x.cpu(), k=20)


@pytest.mark.gpu
def test_c3_layer_shapes():
    """The four EdgeConv layers of DGCNNfeat (src/dgcnn/model.py:62-74): C = 3, 64, 64, 128 at N = 1024, k = 20."""
    import fpsg_b200
    for C in (3, 64, 128):
        x = make_cloud_cn(4, C, 1024, seed=C).cuda().requires_grad_(True)
        out = fpsg_b200.get_graph_feature(x, k=20)
        assert out.shape == (4, 2 * C, 1024, 20)
        out.sum().backward()
        # d/dx of sum(out): the centre half contributes +k, the difference half contributes (in-degree - k)
        idx = fpsg_b200.knn(x.detach(), 20)
        indeg = torch.zeros(4, 1024, device="cuda").scatter_add_(1, idx.reshape(4, -1), torch.ones(4, 1024 * 20, device="cuda"))
        assert torch.allclose(x.grad, indeg[:, None, :].expand(-1, C, -1), rtol=0, atol=1e-4)


@pytest.mark.gpu
def test_out_of_range_idx_is_rejected():
    """A caller-supplied idx outside [0, N) raises (the reference's gather fails on it too) instead of reading out of bounds."""
    import fpsg_b200
    x = torch.randn(2, 8, 50, device="cuda")
    idx = torch.randint(0, 50, (2, 50, 5), device="cuda")
    idx[1, 3, 2] = 50
    with pytest.raises(IndexError):
        fpsg_b200.get_graph_feature(x, k=5, idx=idx)
    idx[1, 3, 2] = -1
    with pytest.raises(IndexError):
        fpsg_b200.get_graph_feature(x, k=5, idx=idx)
