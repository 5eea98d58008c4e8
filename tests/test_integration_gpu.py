"""GPU tests of the drop-in surfaces as the reference uses them: the kaolin shim inside a few_shot.py-style loss,
the dgcnn.model.knn patch feeding a get_graph_feature-style gather, CUDA-graph capture of the C-ABI calls,
batch sharding, and the BASELINE.json c2 shape (1184 pairs) spot-checked against the oracle."""
import ctypes
import sys
import types

import numpy as np
import pytest
import torch

from tests.helpers import assert_close_rel, make_cloud_cn, make_pair

pytestmark = pytest.mark.gpu


def test_fewshot_style_loss_through_kaolin_shim():
    """Mirrors reference src/models/few_shot.py:109-122: pc_metric(syn, ref).sum() for the query set and the
    support set, weighted, then backward -- with `from kaolin.metrics.pointcloud import chamfer_distance`."""
    import fpsg_b200.shims as shims
    from oracle import cpu_oracle as O
    shims.install(force=True, patch_dgcnn=False)
    from kaolin.metrics.pointcloud import chamfer_distance as pc_metric

    n_query, n_support, npts = 5, 32, 2048
    syn_q, ref_q = make_pair(n_query, npts, npts, seed=31)
    syn_s, ref_s = make_pair(n_support, npts, npts, seed=32)
    a_q = syn_q.cuda().requires_grad_(True)
    a_s = syn_s.cuda().requires_grad_(True)
    query_factor, support_factor = 1.0, 0.5
    loss_rec_q = pc_metric(a_q, ref_q.cuda()).sum()
    loss_rec_s = pc_metric(a_s, ref_s.cuda()).sum()
    ttl = query_factor * loss_rec_q + support_factor * loss_rec_s
    ttl.backward()

    oq = O.chamfer_fwd(syn_q.numpy(), ref_q.numpy())
    os_ = O.chamfer_fwd(syn_s.numpy(), ref_s.numpy())
    expect = query_factor * float(oq[4].astype(np.float64).sum()) + support_factor * float(os_[4].astype(np.float64).sum())
    assert abs(ttl.item() - expect) <= 1e-5 * abs(expect)
    g = np.float32(1.0) / np.float32(npts)
    gq, _ = O.chamfer_bwd(syn_q.numpy(), ref_q.numpy(), oq[2], oq[3], np.full((n_query, npts), g, np.float32),
                          np.full((n_query, npts), g, np.float32), True, False)
    gs, _ = O.chamfer_bwd(syn_s.numpy(), ref_s.numpy(), os_[2], os_[3], np.full((n_support, npts), g, np.float32),
                          np.full((n_support, npts), g, np.float32), True, False)
    assert_close_rel(a_q.grad.cpu().numpy(), query_factor * gq, 1e-5, "grad query")
    assert_close_rel(a_s.grad.cpu().numpy(), support_factor * gs, 1e-5, "grad support")


def test_dgcnn_knn_patch_and_edge_feature_contract():
    """`get_graph_feature` (reference src/dgcnn/model.py:23-42) resolves the module-level `knn` at call time and
    does `idx + idx_base` then a flat gather: patching the module global must be enough, and idx must be int64."""
    import fpsg_b200
    import fpsg_b200.shims as shims
    from oracle import cpu_oracle as O

    mod = types.ModuleType("dgcnn.model")

    def knn(x, k):  # stand-in for the reference function that the patch replaces
        raise AssertionError("the un-patched knn must not be called")

    def get_graph_feature(x, k=20):  # a consumer with the reference's index arithmetic (written for this test)
        B, C, N = x.shape
        idx = mod.knn(x, k=k)
        base = torch.arange(0, B, device=x.device).view(-1, 1, 1) * N
        flat = (idx + base).view(-1)
        pts = x.transpose(2, 1).contiguous().view(B * N, C)
        nbr = pts[flat].view(B, N, k, C)
        ctr = pts.view(B, N, 1, C).expand(-1, -1, k, -1)
        return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2).contiguous()

    mod.knn, mod.get_graph_feature = knn, get_graph_feature
    sys.modules["dgcnn.model"] = mod
    try:
        assert shims.patch_dgcnn_knn()
        assert mod.knn is fpsg_b200.knn
        x = make_cloud_cn(3, 3, 500, seed=12)
        feat = mod.get_graph_feature(x.cuda(), k=20)
        assert feat.shape == (3, 6, 500, 20)
        idx = O.knn_f32(x.numpy(), 20)
        pts = x.numpy().transpose(0, 2, 1)
        nbr = np.stack([pts[b][idx[b]] for b in range(3)])  # (B,N,k,C)
        expect = np.concatenate([nbr - pts[:, :, None, :], np.broadcast_to(pts[:, :, None, :], nbr.shape)], axis=3).transpose(0, 3, 1, 2)
        assert np.array_equal(feat.cpu().numpy(), expect)
    finally:
        sys.modules.pop("dgcnn.model", None)


def test_c_abi_calls_are_cuda_graph_capturable():
    """The ABI promises no allocation and no host synchronisation: forward + backward captured in one CUDA graph and
    replayed on new data must give the same result as eager calls."""
    from fpsg_b200 import _lib
    L = _lib.lib()
    B, N, M = 4, 1000, 900
    p1, p2 = make_pair(B, N, M, seed=41)
    q1, q2 = make_pair(B, N, M, seed=42)
    dev = torch.device("cuda")
    a, b = p1.to(dev), p2.to(dev)
    idx1 = torch.empty(B, N, dtype=torch.int64, device=dev)
    idx2 = torch.empty(B, M, dtype=torch.int64, device=dev)
    loss = torch.empty(B, device=dev)
    gl = torch.ones(B, device=dev)
    gp1 = torch.empty(B, N, 3, device=dev)
    nb = L.fpsg_chamfer_fwd_workspace_bytes(B, N, M)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)

    def enqueue(stream):
        _lib.check(L.fpsg_chamfer_loss_fwd(a.data_ptr(), b.data_ptr(), B, N, M, 1.0, 1.0, loss.data_ptr(), None, None,
                                           idx1.data_ptr(), idx2.data_ptr(), ws.data_ptr(), nb, stream))
        _lib.check(L.fpsg_chamfer_loss_bwd(a.data_ptr(), b.data_ptr(), idx1.data_ptr(), idx2.data_ptr(), gl.data_ptr(), 1.0, 1.0,
                                           B, N, M, gp1.data_ptr(), None, None, 0, stream))

    enqueue(torch.cuda.current_stream().cuda_stream)  # warm-up outside capture
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        enqueue(torch.cuda.current_stream().cuda_stream)
    a.copy_(q1)
    b.copy_(q2)
    graph.replay()
    torch.cuda.synchronize()
    got = (loss.clone(), idx1.clone(), idx2.clone(), gp1.clone())
    enqueue(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(got[0], loss) and torch.equal(got[1], idx1) and torch.equal(got[2], idx2) and torch.equal(got[3], gp1)
    from oracle import cpu_oracle as O
    o = O.chamfer_fwd(q1.numpy(), q2.numpy())
    assert (idx1.cpu().numpy() == o[2]).all() and (idx2.cpu().numpy() == o[3]).all()


def test_batch_sharding_is_exact():
    """a shard's results equal the same pairs' results in the unsharded call, bit for bit."""
    import fpsg_b200
    from fpsg_b200.sharded import shard_range
    p1, p2 = make_pair(13, 700, 650, seed=51)
    a = p1.cuda().requires_grad_(True)
    b = p2.cuda()
    full = fpsg_b200.chamfer_loss(a, b)
    full.sum().backward()
    for world in (2, 4):
        parts, grads = [], []
        for r in range(world):
            lo, hi = shard_range(13, r, world)
            s = p1[lo:hi].cuda().requires_grad_(True)
            l = fpsg_b200.chamfer_loss(s, b[lo:hi])
            l.sum().backward()
            parts.append(l.detach())
            grads.append(s.grad)
        assert torch.equal(torch.cat(parts), full.detach())
        assert torch.equal(torch.cat(grads), a.grad)


def test_c2_training_step_shape_spot_checked():
    """BASELINE.json configs[1]: 32 episodes x (32 + 5) = 1184 pairs, N = M = 2048, one call; 48 pairs vs the oracle."""
    import fpsg_b200
    from oracle import cpu_oracle as O
    p1, p2 = make_pair(1184, 2048, 2048, seed=1234 + 2)
    a = p1.cuda().requires_grad_(True)
    d1, d2, i1, i2 = fpsg_b200.chamfer_distance(a, p2.cuda())
    pick = np.linspace(0, 1183, 48).astype(np.int64)
    o = O.chamfer_fwd(p1.numpy()[pick], p2.numpy()[pick])
    assert (i1.cpu().numpy()[pick] == o[2]).all() and (i2.cpu().numpy()[pick] == o[3]).all()
    assert (d1.detach().cpu().numpy()[pick] == o[0]).all() and (d2.detach().cpu().numpy()[pick] == o[1]).all()
    (d1.mean(-1) + d2.mean(-1)).sum().backward()
    assert torch.isfinite(a.grad).all()


def test_c3_knn_shape_and_feature_counts():
    """BASELINE.json configs[2] (k=20, B=32, N=1024) for the DGCNN layer widths C in {3, 64, 128}."""
    import fpsg_b200
    from oracle import cpu_oracle as O
    for C in (3, 64, 128):
        x = make_cloud_cn(32 if C == 3 else 4, C, 1024, seed=100 + C)
        got = fpsg_b200.knn(x.cuda(), 20).cpu().numpy()
        assert (got == O.knn_f32(x.numpy(), 20)).all()


def test_concurrent_streams_and_threads_do_not_interfere():
    """SURVEY.md 8b (threading): the entry points take the stream explicitly, hold no global state and own no memory, so
    calls issued from several host threads onto several CUDA streams at once must give the single-stream results."""
    import threading
    import fpsg_b200
    from tests.helpers import make_cloud_cn, make_pair
    dev = torch.device("cuda", 0)
    jobs = []
    for s in range(4):
        p1, p2 = make_pair(6, 700 + 13 * s, 650 - 7 * s, seed=100 + s)
        x = make_cloud_cn(3, 3 if s % 2 == 0 else 16, 500 + s, seed=s)
        jobs.append((p1.to(dev), p2.to(dev), x.to(dev)))
    want = []
    for p1, p2, x in jobs:  # reference results, one after the other on the default stream
        a = p1.clone().requires_grad_(True)
        loss = fpsg_b200.chamfer_loss(a, p2)
        loss.sum().backward()
        want.append((loss.detach().clone(), a.grad.clone(), fpsg_b200.knn(x, 20)))
    torch.cuda.synchronize()
    got = [None] * len(jobs)
    errors = []

    def work(k):
        try:
            st = torch.cuda.Stream(device=dev)
            st.wait_stream(torch.cuda.default_stream(dev))
            with torch.cuda.stream(st):
                for _ in range(5):  # keep several calls in flight per stream
                    p1, p2, x = jobs[k]
                    a = p1.clone().requires_grad_(True)
                    loss = fpsg_b200.chamfer_loss(a, p2)
                    loss.sum().backward()
                    got[k] = (loss.detach().clone(), a.grad.clone(), fpsg_b200.knn(x, 20))
            st.synchronize()
        except Exception as e:  # pragma: no cover - reported below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (l0, g0, i0), (l1, g1, i1) in zip(want, got):
        assert torch.equal(l0, l1) and torch.equal(g0, g1) and torch.equal(i0, i1)
