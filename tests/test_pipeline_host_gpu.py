"""GPU parity tests for the host-buffer streaming step (`fpsg_chamfer_pipeline_*`, fpsg_b200.ChamferPipeline): the
reference's training-loop shape (src/trainNetwork.py:141-148) with copies overlapped; every batch is checked against the
CPU oracle whatever the interleaving of submits and results."""
import numpy as np
import pytest
import torch

from tests.helpers import assert_close_rel, make_pair

pytestmark = pytest.mark.gpu


def _oracle_step(p1, p2, w1, w2):
    from oracle import cpu_oracle as O
    B, N, M = p1.shape[0], p1.shape[1], p2.shape[1]
    od1, od2, oi1, oi2, _ = O.chamfer_fwd(p1.numpy(), p2.numpy())
    loss = np.float64(w1) * od1.astype(np.float64).mean(1) + np.float64(w2) * od2.astype(np.float64).mean(1)
    g1 = np.full((B, N), np.float32(w1) / np.float32(N), np.float32)
    g2 = np.full((B, M), np.float32(w2) / np.float32(M), np.float32)
    gp1, _ = O.chamfer_bwd(p1.numpy(), p2.numpy(), oi1, oi2, g1, g2, True, False)
    return loss, gp1


@pytest.mark.parametrize("B,N,M,depth,pinned", [(3, 300, 257, 1, False), (4, 513, 600, 2, True), (5, 2048, 2048, 3, True), (2, 4500, 1000, 2, False)])
def test_pipeline_matches_oracle_under_interleaving(B, N, M, depth, pinned):
    import fpsg_b200
    w1, w2 = 0.75, 1.25
    pipe = fpsg_b200.ChamferPipeline(B, N, M, w1=w1, w2=w2, depth=depth)
    batches = [make_pair(B, N, M, seed=100 + s) for s in range(5)]
    if pinned:
        batches = [(a.pin_memory(), b.pin_memory()) for a, b in batches]
    want = [_oracle_step(a, b, w1, w2) for a, b in batches]
    got = []
    nxt = 0
    # keep the pipeline as full as it allows: submit until `depth` batches are in flight, take one result, submit again
    while len(got) < len(batches):
        while nxt < len(batches) and pipe.in_flight < depth:
            pipe.submit(*batches[nxt])
            nxt += 1
        s, loss, gp1 = pipe.result()
        got.append((s, loss.clone().numpy(), gp1.cpu().numpy()))
    for (s, loss, gp1), (oloss, ogp1) in zip(got, want):
        assert_close_rel(loss, oloss, 1e-6, "loss")
        assert abs(s - oloss.sum()) <= 1e-5 * abs(oloss.sum())
        assert_close_rel(gp1, ogp1, 1e-5, "grad_p1")
    pipe.close()


def test_pipeline_device_submits_match_oracle_and_mix_with_host_submits():
    """`submit_device`: inputs already on the GPU (the decoder's output in the reference), no copies, same lanes."""
    import fpsg_b200
    B, N, M, depth = 4, 700, 513, 3
    pipe = fpsg_b200.ChamferPipeline(B, N, M, depth=depth)
    batches = [make_pair(B, N, M, seed=300 + s) for s in range(7)]
    want = [_oracle_step(a, b, 1.0, 1.0) for a, b in batches]
    dev = [(a.cuda(), b.cuda()) for a, b in batches]
    got, nxt = [], 0
    while len(got) < len(batches):
        while nxt < len(batches) and pipe.in_flight < depth:
            if nxt % 3 == 2:
                pipe.submit(*batches[nxt])          # a host batch between device batches
            else:
                pipe.submit_device(*dev[nxt])
            nxt += 1
        s, loss, gp1 = pipe.result()
        got.append((s, loss.clone().numpy(), gp1.cpu().numpy()))
    for (s, loss, gp1), (oloss, ogp1) in zip(got, want):
        assert_close_rel(loss, oloss, 1e-6, "loss")
        assert abs(s - oloss.sum()) <= 1e-5 * abs(oloss.sum())
        assert_close_rel(gp1, ogp1, 1e-5, "grad_p1")
    with pytest.raises(RuntimeError):
        pipe.submit_device(batches[0][0], dev[0][1])  # a CPU tensor is not a device buffer
    pipe.close()


def test_pipeline_argument_errors():
    import fpsg_b200
    from fpsg_b200._lib import FpsgError
    pipe = fpsg_b200.ChamferPipeline(2, 64, 64, depth=1)
    a, b = make_pair(2, 64, 64)
    with pytest.raises(FpsgError):
        pipe.result()                      # nothing in flight
    pipe.submit(a, b)
    with pytest.raises(FpsgError):
        pipe.submit(a, b)                  # depth exceeded
    with pytest.raises(RuntimeError):
        pipe.submit(a.cuda(), b)           # device tensors are not host buffers
    with pytest.raises(RuntimeError):
        pipe.submit(a[:, :10], b)          # wrong shape
    pipe.result()
    with pytest.raises(FpsgError):
        fpsg_b200.ChamferPipeline(2, 64, 64, depth=9)
    pipe.close()
