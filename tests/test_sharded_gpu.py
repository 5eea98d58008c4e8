"""GPU tests of the sharded loss step with the scalar all-reduce inside the loss kernel (SURVEY.md 8e,
`fpsg_chamfer_multi_loss_fwd_bwd_dist`, fpsg_b200.sharded.ShardedChamferStep). One process per GPU; with a single visible
GPU the world is one rank (the exchange then runs against its own table), with two or more it is two ranks over NCCL +
NVLink peer memory. Every rank checks its shard against the CPU oracle and the all-reduced sums against the oracle's sum
over the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import assert_close_rel, make_pair

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, N, M, lag, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from fpsg_b200.sharded import ShardedChamferStep, shard_range
        from oracle import cpu_oracle as O
        lo, hi = shard_range(B, rank, world)
        step = ShardedChamferStep(hi - lo, N, M, dev, lag=lag)
        res = []
        for it in range(37):  # more steps than the ring has slots (32)
            p1, p2 = make_pair(B, N, M, seed=500 + it)
            a, b = p1[lo:hi].contiguous().to(dev), p2[lo:hi].contiguous().to(dev)
            loss, local_sum, gp1, seq = step(a, b)
            torch.cuda.synchronize(dev)
            want_seq = seq - lag
            g = float(step.global_sum(want_seq)) if want_seq >= 1 else None
            od1, od2, oi1, oi2, oloss = O.chamfer_fwd(p1.numpy(), p2.numpy())
            g1 = np.full((B, N), np.float32(1.0) / np.float32(N), np.float32)
            g2 = np.full((B, M), np.float32(1.0) / np.float32(M), np.float32)
            ogp1, _ = O.chamfer_bwd(p1.numpy(), p2.numpy(), oi1, oi2, g1, g2, True, False)
            assert_close_rel(loss.cpu().numpy(), oloss[lo:hi], 1e-6, "per-pair loss of the shard")
            assert_close_rel(gp1.cpu().numpy(), ogp1[lo:hi], 1e-5, "grad_p1 of the shard")
            assert abs(float(local_sum) - float(oloss[lo:hi].astype(np.float64).sum())) <= 1e-5 * abs(float(oloss[lo:hi].sum()))
            res.append((want_seq, g, float(oloss.astype(np.float64).sum())))
        if lag >= 1:  # nothing follows the last `lag` steps: an explicit flush collects their sums
            last = step.flush()
            torch.cuda.synchronize(dev)
            for sq in range(last - lag + 1, last + 1):
                res.append((sq, float(step.global_sum(sq)), None))
        # the reduced value of step s is the oracle's sum over the WHOLE batch of step s (lag 1: delivered one step later)
        sums = {it + 1: r[2] for it, r in enumerate(res) if r[2] is not None}
        for want_seq, g, _ in res:
            if want_seq >= 1:
                assert abs(g - sums[want_seq]) <= 1e-5 * abs(sums[want_seq]), (want_seq, g, sums[want_seq])
        out[rank] = [r[1] for r in res]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lag", [0, 1, 3, 8])
def test_sharded_step_with_in_kernel_allreduce(lag):
    world = 2 if torch.cuda.device_count() >= 2 else 1
    B, N, M = 6, 300, 257
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, N, M, lag, out), nprocs=world, join=True)
    assert len(out) == world
    if world == 2:
        assert out[0] == out[1], "rank-ordered summation: the reduced values are bit-identical on every rank"


def _pipeline_worker(rank, world, port, B, N, M, depth, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import fpsg_b200
        from fpsg_b200.sharded import PeerScalarExchange, shard_range
        from oracle import cpu_oracle as O
        lo, hi = shard_range(B, rank, world)
        pipe = fpsg_b200.ChamferPipeline(hi - lo, N, M, device=dev, depth=depth)
        pipe.set_exchange(PeerScalarExchange(dev))
        nb = max(3 * depth + 2, 35)  # wraps the ring of 32 slots
        batches = [make_pair(B, N, M, seed=700 + s) for s in range(nb)]
        want_global = []
        for p1, p2 in batches:
            _, _, _, _, oloss = O.chamfer_fwd(p1.numpy(), p2.numpy())
            want_global.append((float(oloss[lo:hi].astype(np.float64).sum()), float(oloss.astype(np.float64).sum())))
        got_global = {}
        nxt = done = 0
        while done < nb:
            while nxt < nb and pipe.in_flight < depth:
                p1, p2 = batches[nxt]
                if nxt % 2:
                    pipe.submit(p1[lo:hi].contiguous(), p2[lo:hi].contiguous())
                else:
                    pipe.submit_device(p1[lo:hi].contiguous().to(dev), p2[lo:hi].contiguous().to(dev))
                nxt += 1
            s, _loss, _gp1 = pipe.result()
            done += 1
            assert abs(s - want_global[done - 1][0]) <= 1e-5 * abs(want_global[done - 1][0]), "local sum of the shard"
            if done > depth:  # batch `done` collected the all-reduced sum of batch `done - depth`
                got_global[done - depth] = pipe.previous_global_sum()
        for n, g in got_global.items():
            w = want_global[n - 1][1]
            assert abs(g - w) <= 1e-5 * abs(w), (n, g, w)
        out[rank] = [got_global[n] for n in sorted(got_global)]
        pipe.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("depth", [1, 3, 4, 8])
def test_pipeline_lanes_with_in_kernel_allreduce(depth):
    """The host-buffer pipeline with an exchange: every batch publishes its sum and collects that of batch n - depth (lag = the
    number of lanes); host and device submits mixed."""
    world = 2 if torch.cuda.device_count() >= 2 else 1
    B, N, M = 6, 300, 257
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_pipeline_worker, args=(world, _free_port(), B, N, M, depth, out), nprocs=world, join=True)
    assert len(out) == world
    if world == 2:
        assert out[0] == out[1], "rank-ordered summation: the reduced values are bit-identical on every rank"
