"""world_size-2 gloo tests of the multi-GPU host logic: sharding + scalar all-reduce. The
per-shard compute is the CPU oracle here (injected), so the test runs without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fpsg_b200.sharded import ShardedChamferLoss, shard_range
from tests.helpers import make_pair


def test_shard_range_partitions_exactly():
    for B in (0, 1, 5, 32, 37, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _oracle_loss(p1, p2, w1=1.0, w2=1.0):
    from oracle import cpu_oracle as O
    return torch.from_numpy(O.chamfer_fwd(p1.numpy(), p2.numpy(), w1, w2)[4])


def _worker(rank, world, port, B, N, M, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p1, p2 = make_pair(B, N, M, seed=99)
        lo, hi = shard_range(B, rank, world)
        loss_local, total = ShardedChamferLoss(loss_fn=_oracle_loss)(p1[lo:hi], p2[lo:hi])
        out[rank] = (loss_local.numpy().copy(), total.item())
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("B", [5, 8])
def test_sharded_loss_equals_single_process(B):
    N, M, world = 70, 60, 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, N, M, out), nprocs=world, join=True)
    p1, p2 = make_pair(B, N, M, seed=99)
    full = _oracle_loss(p1, p2).numpy()
    got = np.concatenate([out[r][0] for r in range(world)])
    assert (got == full).all(), "per-pair losses are rank-local and must be identical to the unsharded run"
    for r in range(world):
        assert abs(out[r][1] - float(full.astype(np.float64).sum())) <= 1e-6 * abs(float(full.sum()))
