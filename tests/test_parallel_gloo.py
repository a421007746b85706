"""CPU: the N>1 host logic (replica sharding + the metrics all_gather) with the gloo backend, world_size 2 and 3."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maro_b200.parallel import gather_metrics, owner_of, scatter_actions, shard_range


def test_shard_range_partitions():
    for n in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                cover += list(range(lo, hi))
                for i in range(lo, hi):
                    assert owner_of(i, n, world) == r
            assert cover == list(range(n))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    # each rank "simulates" its shard: metrics row = f(global replica id)
    ids = torch.arange(lo, hi, dtype=torch.int64)
    local = torch.stack([ids * 10, ids * 10 + 1, ids * 10 + 2], 1)
    allm = gather_metrics(local, n)
    acts = scatter_actions(np.arange(n * 4).reshape(n, 4), n, rank, world)
    q.put((rank, allm.numpy().tolist(), acts[:, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 10), (3, 7)])
def test_gather_metrics_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = [[i * 10, i * 10 + 1, i * 10 + 2] for i in range(n)]
    for rank, allm, acts in outs:
        assert allm == want
        lo, hi = shard_range(n, rank, world)
        assert acts == [i * 4 for i in range(lo, hi)]
