"""Host-side multi-GPU logic on CPU: partitioning and the result hand-off over torch.distributed (gloo, world 2)."""
import os
import socket

import numpy as np
import pytest

from hiphase_b200 import sharding


def test_lpt_partition_covers_and_balances():
    rng = np.random.default_rng(0)
    costs = np.exp(rng.uniform(np.log(20), np.log(2000), 5000)).astype(np.int64) * 30
    for world in (1, 2, 4, 8):
        parts = sharding.lpt_partition(costs, world)
        allidx = np.concatenate(parts)
        assert sorted(allidx.tolist()) == list(range(len(costs)))
        loads = np.array([costs[p].sum() for p in parts])
        assert loads.max() <= loads.mean() + costs.max()          # LPT bound


def test_contiguous_shard():
    for n in (0, 1, 7, 1000, 200001):
        for world in (1, 2, 3, 8):
            spans = [sharding.contiguous_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        costs = (np.arange(n_total) % 17 + 1) * 10
        mine = sharding.lpt_partition(costs, world)[rank]
        # "results": a 3-field record derived from the block id (stands in for h-offset / stats)
        rec = np.stack([mine * 3 + 1, mine * mine, costs[mine]], axis=1)
        got = sharding.gather_block_records(mine, rec, n_total)
        ids = np.arange(n_total)
        ok = np.array_equal(got, np.stack([ids * 3 + 1, ids * ids, costs], axis=1))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 101, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
