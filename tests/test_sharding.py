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


def _worker_empty_rank(rank, world, port, q):
    """Rank 1 owns no block at all (world > n_blocks, or an LPT deal that leaves a shard empty)."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = np.arange(3) if rank == 0 else np.zeros(0, np.int64)
        rec = np.stack([mine + 10, mine * 2], axis=1) if rank == 0 else np.zeros((0, 0), np.int64)
        got = sharding.gather_block_records(mine, rec, 3)
        q.put((rank, bool(np.array_equal(got, np.stack([np.arange(3) + 10, np.arange(3) * 2], axis=1)))))
    finally:
        dist.destroy_process_group()


def test_gather_with_an_empty_rank_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker_empty_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


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


def test_c_lpt_partition_matches_the_python_restatement():
    """hp_lpt_partition / hp_block_costs (C, host only) against sharding.lpt_partition / block_costs."""
    from hiphase_b200 import lib
    rng = np.random.default_rng(1)
    n_var = np.exp(rng.uniform(np.log(20), np.log(2000), 3000)).astype(np.uint32)
    n_cells = (n_var.astype(np.uint64) * 30 + rng.integers(0, 50, 3000).astype(np.uint64))
    cost = lib.block_costs(n_var, n_cells)
    assert np.array_equal(cost.astype(np.int64), sharding.block_costs(n_var, n_cells))
    for world in (1, 2, 3, 8):
        shard_of = lib.lpt_partition(cost, world)
        parts = sharding.lpt_partition(cost.astype(np.int64), world)
        for r in range(world):
            assert np.array_equal(np.flatnonzero(shard_of == r), parts[r])
    # more shards than blocks: some shards stay empty, every block is dealt once
    shard_of = lib.lpt_partition(cost[:3], 8)
    assert len(set(shard_of.tolist())) == 3


def test_stream_generator_is_shard_independent():
    """The C++ HG002-scale stream: block b is the same whatever subset / order / thread count generates it."""
    from hiphase_b200 import synth
    a = synth.config_c3_stream(64, first_block=100, threads=1)
    b = synth.stream_blocks(np.arange(100, 164)[::-1].copy(), threads=4)
    nv, noisy = synth.stream_headers(100, 64)
    assert np.array_equal(nv, np.diff(a.var_off.astype(np.int64)))
    for i in (0, 5, 63):
        x, y = a.block(i), b.block(63 - i)
        assert x["n_var"] == y["n_var"] and len(x["reads"]) == len(y["reads"])
        assert all(r[0] == s[0] and np.array_equal(r[1], s[1]) and np.array_equal(r[2], s[2]) for r, s in zip(x["reads"], y["reads"]))
        assert np.array_equal(x["ignored"], y["ignored"])
    # the shape SURVEY 8d asks for: N in [20, 2000], ~30x coverage, every variant covered, reads with >= 2 set alleles
    assert nv.min() >= 20 and nv.max() <= 2000
    cover = np.zeros(a.n_vars + 1, np.int64)
    for blk in range(a.n_blocks):
        v0 = int(a.var_off[blk])
        for r in range(int(a.read_off[blk]), int(a.read_off[blk + 1])):
            cover[v0 + int(a.read_start[r])] += 1; cover[v0 + int(a.read_end[r])] -= 1
            c0, c1 = int(a.cell_off[r]), int(a.cell_off[r + 1])
            assert a.alleles[c0] < 2 and a.alleles[c1 - 1] < 2 and (a.alleles[c0:c1] < 2).sum() >= 2
    assert 25 < a.n_cells / a.n_vars < 35


def test_bench_workloads_deal_every_block_exactly_once():
    """bench.py's sharding of the named configurations (host logic of the N-GPU runs): every block of C5 goes to exactly one
    rank, a rank's chunks partition its shard, the shards' estimated costs are balanced, and both arms sample the same blocks."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for world in (1, 2, 8):
        wl = bench.Workload("c5", world)
        assert wl.n_total == 200000 and wl.scaling == "strong"
        seen = np.zeros(wl.n_total, np.int32)
        costs = []
        for rank in range(world):
            ids = wl.rank_ids(rank)
            seen[ids.astype(np.int64)] += 1
            costs.append(int(wl.cost[ids.astype(np.int64)].sum()))
            chunks = wl.chunk_ids(rank, 8)
            assert sum(len(c) for c in chunks) == len(ids)
            assert np.array_equal(np.sort(np.concatenate(chunks)), np.sort(ids))
            # chunks are miniatures of the shard: their costs agree within a few per cent
            cc = [int(wl.cost[c.astype(np.int64)].sum()) for c in chunks]
            assert max(cc) <= 1.05 * min(cc)
        assert (seen == 1).all()
        assert max(costs) <= 1.001 * min(costs)
    wl = bench.Workload("c3", 1)
    assert wl.n_total == 10000 and "configs[2]" in wl.label
    a, b = wl.sample_ids(3), wl.sample_ids(3)
    assert np.array_equal(a, b) and len(np.unique(a)) == bench.CPU_SAMPLE_BLOCKS and a.max() < wl.n_total
    assert not np.array_equal(wl.sample_ids(3), wl.sample_ids(4))
