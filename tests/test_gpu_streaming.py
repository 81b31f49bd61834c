"""Streaming entry (hp_astar_submit / hp_astar_wait), pinned host memory, the slab pool shared by batches in flight and the
result hand-off, against the CPU oracle.  The reference keeps 40 x threads jobs in flight (src/main.rs:328, 344-355)."""
import numpy as np
import pytest

import oracle_lib as O
from hiphase_b200 import _abi as A
from hiphase_b200 import lib, synth

pytestmark = pytest.mark.gpu


def _check(out, ref):
    assert (out.status == 0).all(), np.bincount(out.status)
    assert np.array_equal(out.h1, ref.h1) and np.array_equal(out.h2, ref.h2)
    assert np.array_equal(out.stats, ref.stats)


def test_jobs_in_flight_match_the_oracle():
    ctx = lib.Context(device=0)
    ctx.set_lanes(3)
    batches = [synth.config_c3_stream(40, first_block=40 * i) for i in range(7)] + [synth.config_c2(n_blocks=24)]
    refs = [O.astar_solve(b, threads=8, want_heuristic=False, want_counters=False) for b in batches]
    flight, done = [], []
    for i, b in enumerate(batches):
        if len(flight) == 3:
            j, k = flight.pop(0)
            done.append((k, ctx.astar_wait(j)))
        flight.append((ctx.astar_submit(b), i))
    assert isinstance(ctx.astar_poll(flight[-1][0]), bool)
    for j, k in flight:
        done.append((k, ctx.astar_wait(j)))
    assert [k for k, _ in done] == list(range(len(batches)))
    for k, out in done:
        _check(out, refs[k])
    # a fourth job without waiting must be refused, not silently queued over a busy lane
    hs = [ctx.astar_submit(batches[0]) for _ in range(3)]
    with pytest.raises(lib.HiPhaseB200Error):
        ctx.astar_submit(batches[0])
    for h in hs:
        _check(ctx.astar_wait(h), refs[0])
    ctx.close()


def test_pinned_buffers_and_sync_call_agree():
    ctx = lib.Context(device=0)
    arena = lib.PinnedArena()
    b = synth.stream_blocks(np.arange(300, 364, dtype=np.uint64), alloc=arena.alloc)
    out_p = A.AstarOut.sized(b.n_vars, b.n_blocks, alloc=arena.empty)
    h = ctx.astar_submit(b, out=out_p)
    ctx.astar_wait(h)
    ref = O.astar_solve(b, threads=8, want_heuristic=False, want_counters=False)
    _check(out_p, ref)
    _check(ctx.astar_solve_batch(b), ref)
    ctx.close()
    arena.close()


def test_device_calls_on_two_streams_overlap_safely():
    """hp_astar_solve_device on different streams takes different lanes; the slab pool is shared on the device."""
    import ctypes as C
    import torch
    ctx = lib.Context(device=0)
    dev = torch.device("cuda", 0)
    b = synth.config_c3_stream(200, first_block=1000)
    ref = O.astar_solve(b, threads=8, want_heuristic=False, want_counters=False)
    max_n = int(np.diff(b.var_off.astype(np.int64)).max())

    def to_dev(a):
        v = a.view(np.int64) if a.dtype == np.uint64 else (a.view(np.int32) if a.dtype == np.uint32 else a)
        return torch.from_numpy(v).to(dev)
    dten = {k: to_dev(getattr(b, k)) for k in A.BlockBatch.FIELDS}
    cp = lambda t, ty: C.cast(t.data_ptr(), ty)
    dbatch = A.hp_block_batch(b.n_blocks, cp(dten["var_off"], A.u64p), cp(dten["read_off"], A.u64p), cp(dten["read_start"], A.u32p),
                              cp(dten["read_end"], A.u32p), cp(dten["cell_off"], A.u64p), cp(dten["alleles"], A.u8p),
                              cp(dten["quals"], A.u8p), cp(dten["ignored"], A.u8p), cp(dten["is_snv"], A.u8p))
    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
    outs = []
    for s in streams:
        o = dict(h1=torch.zeros(b.n_vars, dtype=torch.uint8, device=dev), h2=torch.zeros(b.n_vars, dtype=torch.uint8, device=dev),
                 stats=torch.zeros(b.n_blocks * 7, dtype=torch.int64, device=dev), status=torch.full((b.n_blocks,), -1, dtype=torch.int32, device=dev))
        outs.append(o)
    torch.cuda.synchronize()
    for rep in range(2):
        for s, o in zip(streams, outs):
            do = A.hp_astar_out(cp(o["h1"], A.u8p), cp(o["h2"], A.u8p), C.cast(o["stats"].data_ptr(), C.POINTER(A.hp_phase_stats)),
                                cp(o["status"], A.i32p), A.u64p(), C.POINTER(A.hp_astar_counters)())
            ctx.astar_solve_device(dbatch, b.n_vars, b.n_reads, b.n_cells, max_n, do, s.cuda_stream)
    torch.cuda.synchronize()
    for o in outs:
        assert int((o["status"] != 0).sum()) == 0
        assert np.array_equal(o["h1"].cpu().numpy(), ref.h1) and np.array_equal(o["h2"].cpu().numpy(), ref.h2)
        assert np.array_equal(o["stats"].cpu().numpy().view(np.uint64).reshape(-1, 7), ref.stats.view(np.uint64).reshape(-1, 7))
    ctx.close()


def test_gather_results_single_rank_reorders_by_block_index():
    ctx = lib.Context(device=0)
    ctx.comm_init(None, 0, 1)
    ids = np.array([5, 2, 9, 0, 7, 1, 3, 8, 6, 4], np.uint64)           # this rank solved the blocks in some dealt order
    b = synth.stream_blocks(ids)
    out = ctx.astar_solve_batch(b)
    nv, _ = synth.stream_headers(0, 10)
    all_var_off = np.concatenate([[0], np.cumsum(nv)]).astype(np.uint64)
    allo = ctx.comm_gather_results(ids, b, out, all_var_off)
    ordered = synth.config_c3_stream(10)
    ref = O.astar_solve(ordered, threads=8, want_heuristic=False, want_counters=False)
    _check(allo, ref)
    ctx.close()


def test_block_service_packs_concurrent_one_block_calls():
    """hp_service_solve_one from many threads at once (the reference calls astar_solver from --threads pool workers, one block
    each: src/main.rs:385-408): every caller gets its own block's result, and the calls were packed into fewer launches."""
    import threading
    batch = synth.config_c3_stream(160, first_block=2000)
    ref = O.astar_solve(batch, threads=8, want_heuristic=False, want_counters=False)
    svc = lib.BlockService(device=0, linger_us=500)
    results = [None] * batch.n_blocks
    errors = []

    def worker(t, n_threads):
        try:
            for b in range(t, batch.n_blocks, n_threads):
                results[b] = svc.solve_one(batch, b)
        except Exception as e:           # noqa: BLE001
            errors.append(e)
    threads = [threading.Thread(target=worker, args=(t, 32)) for t in range(32)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for b in range(batch.n_blocks):
        v0, v1 = int(batch.var_off[b]), int(batch.var_off[b + 1])
        h1, h2, st = results[b]
        assert np.array_equal(h1, ref.h1[v0:v1]) and np.array_equal(h2, ref.h2[v0:v1])
        assert tuple(getattr(st, n) for n, _ in A.hp_phase_stats._fields_) == tuple(int(x) for x in ref.stats[b])
    n_batches, n_blocks = svc.counters()
    assert n_blocks == batch.n_blocks and n_batches < n_blocks
    # a block the reference would panic on comes back as an error to its caller only
    bad = A.BlockBatch.from_blocks([{"n_var": 4, "reads": [(0, [0, 1, 0, 1], [5, 5, 5, 5])], "ignored": [0, 1, 0, 0]}])
    with pytest.raises(lib.HiPhaseB200Error):
        svc.solve_one(bad, 0)
    h1, h2, st = svc.solve_one(batch, 3)
    assert np.array_equal(h1, ref.h1[int(batch.var_off[3]):int(batch.var_off[4])])
    svc.close()


def test_full_c3_one_batch_equals_chunks_in_flight():
    """BASELINE.json configs[2] at full size (the 10 000 blocks bench.py times): solved as one batch and as four interleaved
    chunks in flight on the lanes, the results are the same bytes block for block (a block's result does not depend on what it
    was batched with, nor on the team size / device share its launch got), and the size-independent invariants hold."""
    ctx = lib.Context(device=0)
    ctx.set_lanes(4)
    ids = np.arange(10000, dtype=np.uint64)
    whole = synth.stream_blocks(ids)
    out = ctx.astar_solve_batch(whole)
    assert (out.status == 0).all(), np.bincount(out.status)
    st = out.stats
    n = np.diff(whole.var_off.astype(np.int64))
    assert (st["actual_cost"] >= st["estimated_cost"]).all()                                   # phase_stats.rs:163
    assert ((st["phased_variants"] + st["homozygous_variants"] + st["skipped_variants"]) == n).all()
    ign = whole.ignored.astype(bool)
    assert (out.h1[ign] == 2).all() and (out.h2[ign] == 2).all() and (out.h1[~ign] < 2).all() and (out.h2[~ign] < 2).all()
    assert (st["pruned_solutions"] > 0).sum() >= 50                                            # the noisy 2 % prune
    parts = [ids[k::4] for k in range(4)]
    batches = [synth.stream_blocks(p) for p in parts]
    handles = [ctx.astar_submit(b) for b in batches]
    for p, b, h in zip(parts, batches, handles):
        o = ctx.astar_wait(h)
        assert (o.status == 0).all()
        assert np.array_equal(o.stats, st[p.astype(np.int64)])
        for k in range(0, len(p), 97):                                                         # haplotype bytes of every 97th block
            v0, v1 = int(whole.var_off[int(p[k])]), int(whole.var_off[int(p[k]) + 1])
            c0 = int(b.var_off[k])
            assert np.array_equal(o.h1[c0:c0 + v1 - v0], out.h1[v0:v1]) and np.array_equal(o.h2[c0:c0 + v1 - v0], out.h2[v0:v1])
    again = ctx.astar_solve_batch(whole)                                                       # idempotence
    assert np.array_equal(again.h1, out.h1) and np.array_equal(again.h2, out.h2) and np.array_equal(again.stats, st)
    ctx.close()
