"""The realignment pipeline entry (hp_realign_block_batch): graph-WFA -> MaxEditDistance -> local realignment, the
order-dependent switch-off of global realignment for the rest of a block, collapse and filter -- against the oracle's
mapping-by-mapping restatement of load_full_read_segments (src/read_parsing.rs:545-629)."""
import numpy as np
import pytest

import oracle_lib as O
from hiphase_b200 import _abi as A
from hiphase_b200 import lib, synth


def _batch(n_blocks, first_block, fail_min, ratio, **kw):
    d, vtypes = synth.config_realign(n_blocks, first_block, **kw)
    return A.RealignBatch(global_failure_minimum=fail_min, global_failure_ratio=ratio, **d), vtypes


def _same(out, ref):
    assert np.array_equal(out.map_mode, ref.map_mode), (np.bincount(out.map_mode, minlength=4), np.bincount(ref.map_mode, minlength=4))
    assert np.array_equal(out.block_disabled_at, ref.block_disabled_at)
    assert np.array_equal(out.block_failures, ref.block_failures) and np.array_equal(out.block_parsed, ref.block_parsed)
    a, b = out.as_struct().assembled, ref.as_struct().assembled
    assert int(a.n_reads) == int(b.n_reads) and int(a.n_cells) == int(b.n_cells)
    nr, nc = int(b.n_reads), int(b.n_cells)
    assert np.array_equal(out.read_off, ref.read_off)
    assert np.array_equal(out.read_start[:nr], ref.read_start[:nr]) and np.array_equal(out.read_end[:nr], ref.read_end[:nr])
    assert np.array_equal(out.cell_off[:nr + 1], ref.cell_off[:nr + 1])
    assert np.array_equal(out.alleles[:nc], ref.alleles[:nc]) and np.array_equal(out.quals[:nc], ref.quals[:nc])
    assert np.array_equal(out.group_class, ref.group_class)


SMALL = dict(window=9000, n_var=18, n_hom=8, n_reads=40, read_lo=1500, read_hi=4000, sv_max=300)


def test_oracle_replays_the_switch_off_rule():
    """Known-answer scenario for the rule itself (read_parsing.rs:593-600): with max_edit_distance 8 roughly half of the reads
    fail graph-WFA; with global-failure-count 5 / ratio 0.3 the block trips and every later mapping is local."""
    params = A.hp_params(1000, 3, 500, 8)
    batch, _ = _batch(2, 0, 5, 0.3, err=0.004, **SMALL)
    ref = O.realign_block_batch(batch, params)
    assert ref.rc == 0
    for b in range(batch.n_blocks):
        m0, m1 = int(batch.map_off[b]), int(batch.map_off[b + 1])
        mode = ref.map_mode[m0:m1]
        d = int(ref.block_disabled_at[b])
        assert d != 0xffffffff, "the scenario is meant to trip the rule"
        assert mode[d] == A.HP_MAP_LOCAL_FAILED                                  # the mapping that trips it is itself a failure
        assert not (mode[:d + 1] == A.HP_MAP_LOCAL_DISABLED).any()
        assert np.isin(mode[d + 1:], (A.HP_MAP_LOCAL_DISABLED, A.HP_MAP_SKIPPED)).all()
        # replay of the counters from the modes alone
        fails = parsed = 0
        for k, md in enumerate(mode[:d + 1]):
            if md == A.HP_MAP_SKIPPED:
                continue
            parsed += 1; fails += md == A.HP_MAP_LOCAL_FAILED
            tripped = fails >= 5 and fails / parsed >= 0.3
            assert tripped == (k == d)
    # with the reference defaults (50 / 0.5) this block never trips
    batch2, _ = _batch(1, 0, 50, 0.5, err=0.004, **SMALL)
    ref2 = O.realign_block_batch(batch2, params)
    assert ref2.rc == 0 and int(ref2.block_disabled_at[0]) == 0xffffffff and (ref2.map_mode == A.HP_MAP_LOCAL_FAILED).any()


@pytest.mark.gpu
@pytest.mark.parametrize("fail_min,ratio,max_ed", [(5, 0.3, 8), (50, 0.5, 8), (3, 0.9, 6), (1, 0.0, 500)])
def test_gpu_pipeline_matches_the_oracle(fail_min, ratio, max_ed):
    params = A.hp_params(1000, 3, 500, max_ed)
    ctx = lib.Context(params, device=0)
    batch, vtypes = _batch(3, 10, fail_min, ratio, err=0.004, **SMALL)
    ref = O.realign_block_batch(batch, params)
    assert ref.rc == 0
    out = ctx.realign_block_batch(batch)
    _same(out, ref)
    # ... and the assembled reads phase identically
    is_snv = np.concatenate([(np.array(v) == 0).astype(np.uint8) for v in vtypes])
    got = ctx.astar_solve_batch(out.block_batch(is_snv=is_snv))
    exp = O.astar_solve(ref.block_batch(is_snv=is_snv), params, want_heuristic=False, want_counters=False)
    assert np.array_equal(got.h1, exp.h1) and np.array_equal(got.h2, exp.h2) and np.array_equal(got.stats, exp.stats)
    ctx.close()


@pytest.mark.gpu
def test_gpu_pipeline_trips_at_fifty_failures():
    """The reference defaults (--global-failure-count 50, --max-global-failure-ratio 0.5) on a block with enough failing reads."""
    params = A.hp_params(1000, 3, 500, 6)
    ctx = lib.Context(params, device=0)
    batch, _ = _batch(1, 20, 50, 0.5, err=0.006, window=9000, n_var=18, n_reads=160, read_lo=1500, read_hi=4000, sv_max=300)
    ref = O.realign_block_batch(batch, params)
    assert ref.rc == 0 and int(ref.block_disabled_at[0]) != 0xffffffff and int(ref.block_failures[0]) >= 50
    _same(ctx.realign_block_batch(batch), ref)
    ctx.close()


@pytest.mark.gpu
def test_contexts_on_several_threads_agree_with_the_oracle():
    """The reference calls the loop from its pool workers (main.rs:385-408); here a worker owns a context.  Four threads run
    different batches side by side, several times over, with A* behind each: every result equals the oracle's."""
    import threading
    params = A.hp_params(1000, 3, 500, 8)
    work = [_batch(2, 30 + 5 * k, 5, 0.3, err=0.004, **SMALL) for k in range(4)]
    refs = [O.realign_block_batch(b, params) for b, _ in work]
    assert all(r.rc == 0 for r in refs)
    errors = []

    def worker(k):
        try:
            ctx = lib.Context(params, device=0)
            b, vtypes = work[k]
            is_snv = np.concatenate([(np.array(v) == 0).astype(np.uint8) for v in vtypes])
            exp = O.astar_solve(refs[k].block_batch(is_snv=is_snv), params, want_heuristic=False, want_counters=False)
            for _ in range(3):
                out = ctx.realign_block_batch(b)
                _same(out, refs[k])
                got = ctx.astar_solve_batch(out.block_batch(is_snv=is_snv))
                assert np.array_equal(got.h1, exp.h1) and np.array_equal(got.h2, exp.h2) and np.array_equal(got.stats, exp.stats)
            ctx.close()
        except BaseException as e:      # noqa: BLE001 -- reported by the main thread
            errors.append((k, repr(e)))
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


# ---- CIGAR projection (the pre-WFA half of global_realignment, read_parsing.rs:672-742) --------------------------------
def _plan_inputs(n_blocks, seed0, **kw):
    map_block, seg_off, sr, sd, sl, het_first, het_pos, hom_first, hom_pos, reads = [], [0], [], [], [], [0], [], [0], [], []
    for b in range(n_blocks):
        blk = synth.gen_local_block(np.random.default_rng(seed0 + b), p_ignored=0.0, **kw)
        het_pos += [v.position() for v in blk["variants"]]; het_first.append(len(het_pos))
        hom_pos += [v.position() for v in blk["homs"]]; hom_first.append(len(hom_pos))
        for (lo, hi, rp, segs, seq, q) in blk["jobs"]:
            map_block.append(b)
            for (a, r, n) in segs:
                sr.append(a); sd.append(r); sl.append(n)
            seg_off.append(len(sr)); reads.append((rp, segs, seq, q))
    return A.PlanBatch(map_block, seg_off, sr, sd, sl, het_first, het_pos, hom_first, hom_pos), reads


def test_plan_oracle_matches_the_python_mirror():
    from hiphase_b200.read_parsing import AlignedRead, plan_global_realignment
    pb, reads = _plan_inputs(3, 900, window=9000, n_var=12, n_hom=9, n_reads=30, read_lo=300, read_hi=3000, sv_max=300)
    ref = O.wfa_plan_batch(pb)
    assert ref.rc == 0
    skipped = 0
    for j, (rp, segs, seq, q) in enumerate(reads):
        b = int(pb.map_block[j])
        h0, m0 = int(pb.het_first[b]), int(pb.hom_first[b])
        plan = plan_global_realignment(AlignedRead(rp, segs, seq.tobytes(), q), pb.het_pos[h0:int(pb.het_first[b + 1])].tolist(),
                                       pb.hom_pos[m0:int(pb.hom_first[b + 1])].tolist())
        if plan is None:
            assert ref.het_lo[j] == ref.het_hi[j]; skipped += 1
            continue
        got = dict(ref_start=int(ref.ref_start[j]), ref_end=int(ref.ref_end[j]), het_lo=int(ref.het_lo[j]) - h0, het_hi=int(ref.het_hi[j]) - h0,
                   hom_lo=int(ref.hom_lo[j]) - m0 if ref.hom_hi[j] > ref.hom_lo[j] else 0, hom_hi=int(ref.hom_hi[j]) - m0 if ref.hom_hi[j] > ref.hom_lo[j] else 0,
                   read_start=int(ref.read_start[j]), read_end=int(ref.read_end[j]))
        assert got == plan, (j, got, plan)
    assert skipped > 0 and skipped < len(reads)


@pytest.mark.gpu
def test_gpu_plan_matches_the_oracle():
    ctx = lib.Context(device=0)
    pb, _ = _plan_inputs(4, 950, window=9000, n_var=12, n_hom=9, n_reads=60, read_lo=300, read_hi=3000, sv_max=300)
    ref = O.wfa_plan_batch(pb)
    out = ctx.wfa_plan_batch(pb)
    for f in A.PlanOut.FIELDS:
        assert np.array_equal(getattr(out, f), getattr(ref, f)), f
    with pytest.raises(lib.HiPhaseB200Error):                           # a mapping without aligned pairs: the reference asserts
        bad = A.PlanBatch([0], [0, 0], [], [], [], [0, 1], [5], [0, 0], [])
        ctx.wfa_plan_batch(bad)
    ctx.close()
