"""Matrix assembly: rows -> ReadSegment::new -> ReadSegment::collapse -> min-matched-alleles filter -> A* block batch
(src/data_types/read_segments.rs:40-121, 151-155; src/read_parsing.rs:612-629)."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import golden
from hiphase_b200 import _abi as A
from hiphase_b200 import synth


def _reads(asm):
    nr = int(asm.as_struct().n_reads)
    out = []
    for r in range(nr):
        c0, c1 = int(asm.cell_off[r]), int(asm.cell_off[r + 1])
        out.append((int(asm.read_start[r]), int(asm.read_end[r]), asm.alleles[c0:c1].tolist(), asm.quals[c0:c1].tolist()))
    return out


def _random_rows(seed, n_blocks=6, n_var=80, n_groups=25):
    rng = np.random.default_rng(seed)
    blocks = []
    for _ in range(n_blocks):
        groups = []
        for _ in range(n_groups):
            k = int(rng.choice([1, 1, 1, 2, 2, 3]))
            base = int(rng.integers(0, n_var - 10))
            grp = []
            truth = rng.integers(0, 2, n_var)
            for _ in range(k):
                s = max(0, base + int(rng.integers(-8, 9)))
                ln = int(rng.integers(0, min(30, n_var - s) + 1))
                a = truth[s:s + ln].astype(np.uint8)
                u = rng.random(ln)
                a = np.where(u < 0.15, 2, np.where(u < 0.3, 3, np.where(u < 0.36, 1 - a, a))).astype(np.uint8)
                q = np.where(a < 2, rng.integers(1, 200, ln), np.where(rng.random(ln) < 0.5, 0, rng.integers(0, 50, ln))).astype(np.uint8)
                grp.append((s, a, q))
            groups.append(grp)
        blocks.append(dict(n_var=n_var, groups=groups))
    return blocks


def test_oracle_collapse_golden_through_the_batch_entry():
    g = golden("read_segments.json")["collapse"]
    rows = A.RowsBatch([dict(n_var=7, groups=[[(0, a, q) for a, q in zip(g["rows_alleles"], g["rows_quals"])],
                                               [(0, g["rows_alleles"][0], g["rows_quals"][0])]])])
    o = O.assemble_blocks(rows)
    assert o.rc == 0
    s, e = g["region"]
    first, single = _reads(o)
    assert first == (s, e, g["expected_alleles"][s:e], g["expected_quals"][s:e])
    # "stupid collapsing": a single mapping collapses to itself, clipped to its set region (read_segments.rs:40-62)
    a0 = g["rows_alleles"][0]
    s1 = min(i for i, x in enumerate(a0) if x < 2); e1 = max(i for i, x in enumerate(a0) if x < 2) + 1
    assert single == (s1, e1, a0[s1:e1], g["rows_quals"][0][s1:e1])


def test_oracle_hand_cases():
    rows = A.RowsBatch([dict(n_var=10, groups=[
        [(2, [2, 0, 1, 2], [9, 5, 6, 9])],                              # Ambiguous at the ends is clipped away
        [(0, [0, 3, 3, 3], [7, 0, 0, 0])],                              # one set allele: phasable only (min 2)
        [(0, [2, 3, 2], [0, 0, 0])],                                    # nothing set: dropped
        [(1, [0, 1], [4, 4]), (2, [0, 1], [9, 3])],                     # conflict at 2 -> Ambiguous/0; 1 and 3 from one row each
        [(1, [2, 0, 2, 1], [0, 3, 8, 3]), (2, [0, 1], [5, 5])],         # a row's inner Ambiguous keeps its slot: first non-NoOverlap wins
        [(4, [1, 1], [0, 2]), (4, [1, 0], [0, 2])],                     # equal alleles, both quality 0: the reference asserts
    ])])
    o = O.assemble_blocks(rows)
    assert o.rc == 0
    assert o.group_class[:6].tolist() == [A.HP_GROUP_KEPT, A.HP_GROUP_PHASABLE, A.HP_GROUP_DROPPED, A.HP_GROUP_KEPT, A.HP_GROUP_KEPT, A.HP_GROUP_ASSERT]
    assert o.group_num_set[:5].tolist() == [2, 1, 0, 2, 2]
    assert _reads(o) == [(3, 5, [0, 1], [5, 6]), (1, 4, [0, 2, 1], [4, 0, 3]), (2, 5, [0, 2, 1], [5, 8, 3])]


def test_oracle_matches_python_glue_on_single_mapping_groups():
    """synth.blocks_from_wfa_rows (the Python glue used by the pipeline tests) == the oracle for one mapping per read."""
    blocks = _random_rows(3, n_groups=30)
    for b in blocks:
        b["groups"] = [g[:1] for g in b["groups"]]
    o = O.assemble_blocks(A.RowsBatch(blocks))
    got = _reads(o)
    exp = []
    for b in blocks:
        for (s, a, q), in b["groups"]:
            st = np.flatnonzero(np.asarray(a) < 2)
            if len(st) >= 2:
                lo, hi = int(st[0]), int(st[-1]) + 1
                exp.append((s + lo, s + hi, list(map(int, a[lo:hi])), list(map(int, q[lo:hi]))))
    assert got == exp


@pytest.fixture(scope="module")
def ctx():
    from hiphase_b200 import lib
    c = lib.Context(device=0)
    yield c
    c.close()


def _same(out, ref):
    assert int(out.as_struct().n_reads) == int(ref.as_struct().n_reads) and int(out.as_struct().n_cells) == int(ref.as_struct().n_cells)
    nr, nc = int(ref.as_struct().n_reads), int(ref.as_struct().n_cells)
    assert np.array_equal(out.read_off, ref.read_off)
    assert np.array_equal(out.read_start[:nr], ref.read_start[:nr]) and np.array_equal(out.read_end[:nr], ref.read_end[:nr])
    assert np.array_equal(out.cell_off[:nr + 1], ref.cell_off[:nr + 1])
    assert np.array_equal(out.alleles[:nc], ref.alleles[:nc]) and np.array_equal(out.quals[:nc], ref.quals[:nc])
    assert np.array_equal(out.group_class, ref.group_class) and np.array_equal(out.group_num_set, ref.group_num_set)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,min_set", [(1, 2), (2, 1), (5, 4)])
def test_gpu_assemble_matches_oracle(ctx, seed, min_set):
    rows = A.RowsBatch(_random_rows(seed), min_matched_alleles=min_set)
    ref = O.assemble_blocks(rows)
    assert ref.rc == 0 and int(ref.as_struct().n_reads) > 20
    _same(ctx.assemble_blocks(rows), ref)


@pytest.mark.gpu
def test_gpu_assemble_edge_cases(ctx):
    g = golden("read_segments.json")["collapse"]
    for blocks in ([dict(n_var=7, groups=[[(0, a, q) for a, q in zip(g["rows_alleles"], g["rows_quals"])]])],
                   [dict(n_var=5, groups=[])], [dict(n_var=5, groups=[[]])],
                   [dict(n_var=10, groups=[[(4, [1, 1], [0, 2]), (4, [1, 0], [0, 2])], [(0, [], [])], [(9, [1], [3])]])]):
        rows = A.RowsBatch(blocks)
        _same(ctx.assemble_blocks(rows), O.assemble_blocks(rows))


@pytest.mark.gpu
def test_gpu_pipeline_wfa_rows_to_astar(ctx):
    """WFA rows (GPU) -> matrix assembly (GPU) -> A* (GPU) equals the same chain through the oracle."""
    batch, jb, meta = synth.config_c4(2, window=30000, n_het=30, n_hom=40, n_reads=40, read_lo=4000, read_hi=9000, sv_max=600)
    out = ctx.wfa_align_batch(batch)
    blocks = [dict(n_var=m["n_het"], groups=[]) for m in meta]
    for j in range(batch.n_jobs):
        if out.status[j] != 0:
            continue
        m = meta[jb[j]]
        r0, r1 = int(batch.row_off[j]), int(batch.row_off[j + 1])
        blocks[jb[j]]["groups"].append([(int(batch.het_lo[j]) - m["het_base"], out.alleles[r0:r1], out.quals[r0:r1])])
    rows = A.RowsBatch(blocks)
    asm = ctx.assemble_blocks(rows)
    ref = O.assemble_blocks(rows)
    _same(asm, ref)
    is_snv = np.concatenate([(np.array(m["vtypes"]) == 0).astype(np.uint8) for m in meta])
    bb = asm.block_batch(rows, is_snv=is_snv)
    got = ctx.astar_solve_batch(bb)
    exp = O.astar_solve(bb, want_heuristic=False, want_counters=False)
    assert np.array_equal(got.h1, exp.h1) and np.array_equal(got.h2, exp.h2) and np.array_equal(got.stats, exp.stats)
    # and it is the batch the Python glue of the older pipeline test builds
    old = synth.blocks_from_wfa_rows(batch, out, jb, meta)
    for f in ("read_off", "read_start", "read_end", "cell_off", "alleles", "quals"):
        assert np.array_equal(getattr(bb, f), getattr(old, f)), f
