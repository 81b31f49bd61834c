"""Local realignment: local_realignment (src/read_parsing.rs:121-503), Variant::match_allele /
closest_allele_clip (variants.rs:598-641), sequence_alignment::edit_distance (sequence_alignment.rs:6-38).

CPU tests pin the oracle on the reference's known-answer vectors (tests/golden/local_realign.json) and cross-check it
against an independent Python restatement; GPU tests compare the CUDA path with the oracle through the C ABI."""
import numpy as np
import pytest

import oracle_lib as O
import pyref
from helpers import golden
from hiphase_b200 import _abi as A
from hiphase_b200 import synth
from hiphase_b200.variants import Variant, VariantType, variant_table

_VT = {"snv": 0, "insertion": 1, "deletion": 2, "indel": 3, "sv_insertion": 4, "sv_deletion": 5, "tr": 9}


def _variant(spec):
    t, pos, ref_len, a0, a1, i0, i1 = spec
    return Variant(0, _VT[t], pos, ref_len, a0.encode(), a1.encode(), i0, i1)


def _table_batch(variants):
    """A LocalBatch that only carries the variant table (for match_allele / closest_allele probes)."""
    return A.LocalBatch(variant_table(variants), [0], [len(variants)], [0], [0, 0], [], [], [], np.zeros(1, np.uint8),
                        np.zeros(1, np.uint8), [0, 0])


def _single_job(variants, read_pos, segs, seq, quals):
    return A.LocalBatch(variant_table(variants), [0], [len(variants)], [read_pos], [0, len(segs)], [s[0] for s in segs],
                        [s[1] for s in segs], [s[2] for s in segs], np.frombuffer(bytes(seq), np.uint8),
                        np.asarray(quals, np.uint8), [0, len(seq)])


def _py_variants(batch, lo, hi):
    out = []
    for k in range(lo, hi):
        a0 = bytes(batch.allele_bytes[int(batch.allele0_off[k]): int(batch.allele0_off[k]) + int(batch.allele0_len[k])])
        a1 = bytes(batch.allele_bytes[int(batch.allele1_off[k]): int(batch.allele1_off[k]) + int(batch.allele1_len[k])])
        out.append(dict(pos=int(batch.position[k]), ref_len=int(batch.ref_len[k]), prefix_len=int(batch.prefix_len[k]),
                        postfix_len=int(batch.postfix_len[k]), a0=a0, a1=a1, vtype=int(batch.vtype[k]), ignored=int(batch.ignored[k])))
    return out


# ---------------------------------------------------------------- CPU: oracle vs the reference's golden vectors
def test_oracle_edit_distance_golden():
    for a, b, d in golden("local_realign.json")["edit_distance"]:
        assert O.edit_distance(a, b) == d
        assert pyref.py_edit_distance(a, b) == d


def test_oracle_match_allele_golden():
    g = golden("local_realign.json")
    for spec, cases in g["match_allele"]:
        v = _variant(spec)
        b = _table_batch([v])
        for obs, exp in cases:
            assert O.match_allele(b, 0, obs.encode()) == exp
            assert v.match_allele(obs.encode()) == exp


def test_oracle_reference_adjustment_golden():
    g = golden("local_realign.json")["reference_adjustment"]
    v = _variant(g["variant"])
    assert v.get_prefix_len() == 0 and v.get_postfix_len() == 0
    v.add_reference_prefix(g["prefix"].encode()); v.add_reference_postfix(g["postfix"].encode())
    assert v.get_truncated_allele0() == g["truncated_allele0"].encode() and v.get_truncated_allele1() == g["truncated_allele1"].encode()
    v.truncate_reference_postfix(g["truncate"])
    assert v.get_type() == VariantType.Indel and v.position() == 20 and v.get_ref_len() == 2
    assert v.get_prefix_len() == g["prefix_len"] and v.get_postfix_len() == g["postfix_len"]
    b = _table_batch([v])
    for obs, exp in g["match_allele"]:
        assert O.match_allele(b, 0, obs.encode()) == exp
    for obs, exp in g["closest_allele"]:
        assert list(O.closest_allele_clip(b, 0, obs.encode())) == exp


def test_oracle_hand_cases():
    """Small scenarios whose answers follow directly from the text of read_parsing.rs."""
    ref = b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"          # 40 bases
    snv = Variant(0, 0, 10, 1, b"G", b"T")                    # ref[10] == 'G'
    snv.add_reference_prefix(ref[7:10]); snv.add_reference_postfix(ref[11:14])
    # read == reference over [0, 40): exact REF; all base qualities 40 -> factor 1 -> SNV_QUAL 80
    o = O.local_realign(_single_job([snv], 0, [(0, 0, 40)], ref, [40] * 40))
    assert (o.alleles[0], o.quals[0], o.match_class[0], o.status[0]) == (0, 80, 3, 0)
    # base qualities 20 -> harmonic mean 20 -> factor 0.5 -> 40 (:293-327)
    o = O.local_realign(_single_job([snv], 0, [(0, 0, 40)], ref, [20] * 40))
    assert (o.alleles[0], o.quals[0]) == (0, 40)
    # ALT base in the read
    alt = bytearray(ref); alt[10] = ord("T")
    o = O.local_realign(_single_job([snv], 0, [(0, 0, 40)], alt, [40] * 40))
    assert (o.alleles[0], o.quals[0], o.match_class[0]) == (1, 80, 3)
    # a third base: equidistant -> Ambiguous, inexact, but it overlaps and keeps its quality (:280-287)
    oth = bytearray(ref); oth[10] = ord("A")
    o = O.local_realign(_single_job([snv], 0, [(0, 0, 40)], oth, [40] * 40))
    assert (o.alleles[0], o.quals[0], o.match_class[0]) == (2, 80, 1)
    assert o.edit_distance[:2].tolist() == [1, 1]
    # a zero base quality: 1/0 = inf -> harmonic 0 -> the floor of 1 (:327)
    o = O.local_realign(_single_job([snv], 0, [(0, 0, 40)], ref, [40] * 9 + [0] + [40] * 30))
    assert (o.alleles[0], o.quals[0]) == (0, 1)
    # the read stops before the variant: NoOverlap; it covers the position but not the anchors: Ambiguous (:340-349)
    o = O.local_realign(_single_job([snv], 0, [(0, 0, 5)], ref[:5], [40] * 5))
    assert (o.alleles[0], o.quals[0], o.match_class[0]) == (3, 0, 0)
    # ignored variants are NoOverlap whatever the read says (:179-185)
    ign = Variant(0, 0, 10, 1, b"G", b"T"); ign.set_ignored()
    o = O.local_realign(_single_job([ign], 0, [(0, 0, 40)], ref, [40] * 40))
    assert (o.alleles[0], o.quals[0], o.match_class[0]) == (3, 0, 0)
    # SV deletion of ref[21..30): fully deleted in the read -> ALT, SV_INDEL_QUAL 20, exact; the SNV inside it is masked
    # to Ambiguous (:186-193, :399-428); the same read without the gap -> REF with quality 20
    svd = Variant(0, 5, 20, 10, ref[20:30], ref[20:21])
    inner = Variant(0, 0, 25, 1, b"C", b"A")
    read = ref[:21] + ref[30:]
    o = O.local_realign(_single_job([snv, svd, inner], 0, [(0, 0, 21), (30, 21, 10)], read, [40] * len(read)))
    assert o.alleles.tolist() == [0, 1, 2] and o.quals.tolist() == [80, 20, 0] and o.match_class.tolist() == [3, 3, 1]
    o = O.local_realign(_single_job([snv, svd, inner], 0, [(0, 0, 40)], ref, [40] * 40))
    assert o.alleles.tolist() == [0, 0, 0] and o.quals.tolist() == [80, 20, 80]
    # an unhandled variant type is a panic in the reference (:452-454)
    dup = Variant(0, 6, 10, 1, b"G", b"GG")
    assert O.local_realign(_single_job([dup], 0, [(0, 0, 40)], ref, [40] * 40)).status[0] == A.HP_LOCAL_UNHANDLED_TYPE


def test_oracle_vs_python_restatement():
    batch = synth.config_local(2, full_rows=True, n_reads=25, sv_max=300)
    o = O.local_realign(batch)
    assert o.rc == 0
    for j in range(batch.n_jobs):
        s0, s1 = int(batch.seg_off[j]), int(batch.seg_off[j + 1])
        segs = [(int(batch.seg_ref_start[s]), int(batch.seg_read_start[s]), int(batch.seg_len[s])) for s in range(s0, s1)]
        r0, r1 = int(batch.read_off[j]), int(batch.read_off[j + 1])
        a, q, c, st = pyref.py_local_realignment(_py_variants(batch, int(batch.var_lo[j]), int(batch.var_hi[j])), int(batch.read_pos[j]),
                                                 segs, batch.read_bytes[r0:r1], batch.read_quals[r0:r1].tolist())
        c0, c1 = int(batch.row_off[j]), int(batch.row_off[j + 1])
        assert o.alleles[c0:c1].tolist() == a and o.quals[c0:c1].tolist() == q and o.match_class[c0:c1].tolist() == c
        assert o.status[j] == st


# ---------------------------------------------------------------- GPU: CUDA path vs oracle, through the C ABI
@pytest.fixture(scope="module")
def ctx():
    from hiphase_b200 import lib
    c = lib.Context(device=0)
    yield c
    c.close()


def _same(out, ref):
    assert np.array_equal(out.status, ref.status)
    assert np.array_equal(out.alleles, ref.alleles)
    assert np.array_equal(out.quals, ref.quals)
    assert np.array_equal(out.match_class, ref.match_class)
    assert np.array_equal(out.edit_distance, ref.edit_distance)


@pytest.mark.gpu
def test_gpu_edit_distance_golden(ctx):
    g = golden("local_realign.json")["edit_distance"]
    d = ctx.edit_distance_batch([(a, b) for a, b, _ in g])
    assert d.tolist() == [x for _, _, x in g]


@pytest.mark.gpu
def test_gpu_edit_distance_random(ctx):
    rng = np.random.default_rng(7)
    pairs = []
    for la, lb, alpha in [(0, 0, 4), (0, 9, 4), (1, 1, 4), (63, 64, 4), (64, 65, 4), (65, 65, 4), (65, 300, 4), (128, 129, 4), (200, 70, 5),
                          (640, 700, 4), (2047, 2049, 4), (2100, 2500, 4), (4100, 4097, 4), (300, 16000, 4), (30, 5000, 6)] + \
                         [(int(rng.integers(0, 400)), int(rng.integers(0, 400)), int(rng.integers(2, 7))) for _ in range(60)]:
        letters = np.frombuffer(b"ACGTNX", np.uint8)[:alpha]
        a = letters[rng.integers(0, alpha, la)]
        if rng.random() < 0.6 and la and lb:          # b = a mutated copy of a, so that distances are not saturated
            b = a.copy()
            idx = rng.random(len(b)) < 0.05
            b[idx] = letters[rng.integers(0, alpha, int(idx.sum()))]
            b = np.delete(b, np.flatnonzero(rng.random(len(b)) < 0.03))
            b = np.concatenate([b, letters[rng.integers(0, alpha, max(0, lb - len(b)))]])[: max(lb, 1)]
        else:
            b = letters[rng.integers(0, alpha, lb)]
        pairs.append((a, b))
    d = ctx.edit_distance_batch(pairs)
    for (a, b), x in zip(pairs, d):
        assert int(x) == O.edit_distance(a, b), (len(a), len(b))


def _mutated(rng, a, p_sub=0.01, p_del=0.003, p_ins=0.003):
    letters = np.frombuffer(b"ACGT", np.uint8)
    b = a.copy()
    idx = rng.random(len(b)) < p_sub
    b[idx] = letters[rng.integers(0, 4, int(idx.sum()))]
    b = np.delete(b, np.flatnonzero(rng.random(len(b)) < p_del))
    for pos in sorted(rng.integers(0, len(b), int(len(b) * p_ins)), reverse=True):
        b = np.insert(b, pos, letters[rng.integers(0, 4)])
    return b


@pytest.mark.gpu
def test_gpu_edit_distance_beyond_one_panel(ctx):
    """Both sequences longer than one panel of the bit-vector kernel (16384 pattern rows): two and three panels, a pattern that
    ends exactly on a panel boundary, identical sequences.  The reference's full grid has no length limit
    (sequence_alignment.rs:6-38)."""
    rng = np.random.default_rng(99)
    letters = np.frombuffer(b"ACGT", np.uint8)
    a1 = letters[rng.integers(0, 4, 16385 + 64)]
    a2 = letters[rng.integers(0, 4, 20000)]
    a3 = letters[rng.integers(0, 4, 33100)]
    a4 = letters[rng.integers(0, 4, 2 * 16384)]
    pairs = [(a1, a1), (a2, _mutated(rng, a2)), (_mutated(rng, a3), a3), (a4, np.concatenate([_mutated(rng, a4), a2[:300]])),
             (a2, letters[rng.integers(0, 4, 17000)])]
    d = ctx.edit_distance_batch(pairs)
    assert int(d[0]) == 0
    for (a, b), x in zip(pairs[1:], d[1:]):
        assert int(x) == O.edit_distance(a, b), (len(a), len(b))


@pytest.mark.gpu
def test_gpu_local_realign_sv_insertion_beyond_one_panel(ctx):
    """A 17 kb SV insertion inside a read that carries a noisy copy of it: closest_allele_clip compares two sequences that are
    both longer than one panel (variants.rs:598-641)."""
    rng = np.random.default_rng(17)
    letters = np.frombuffer(b"ACGT", np.uint8)
    ref = letters[rng.integers(0, 4, 240)].tobytes()
    ins = letters[rng.integers(0, 4, 17000)]
    sv = Variant(0, 4, 100, 1, ref[100:101], ref[100:101] + ins.tobytes())
    snv = Variant(0, 0, 180, 1, ref[180:181], b"A" if ref[180:181] != b"A" else b"C")
    got = _mutated(rng, ins).tobytes()
    read = ref[:101] + got + ref[101:]
    job = _single_job([sv, snv], 0, [(0, 0, 101), (101, 101 + len(got), len(ref) - 101)], read, [30] * len(read))
    ref_out = O.local_realign(job)
    assert int(ref_out.alleles[0]) == 1 and 0 < int(ref_out.edit_distance.reshape(-1, 2)[0, 1]) < 2000
    _same(ctx.local_realign_batch(job), ref_out)


@pytest.mark.gpu
def test_gpu_variant_mirror_golden():
    g = golden("local_realign.json")["reference_adjustment"]
    v = _variant(g["variant"])
    v.add_reference_prefix(g["prefix"].encode()); v.add_reference_postfix(g["postfix"].encode()); v.truncate_reference_postfix(g["truncate"])
    for obs, exp in g["closest_allele"]:
        assert list(v.closest_allele(obs.encode())) == exp


@pytest.mark.gpu
@pytest.mark.parametrize("full_rows", [False, True])
def test_gpu_local_realign_matches_oracle(ctx, full_rows):
    batch = synth.config_local(6, full_rows=full_rows)
    ref = O.local_realign(batch)
    out = ctx.local_realign_batch(batch)
    _same(out, ref)
    assert (ref.match_class & 1).sum() > 1000 and (ref.edit_distance.reshape(-1, 2).sum(1) > 0).sum() > 100


@pytest.mark.gpu
def test_gpu_local_realign_hand_cases(ctx):
    ref = b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"
    snv = Variant(0, 0, 10, 1, b"G", b"T")
    snv.add_reference_prefix(ref[7:10]); snv.add_reference_postfix(ref[11:14])
    svd = Variant(0, 5, 20, 10, ref[20:30], ref[20:21])
    inner = Variant(0, 0, 25, 1, b"C", b"A")
    dup = Variant(0, 6, 10, 1, b"G", b"GG")
    read = ref[:21] + ref[30:]
    for job in (_single_job([snv, svd, inner], 0, [(0, 0, 21), (30, 21, 10)], read, [40] * len(read)),
                _single_job([snv, svd, inner], 0, [(0, 0, 40)], ref, [40] * 40),
                _single_job([snv], 0, [(0, 0, 40)], ref, [40] * 9 + [0] + [40] * 30),
                _single_job([snv], 0, [(0, 0, 5)], ref[:5], [40] * 5),
                _single_job([snv, svd, inner], 3, [(12, 0, 5), (19, 5, 1), (33, 6, 7)], ref[12:17] + ref[19:20] + ref[33:40], [33] * 13),
                _single_job([dup], 0, [(0, 0, 40)], ref, [40] * 40),
                _single_job([svd, dup], 0, [(0, 0, 21), (30, 21, 10)], read, [40] * len(read))):
        _same(ctx.local_realign_batch(job), O.local_realign(job))


@pytest.mark.gpu
def test_gpu_local_realign_sv_heavy(ctx):
    """Many SV insertions / deletions: long x long edit distances (warp-systolic path) and deletion masking."""
    batch = synth.config_local(3, full_rows=True, n_var=60, sv_max=3000, err=0.01)
    _same(ctx.local_realign_batch(batch), O.local_realign(batch))


@pytest.mark.gpu
def test_gpu_local_realignment_mirror():
    """read_parsing.local_realignment(read, variants) with the reference's argument order; a CIGAR-described read."""
    from hiphase_b200.read_parsing import AlignedRead, local_realignment
    ref = b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT"
    snv = Variant(0, 0, 10, 1, b"G", b"T")
    snv.add_reference_prefix(ref[7:10]); snv.add_reference_postfix(ref[11:14])
    svd = Variant(0, 5, 20, 10, ref[20:30], ref[20:21])
    inner = Variant(0, 0, 25, 1, b"C", b"A")
    read = AlignedRead.from_cigar(0, [("S", 2), ("M", 21), ("D", 9), ("M", 10)], b"TT" + ref[:21] + ref[30:], [40] * 33)
    assert read.segments == [(0, 2, 21), (30, 23, 10)]
    alleles, quals, stats = local_realignment(read, [snv, svd, inner])
    assert alleles == [0, 1, 2] and quals == [80, 20, 0]
    assert stats.exact_matches[0] == 1 and stats.exact_matches[5] == 1 and stats.failed_matches[0] == 1
    assert stats.allele0_matches[0] == 1 and stats.allele1_matches[5] == 1 and stats.num_alleles == 2 and stats.local_aligned == 1
    with pytest.raises(RuntimeError):
        local_realignment(read, [Variant(0, 6, 10, 1, b"G", b"GG")])


def test_plan_global_realignment_window():
    """read_parsing.rs:672-742: window, overlap index ranges and read slice from the aligned segments."""
    from hiphase_b200.read_parsing import AlignedRead, plan_global_realignment
    read = AlignedRead.from_cigar(100, [("S", 5), ("M", 20), ("D", 10), ("M", 30), ("I", 4), ("M", 6), ("S", 3)], b"A" * 68, [30] * 68)
    assert read.segments == [(100, 5, 20), (130, 25, 30), (160, 59, 6)]
    plan = plan_global_realignment(read, [50, 100, 125, 165, 166, 400], [10, 120, 165, 170])
    assert plan == dict(ref_start=100, ref_end=166, het_lo=1, het_hi=4, hom_lo=1, hom_hi=3, read_start=5, read_end=65)
    assert plan_global_realignment(read, [50, 99, 166], [120]) is None          # no het overlap: the job is skipped (:703-712)
    assert plan_global_realignment(read, [110], [])["hom_lo"] == 0             # first_hom_overlap.unwrap_or(0)
