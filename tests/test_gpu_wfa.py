"""Parity of the CUDA graph-WFA path (through the C ABI) against the reference's known-answer vectors and the
CPU oracle: status, edit distance, traversed-node set and the allele / quality rows, bit-exact."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import golden, wfa_batch_single
from hiphase_b200 import _abi as A
from hiphase_b200 import lib, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = lib.Context(device=0)
    yield c
    c.close()


def _flat(nodes):
    seq, seq_off, par, par_off = [], [0], [], [0]
    for n in nodes:
        seq += list(n["seq"]); seq_off.append(len(seq))
        par += sorted(n["parents"]); par_off.append(len(par))
    return np.array(seq, np.uint8), seq_off, np.array(par, np.uint32), par_off


@pytest.mark.parametrize("case", golden("wfa_graphs.json")["hand_built"], ids=lambda c: c["name"])
def test_hand_built_graphs(ctx, case):
    seq, seq_off, par, par_off = _flat(case["nodes"])
    for c in case["cases"]:
        st, score, nodes = ctx.wfa_graph_align(seq, seq_off, par, par_off, c["read"])
        assert st == A.HP_WFA_OK and score == c["score"], (c, st, score)
        if "nodes" in c:
            assert nodes == c["nodes"]


def test_graph_rules_rejected(ctx):
    with pytest.raises(lib.HiPhaseB200Error):
        ctx.wfa_graph_align([1, 2], [0, 1, 2], [], [0, 0, 0], [1, 2])            # second node without a parent
    with pytest.raises(lib.HiPhaseB200Error):
        ctx.wfa_graph_align([1, 2], [0, 1, 2], [0, 1], [0, 1, 2], [1, 2])        # root with a parent


def test_max_edit_distance_status(ctx):
    st, score, nodes = ctx.wfa_graph_align(list(range(8)), [0, 8], [], [0, 0], [9] * 8, max_edit_distance=3)
    assert st == A.HP_WFA_MAX_EDIT_DISTANCE and score == 3 and nodes == []


@pytest.mark.parametrize("case", golden("wfa_graphs.json")["from_variants"], ids=lambda c: c["name"])
def test_from_variants_golden(ctx, case):
    reads = [c["read"] for c in case["cases"]] or ["A"]
    batch = wfa_batch_single(case["reference"], case["hets"], case["homs"], case["ref_start"], case["ref_end"], reads)
    out = ctx.wfa_align_batch(batch, trav_words=1)
    ref = O.wfa_align(batch, trav_words=1)
    assert (out.n_nodes == case["num_nodes"]).all()
    for j, c in enumerate(case["cases"]):
        assert out.status[j] == A.HP_WFA_OK and out.score[j] == c["score"]
        assert [i for i in range(case["num_nodes"]) if (int(out.traversed[j]) >> i) & 1] == c["nodes"]
    assert np.array_equal(out.alleles, ref.alleles) and np.array_equal(out.quals, ref.quals)
    assert np.array_equal(out.status, ref.status) and np.array_equal(out.score, ref.score)


def _assert_wfa_parity(out, ref):
    assert np.array_equal(out.status, ref.status), (np.flatnonzero(out.status != ref.status)[:10], out.status[:20], ref.status[:20])
    assert np.array_equal(out.score, ref.score)
    assert np.array_equal(out.n_nodes, ref.n_nodes)
    assert np.array_equal(out.traversed, ref.traversed)
    assert np.array_equal(out.alleles, ref.alleles) and np.array_equal(out.quals, ref.quals)


def test_c4_small_vs_oracle(ctx):
    batch, jb, meta = synth.config_c4(2, window=30000, n_het=30, n_hom=40, n_reads=24, read_lo=4000, read_hi=9000, sv_max=600)
    out = ctx.wfa_align_batch(batch, trav_words=8)
    ref = O.wfa_align(batch, threads=8, trav_words=8)
    assert ref.failures == 0
    _assert_wfa_parity(out, ref)
    assert (out.status == A.HP_WFA_OK).sum() > 10


def test_c4_noisy_reads_and_pruning_off(ctx):
    # high error rate: deep edit distances, wide wavefronts, max-ED failures; and the same with pruning disabled
    batch, jb, meta = synth.config_c4(1, window=12000, n_het=25, n_hom=25, n_reads=16, read_lo=1500, read_hi=3000, sv_max=300,
                                      err=0.03, p_noisy=0.3, err_noisy=0.25)
    for prune, max_ed in ((500, 500), (0, 120), (40, 500)):
        params = A.hp_params(1000, 3, prune, max_ed)
        c2 = lib.Context(params, device=0)
        out = c2.wfa_align_batch(batch, trav_words=4)
        ref = O.wfa_align(batch, params, threads=8, trav_words=4)
        _assert_wfa_parity(out, ref)
        c2.close()


def test_skipped_and_empty_jobs(ctx):
    case = [c for c in golden("wfa_graphs.json")["from_variants"] if c["name"] == "test_multiple_variants"][0]
    batch = wfa_batch_single(case["reference"], case["hets"], case["homs"], 0, 5, ["AAAAA", "", "ACACA"])
    batch.het_hi[1] = batch.het_lo[1]          # job 1 overlaps no het variant -> skipped (read_parsing.rs:703-712)
    batch = A.WfaBatch({k: getattr(batch, k) for k in ("position", "ref_len", "allele0_off", "allele0_len", "allele1_off", "allele1_len",
                                                        "index_allele0", "vtype", "ignored", "allele_bytes")},
                       batch.reference, batch.ref_start, batch.ref_end, batch.het_lo, batch.het_hi, batch.hom_lo, batch.hom_hi,
                       batch.read_bytes, batch.read_off)
    out = ctx.wfa_align_batch(batch, trav_words=1)
    ref = O.wfa_align(batch, trav_words=1)
    assert out.status.tolist() == [A.HP_WFA_OK, A.HP_WFA_SKIPPED, A.HP_WFA_OK] == ref.status.tolist()
    assert np.array_equal(out.alleles, ref.alleles) and np.array_equal(out.quals, ref.quals)


def test_c4_rows_feed_astar(ctx):
    # K2 -> matrix rows -> K1: the blocks assembled from the CUDA rows phase identically to the oracle pipeline
    batch, jb, meta = synth.config_c4(2, window=30000, n_het=30, n_hom=40, n_reads=40, read_lo=4000, read_hi=9000, sv_max=600)
    out = ctx.wfa_align_batch(batch)
    blocks = synth.blocks_from_wfa_rows(batch, out, jb, meta)
    got = ctx.astar_solve_batch(blocks)
    ref_rows = O.wfa_align(batch, threads=8)
    ref_blocks = synth.blocks_from_wfa_rows(batch, ref_rows, jb, meta)
    want = O.astar_solve(ref_blocks, threads=4)
    assert np.array_equal(got.h1, want.h1) and np.array_equal(got.h2, want.h2) and np.array_equal(got.stats, want.stats)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2], ids=["device-build", "host-build", "device-build-regrow"])
def test_graph_build_modes_agree(mode):
    """Row f3: from_reference_variants_with_hom on the device (default) vs the host builder vs the oracle: same node
    counts, traversed sets, scores and rows."""
    from hiphase_b200 import lib
    c = lib.Context(device=0)
    c.set_wfa_build_mode(mode)
    batch, jb, meta = synth.config_c4(2, window=30000, n_het=30, n_hom=40, n_reads=24, read_lo=4000, read_hi=9000, sv_max=600)
    ref = O.wfa_align(batch, threads=4, trav_words=8)
    out = c.wfa_align_batch(batch, trav_words=8)
    _assert_wfa_parity(out, ref)
    assert np.array_equal(out.n_nodes, ref.n_nodes) and np.array_equal(out.traversed, ref.traversed)
    c.close()


def _dense_snv_batch(n_het, n_reads, seed, spacing=3, p_err=0.004):
    """One window with n_het SNVs every `spacing` bases (3 nodes per variant: > 1024 nodes from 342 variants on) and reads
    drawn from the two haplotypes with a few errors."""
    rng = np.random.default_rng(seed)
    L = n_het * spacing + 40
    ref = rng.integers(0, 4, L)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    hets, truth = [], rng.integers(0, 2, n_het)
    for i in range(n_het):
        pos = 20 + i * spacing
        alt = (int(ref[pos]) + 1 + int(rng.integers(0, 3))) % 4
        hets.append({"type": "snv", "pos": pos, "ref_len": 1, "a0": bytes([acgt[ref[pos]]]), "a1": bytes([acgt[alt]])})
    reads = []
    for r in range(n_reads):
        h = r & 1
        seq = ref.copy()
        for i, v in enumerate(hets):
            if truth[i] ^ h:
                seq[v["pos"]] = b"ACGT".index(v["a1"])
        flip = rng.random(L) < p_err
        seq[flip] = (seq[flip] + 1) % 4
        reads.append(bytes(acgt[seq]))
    return wfa_batch_single(bytes(acgt[ref]), hets, [], 0, L, reads)


@pytest.mark.parametrize("n_het,words", [(360, 32), (1200, 64)], ids=["1082-nodes", "3602-nodes"])
def test_graphs_with_more_than_1024_nodes(ctx, n_het, words):
    """The reference's WFAGraph has no node limit (wfa_graph.rs:24-68); the kernel's node mask is one word per lane up to
    1024 nodes and four words per lane up to 4096 (round-1 review: these jobs used to be refused)."""
    batch = _dense_snv_batch(n_het, 6, seed=n_het)
    ref = O.wfa_align(batch, threads=6, trav_words=words)
    assert int(ref.n_nodes.max()) > 1024 and (ref.status == A.HP_WFA_OK).all()
    out = ctx.wfa_align_batch(batch, trav_words=words)
    _assert_wfa_parity(out, ref)


def test_graph_beyond_the_kernel_range_reports_its_own_status(ctx):
    batch = _dense_snv_batch(1400, 2, seed=5)                       # 4202 nodes
    out = ctx.wfa_align_batch(batch, trav_words=1)
    assert out.status.tolist() == [A.HP_WFA_GRAPH_TOO_LARGE] * 2


def test_device_builder_overflow_falls_back_to_the_host_builder(ctx):
    """More than 64 ALT branches open at one position (nested deletions): the device builder's fixed lists overflow and the
    job is rebuilt on the host in the retry pass instead of being refused."""
    rng = np.random.default_rng(3)
    L = 400
    acgt = np.frombuffer(b"ACGT", np.uint8)
    ref = rng.integers(0, 4, L)
    refb = bytes(acgt[ref])
    hets = []
    for i in range(70):                                              # 70 deletions, all spanning position 200
        pos = 100 + i
        rl = 150 - i + (i % 3)
        hets.append({"type": "deletion", "pos": pos, "ref_len": rl, "a0": refb[pos:pos + rl], "a1": refb[pos:pos + 1]})
    reads = [refb, refb[:100] + refb[249:]]
    batch = wfa_batch_single(refb, hets, [], 0, L, reads)
    ref_out = O.wfa_align(batch, threads=2, trav_words=4)
    out = ctx.wfa_align_batch(batch, trav_words=4)
    _assert_wfa_parity(out, ref_out)


def test_piece_filter_is_exact():
    """The piece filter answers MaxEditDistance for reads whose number of 10-base pieces that occur in no path of the graph
    exceeds max_edit_distance (a lower bound of the edit distance of wfa_graph.rs:350-650).  Filter on, filter off and the
    oracle agree on every output; the filter fires on the hopeless reads and on nothing else."""
    batch, jb, meta = synth.config_c4(2, window=20000, n_het=30, n_hom=30, n_reads=24, read_lo=2500, read_hi=6000, sv_max=300,
                                      err=0.01, p_noisy=0.35, err_noisy=0.3)
    # reads at the limit: edit distance close to max_edit_distance on either side (errors every ~18 bases over the whole read)
    for max_ed in (100, 200):
        params = A.hp_params(1000, 3, 500, max_ed)
        ref = O.wfa_align(batch, params, threads=8, trav_words=4)
        n_max = int((ref.status == A.HP_WFA_MAX_EDIT_DISTANCE).sum())
        assert n_max > 3 and (ref.status == A.HP_WFA_OK).sum() > 3
        outs = []
        for on in (1, 0):
            c = lib.Context(params, device=0)
            c.set_wfa_filter(on)
            out = c.wfa_align_batch(batch, trav_words=4)
            _assert_wfa_parity(out, ref)
            outs.append(c.wfa_filtered())
            c.close()
        assert outs[1] == 0 and 0 < outs[0] <= n_max, (outs, n_max)


def test_piece_filter_keeps_reads_near_the_limit():
    """Reads with one substitution in (almost) every piece are within max_edit_distance although most of their pieces are
    absent from the graph: the bound counts pieces, not bases, and must not fire below the limit."""
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    L = 3300                                              # 330 pieces > max_edit_distance 300: the filter is tried
    ref = rng.integers(0, 4, L)
    hets = [{"type": "snv", "pos": p, "ref_len": 1, "a0": bytes([acgt[ref[p]]]), "a1": bytes([acgt[(ref[p] + 1) % 4]])} for p in range(100, L - 100, 211)]
    reads = []
    for n_err in (280, 295, 300, 301, 305, 320):         # max_edit_distance 300: three on each side
        seq = ref.copy()
        pos = rng.choice(L // 10, n_err - 20, replace=False) * 10 + rng.integers(0, 10, n_err - 20)   # one per piece ...
        seq[pos] = (seq[pos] + 2) % 4
        extra = rng.choice(L, 20, replace=False)                                                  # ... and a few anywhere
        seq[extra] = (seq[extra] + 1) % 4
        reads.append(bytes(acgt[seq]))
    batch = wfa_batch_single(bytes(acgt[ref]), hets, [], 0, L, reads)
    params = A.hp_params(1000, 3, 500, 300)
    ref_out = O.wfa_align(batch, params, threads=6, trav_words=1)
    assert set(ref_out.status.tolist()) == {A.HP_WFA_OK, A.HP_WFA_MAX_EDIT_DISTANCE}
    c = lib.Context(params, device=0)
    out = c.wfa_align_batch(batch, trav_words=1)
    _assert_wfa_parity(out, ref_out)
    assert c.wfa_filtered() <= int((ref_out.status == A.HP_WFA_MAX_EDIT_DISTANCE).sum())
    c.close()
