"""The CPU oracle against every known-answer vector the reference's unit tests hold for the hot path
(tests/golden/*.json, transcribed from src/data_types/read_segments.rs:213-308, src/astar_phaser.rs:662-798,
src/wfa_graph.rs:676-1208)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from helpers import golden, wfa_batch_single
from hiphase_b200 import _abi as A


def test_read_segment_constructor():
    for c in golden("read_segments.json")["constructor"]:
        s, e = O.region(c["alleles"])
        assert [s, e] == c["region"]
        assert c["alleles"][s:e] == c["clipped_alleles"] and c["quals"][s:e] == c["clipped_quals"]


@pytest.mark.parametrize("key", ["score_haplotype", "score_partial_haplotype"])
def test_read_segment_scores(key):
    g = golden("read_segments.json")[key]
    s, e = O.region(g["alleles"])
    if "region" in g:
        assert [s, e] == g["region"]
        assert sum(1 for a in g["alleles"][s:e] if a < 2) == g["num_set"]
    for c in g["cases"]:
        assert O.score_partial(s, g["alleles"][s:e], g["quals"][s:e], c["hap"], c["offset"]) == c["score"]


def test_read_segment_collapse():
    g = golden("read_segments.json")["collapse"]
    oa, oq, reg = O.collapse(g["rows_alleles"], g["rows_quals"])
    assert list(reg) == g["region"]
    # outside the region the oracle reports NoOverlap/0, like ReadSegment::allele()/qual()
    assert oa.tolist() == g["expected_alleles"] and oq.tolist() == g["expected_quals"]
    s, e = reg
    assert O.score_partial(s, oa[s:e], oq[s:e], g["score_case"]["hap"], 0) == g["score_case"]["score"]
    # "stupid collapsing": a single mapping collapses to itself
    oa1, oq1, reg1 = O.collapse(g["rows_alleles"][:1], g["rows_quals"][:1])
    s1, e1 = O.region(g["rows_alleles"][0])
    assert (s1, e1) == reg1 and oa1[s1:e1].tolist() == g["rows_alleles"][0][s1:e1]


def test_astar_node_paths():
    g = golden("astar_node.json")
    batch = A.BlockBatch.from_blocks([{"n_var": g["n"], "reads": [(r["start"], r["alleles"], r["quals"]) for r in g["reads"]]}])
    H = np.array(g["H"], np.uint64)
    for p in g["paths"]:
        n = len(p["a1"])
        a1, a2 = np.array(p["a1"], np.uint8), np.array(p["a2"], np.uint8)
        frozen, total, hets = (np.zeros(n, np.uint64) for _ in range(3))
        bs = batch.as_struct()
        rc = O.lib().hpo_astar_node_path(C.byref(bs), A.ptr(H, A.u64p), n, A.ptr(a1, A.u8p), A.ptr(a2, A.u8p),
                                         A.ptr(frozen, A.u64p), A.ptr(total, A.u64p), A.ptr(hets, A.u64p))
        assert rc == 0
        assert frozen.tolist() == p["frozen"], p["name"]
        assert total.tolist() == p["total"], p["name"]
        assert hets.tolist() == p["hets"], p["name"]


def test_tracker():
    g = golden("astar_node.json")["tracker"]
    ops = np.array([o for o, _ in g["ops"]], np.uint8)
    vals = np.array([v for _, v in g["ops"]], np.uint32)
    lens = np.zeros(len(ops), np.uint64)
    assert O.lib().hpo_tracker_script(g["max_len"], len(ops), A.ptr(ops, A.u8p), A.ptr(vals, A.u32p), A.ptr(lens, A.u64p)) == 0
    assert lens.tolist() == g["lens"]


@pytest.mark.parametrize("case", golden("wfa_graphs.json")["hand_built"], ids=lambda c: c["name"])
def test_wfa_hand_built(case):
    g = O.Graph(1000)
    for i, n in enumerate(case["nodes"]):
        assert g.add_node(n["seq"], n["parents"]) == i
    for c in case["cases"]:
        for seed in (0, 1, 7):   # result must not depend on the diagonal visiting order
            st, score, nodes = g.edit_distance(c["read"], shuffle_seed=seed)
            assert st == A.HP_WFA_OK and score == c["score"]
            if "nodes" in c:
                assert nodes == c["nodes"]


def test_wfa_add_node_errors():
    g = O.Graph(1000)
    assert g.add_node([1], [0]) == -1        # first node must have no parents (wfa_graph.rs:304)
    assert g.add_node([1], []) == 0
    assert g.add_node([1], []) == -1         # later nodes need a parent (:309)
    assert g.add_node([1], [1]) == -1        # parent must precede (:314)
    assert g.add_node([1], [0]) == 1


@pytest.mark.parametrize("case", golden("wfa_graphs.json")["from_variants"], ids=lambda c: c["name"])
def test_wfa_from_variants(case):
    reads = [c["read"] for c in case["cases"]] or ["A"]
    batch = wfa_batch_single(case["reference"], case["hets"], case["homs"], case["ref_start"], case["ref_end"], reads)
    g = O.Graph.from_job(batch, 0, 1000)
    assert g is not None and g.num_nodes() == case["num_nodes"]
    amap = g.allele_map()
    assert {str(k): [list(x) for x in v] for k, v in amap.items()} == case["map"]
    for c in case["cases"]:
        for seed in (0, 3):
            st, score, nodes = g.edit_distance(c["read"].encode(), shuffle_seed=seed)
            assert (st, score, nodes) == (A.HP_WFA_OK, c["score"], c["nodes"])


def test_wfa_max_edit_distance():
    g = O.Graph(3)
    g.add_node([0, 1, 2, 3, 4, 5, 6, 7], [])
    st, score, nodes = g.edit_distance([9] * 8)
    assert st == A.HP_WFA_MAX_EDIT_DISTANCE and score == 3


def test_wfa_rows_from_traversal():
    # read_parsing.rs:790-835 on the complex 9-node case: alleles from traversed nodes, conflict -> Ambiguous, quals doubled
    case = [c for c in golden("wfa_graphs.json")["from_variants"] if c["name"] == "test_complex_problem"][0]
    reads = [c["read"] for c in case["cases"]]
    batch = wfa_batch_single(case["reference"], case["hets"], case["homs"], case["ref_start"], case["ref_end"], reads)
    out = O.wfa_align(batch, trav_words=1)
    assert out.failures == 0
    rows = out.alleles.reshape(len(reads), 3).tolist()
    quals = out.quals.reshape(len(reads), 3).tolist()
    # reference path: del0 ref(0), del1 ref(0), multi-allelic snv untouched (its REF node carries no allele) -> 3
    assert rows[0] == [0, 0, 3] and quals[0] == [20, 20, 0]
    assert rows[1] == [1, 3, 3] and quals[1] == [20, 0, 0]            # first deletion only: skips the others
    assert rows[3] == [0, 0, 0] and quals[3] == [20, 20, 160]         # third-0
    assert rows[4] == [0, 0, 1]                                       # third-1
    assert rows[7] == [2, 1, 3] and quals[7] == [0, 20, 0]            # nodes 0,1,2,3,7,8: var0 both alleles -> ambiguous
    assert rows[8] == [2, 1, 2] and quals[8] == [0, 20, 0]
    for j, c in enumerate(case["cases"]):
        assert out.score[j] == c["score"]
        assert [i for i in range(9) if out.traversed[j] >> i & 1] == c["nodes"]
