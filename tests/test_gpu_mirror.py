"""The host-side mirrors of the reference interface (same names / argument meaning / error behaviour), written like
the reference's own unit tests (src/wfa_graph.rs:676-744, src/astar_phaser.rs:639-660)."""
import numpy as np
import pytest

import oracle_lib as O
from hiphase_b200 import _abi as A
from hiphase_b200.astar_phaser import astar_solver
from hiphase_b200.read_segments import AlleleType, ReadSegment
from hiphase_b200.wfa_graph import WFAGraph, WFAGraphError, WFAResult

pytestmark = pytest.mark.gpu


def test_basic_variant_like_the_reference_test():
    graph = WFAGraph()
    v1 = [0, 1, 2, 4, 5]
    graph.add_node(v1[0:2], [])
    graph.add_node([2], [0])
    graph.add_node([3], [0])
    graph.add_node(v1[3:], [1, 2])
    assert graph.edit_distance(v1) == WFAResult(0, [0, 1, 3])
    assert graph.edit_distance([0, 1, 3, 4, 5]) == WFAResult(0, [0, 2, 3])
    assert graph.edit_distance([1, 2, 3, 5]) == WFAResult(2, [0, 1, 3])
    assert graph.edit_distance([]) == WFAResult(5, [0, 1, 2, 3])
    assert graph.edit_distance([0, 1, 4, 5]) == WFAResult(1, [0, 1, 2, 3])


def test_add_node_errors_and_max_ed():
    g = WFAGraph(3)
    with pytest.raises(ValueError):
        g.add_node([1], [0])
    g.add_node(list(range(8)), [])
    with pytest.raises(ValueError):
        g.add_node([1], [])
    with pytest.raises(WFAGraphError) as e:
        g.edit_distance([9] * 8)
    assert e.value.distance == 3


def test_astar_solver_simple_reads():
    # get_simple_reads (astar_phaser.rs:642-660): an all-0 read with qual 2 and an all-1 read with qual 3
    n = 6
    reads = [ReadSegment("read_name", [AlleleType.Reference] * n, [2] * n), ReadSegment("read_name_2", [AlleleType.Alternate] * n, [3] * n)]
    variants = [{"ignored": False, "is_snv": True} for _ in range(n)]
    res = astar_solver(variants, reads, 1000, 3)
    assert res.haplotype_1.tolist() == [0] * n and res.haplotype_2.tolist() == [1] * n
    assert res.statistics.actual_cost == 0 and res.statistics.phased_variants == n and res.statistics.phased_snvs == n
    ref = O.astar_solve(A.BlockBatch.from_blocks([{"n_var": n, "reads": [(r.start, r.alleles, r.quals) for r in reads]}]))
    assert ref.h1.tolist() == res.haplotype_1.tolist() and int(ref.stats[0]["estimated_cost"]) == res.statistics.estimated_cost


def test_read_segment_constructor_clipping():
    rs = ReadSegment("read_name", [3, 0, 1, 0, 0, 1, 2, 2, 3, 3], [0, 1, 2, 3, 4, 5, 6, 7, 0, 0])
    assert (rs.start, rs.end) == (1, 6) and rs.alleles.tolist() == [0, 1, 0, 0, 1] and rs.quals.tolist() == [1, 2, 3, 4, 5]
