"""Mirror of src/wfa_graph.rs's public surface on top of the CUDA path.

  WFAGraph(max_edit_distance).add_node(sequence, parent_nodes) -> index            (wfa_graph.rs:298)
  WFAGraph.edit_distance(other) / edit_distance_with_pruning(other, prune) -> WFAResult   (wfa_graph.rs:338, 350)
  WFAGraphError                                                                    (wfa_graph.rs:13-17)

Graph construction from variants (from_reference_variants_with_hom, wfa_graph.rs:119) is batched through
lib.Context.wfa_align_batch, which also returns the allele / quality rows of read_parsing.rs:790-835.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _abi as A
from . import lib


class WFAGraphError(Exception):
    """WFAGraphError::MaxEditDistance { distance }"""

    def __init__(self, distance):
        super().__init__("Max_edit_distance (%d) reached during WFA solving" % distance)
        self.distance = distance


@dataclass
class WFAResult:
    """wfa_graph.rs:655-670"""
    score: int
    traversed_nodes: List[int] = field(default_factory=list)


class WFAGraph:
    def __init__(self, max_edit_distance=1000, device=0):
        self._seqs, self._parents = [], []
        self.max_edit_distance = int(max_edit_distance)
        self._device = device
        self._ctx = None

    def get_num_nodes(self):
        return len(self._seqs)

    def add_node(self, sequence, parent_nodes):
        """wfa_graph.rs:298-331 (same three error conditions)."""
        idx = len(self._seqs)
        parent_nodes = sorted(int(p) for p in parent_nodes)
        if idx == 0 and parent_nodes:
            raise ValueError("First node must have no parent nodes.")
        if idx > 0 and not parent_nodes:
            raise ValueError("All nodes after the first must have at least one parent node.")
        if any(p >= idx for p in parent_nodes):
            raise ValueError("All parent nodes must come before this node.")
        self._seqs.append(np.asarray(list(sequence), dtype=np.uint8))
        self._parents.append(parent_nodes)
        return idx

    def edit_distance(self, other_sequence):
        return self.edit_distance_with_pruning(other_sequence, None)

    def edit_distance_with_pruning(self, other_sequence, prune_distance):
        if self._ctx is None:
            self._ctx = lib.Context(device=self._device)
        seq = np.concatenate(self._seqs) if self._seqs and sum(len(s) for s in self._seqs) else np.zeros(0, np.uint8)
        seq_off = np.concatenate([[0], np.cumsum([len(s) for s in self._seqs])])
        par = np.array([p for ps in self._parents for p in ps], np.uint32)
        par_off = np.concatenate([[0], np.cumsum([len(ps) for ps in self._parents])])
        st, score, nodes = self._ctx.wfa_graph_align(seq, seq_off, par, par_off, np.asarray(list(other_sequence), np.uint8),
                                                     prune_distance, self.max_edit_distance)
        if st == A.HP_WFA_MAX_EDIT_DISTANCE:
            raise WFAGraphError(self.max_edit_distance)
        return WFAResult(score, nodes)
