"""hiphase_b200 -- B200-native implementation of HiPhase's per-block phasing hot path.

  astar_phaser   mirror of src/astar_phaser.rs  (astar_solver, AstarResult) on top of the C ABI
  wfa_graph      mirror of src/wfa_graph.rs     (WFAGraph, WFAResult, WFAGraphError)
  read_segments  mirror of src/data_types/read_segments.rs (AlleleType, ReadSegment)
  variants       mirror of src/data_types/variants.rs allele matching + sequence_alignment::edit_distance
  read_parsing   mirror of read_parsing::local_realignment (AlignedRead, ReadStats)
  lib            ctypes loader of csrc/libhiphase_b200.so (hand-written sm_100a kernels behind include/hiphase_b200.h)
  synth          seeded synthetic workloads of BASELINE.json
"""
from . import _abi  # noqa: F401

__all__ = ["_abi"]
