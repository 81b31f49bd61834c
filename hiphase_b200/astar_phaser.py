"""Mirror of src/astar_phaser.rs's public surface on top of the CUDA path.

  astar_solver(variants, read_segments, min_queue_size, queue_increment) -> AstarResult      (astar_phaser.rs:426)
  astar_solver_batch([...blocks...])                                                         (many blocks per launch)

`variants` only contributes what the reference reads from it on this path: is_ignored() and get_type()==Snv
(astar_phaser.rs:438, 446, 606).  All computation happens in hiphase_b200/csrc (no host fallback).
"""
from dataclasses import dataclass

import numpy as np

from . import _abi as A
from . import lib
from .read_segments import AlleleType, ReadSegment  # noqa: F401

_CTX = {}


def _context(min_queue_size, queue_increment, device):
    key = (int(min_queue_size), int(queue_increment), int(device))
    if key not in _CTX:
        _CTX[key] = lib.Context(A.hp_params(key[0], key[1], 500, 500), device=device)
    return _CTX[key]


@dataclass
class PhaseStats:
    """writers/phase_stats.rs:131-173 (the fields astar_solver fills, astar_phaser.rs:618-621)"""
    pruned_solutions: int
    estimated_cost: int
    actual_cost: int
    phased_variants: int
    phased_snvs: int
    homozygous_variants: int
    skipped_variants: int


@dataclass
class AstarResult:
    """astar_phaser.rs:408-415"""
    haplotype_1: np.ndarray
    haplotype_2: np.ndarray
    statistics: PhaseStats


def _block(variants, read_segments):
    ignored = np.array([bool(v.is_ignored()) if hasattr(v, "is_ignored") else bool(v["ignored"]) for v in variants], np.uint8)
    is_snv = np.array([(v.get_type() == 0) if hasattr(v, "get_type") else bool(v.get("is_snv", True)) for v in variants], np.uint8)
    return {"n_var": len(variants), "reads": [(rs.start, rs.alleles, rs.quals) for rs in read_segments],
            "ignored": ignored, "is_snv": is_snv}


def astar_solver_batch(blocks, min_queue_size=1000, queue_increment=3, device=0):
    """blocks: iterable of (variants, read_segments).  One launch for all of them."""
    batch = A.BlockBatch.from_blocks([_block(v, r) for v, r in blocks])
    out = _context(min_queue_size, queue_increment, device).astar_solve_batch(batch)
    results = []
    for b in range(batch.n_blocks):
        if out.status[b] != A.HP_BLOCK_OK:
            # the reference panics here (astar_phaser.rs:439, 529, 631)
            raise RuntimeError("astar_solver: block %d rejected with status %d" % (b, int(out.status[b])))
        v0, v1 = int(batch.var_off[b]), int(batch.var_off[b + 1])
        st = PhaseStats(*[int(out.stats[b][k]) for k in out.stats.dtype.names])
        results.append(AstarResult(out.h1[v0:v1].copy(), out.h2[v0:v1].copy(), st))
    return results


def astar_solver(variants, read_segments, min_queue_size=1000, queue_increment=3, device=0):
    """Call-site-1 shape (src/phaser.rs:541-543); phase_block is only used for log text in the reference."""
    return astar_solver_batch([(variants, read_segments)], min_queue_size, queue_increment, device)[0]
