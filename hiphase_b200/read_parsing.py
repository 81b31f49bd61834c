"""Mirror of the realignment entry points of src/read_parsing.rs on top of the CUDA path.

  local_realignment(read, variant_calls) -> (alleles, quals, ReadStats)        read_parsing.rs:121-503
  ReadStats (the fields local_realignment fills)                               writers/read_stats.rs, read_parsing.rs:457-503

`read` is anything with the four things the reference takes from a bam::Record: pos, aligned segments (the gap-free
M/=/X runs of rust_htslib's aligned_pairs()), the decoded sequence and the base qualities.  Batches go through
lib.Context.local_realign_batch; this module is the one-read convenience with the reference's argument order.
"""
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

from . import _abi as A
from . import lib
from .variants import variant_table

N_TYPES = 11      # VariantType::Unknown as usize + 1


@dataclass
class AlignedRead:
    """What local_realignment reads from a bam::Record (read_parsing.rs:137-156)."""
    pos: int
    segments: Sequence[Tuple[int, int, int]]      # (reference start, read start, length), ascending
    seq: bytes
    qual: Sequence[int]

    @staticmethod
    def from_cigar(pos, cigar, seq, qual):
        """cigar: [(op, len)] with op in 'M=XIDNSHP' -- the aligned_pairs() of rust_htslib, grouped into runs."""
        segs, r, q = [], int(pos), 0
        for op, n in cigar:
            if op in "M=X":
                segs.append((r, q, n)); r += n; q += n
            elif op in "IS":
                q += n
            elif op in "DN":
                r += n
        return AlignedRead(int(pos), segs, bytes(seq), list(qual))


@dataclass
class ReadStats:
    num_reads: int = 0
    skipped_reads: int = 0
    num_alleles: int = 0
    exact_matches: List[int] = field(default_factory=lambda: [0] * N_TYPES)
    inexact_matches: List[int] = field(default_factory=lambda: [0] * N_TYPES)
    failed_matches: List[int] = field(default_factory=lambda: [0] * N_TYPES)
    allele0_matches: List[int] = field(default_factory=lambda: [0] * N_TYPES)
    allele1_matches: List[int] = field(default_factory=lambda: [0] * N_TYPES)
    global_aligned: int = 0
    local_aligned: int = 0


def read_stats(alleles, match_class, vtypes):
    """The statistics block of local_realignment (read_parsing.rs:457-503) from one output row."""
    st = ReadStats()
    overlaps = 0
    for al, mc, vt in zip(alleles, match_class, vtypes):
        if not (mc & A.HP_LOCAL_OVERLAPS):
            continue
        if al == 2:
            st.failed_matches[vt] += 1
            continue
        (st.exact_matches if mc & A.HP_LOCAL_EXACT else st.inexact_matches)[vt] += 1
        (st.allele0_matches if al == 0 else st.allele1_matches)[vt] += 1
        overlaps += 1
        st.num_alleles += 1
    st.skipped_reads = 1 if overlaps == 0 else 0
    st.local_aligned = 1 - st.skipped_reads
    return st


_CTX = None


def local_realignment(read: AlignedRead, variant_calls, ctx=None):
    """-> (alleles, quals, ReadStats); raises like the reference panics on an unhandled variant type."""
    global _CTX
    if ctx is None:
        if _CTX is None:
            _CTX = lib.Context(device=0)
        ctx = _CTX
    n = len(variant_calls)
    segs = list(read.segments)
    batch = A.LocalBatch(variant_table(variant_calls), [0], [n], [read.pos], [0, len(segs)], [s[0] for s in segs], [s[1] for s in segs],
                         [s[2] for s in segs], np.frombuffer(bytes(read.seq), np.uint8) if len(read.seq) else np.zeros(1, np.uint8),
                         np.asarray(read.qual, np.uint8) if len(read.qual) else np.zeros(1, np.uint8), [0, len(read.seq)])
    out = ctx.local_realign_batch(batch)
    if out.status[0] == A.HP_LOCAL_UNHANDLED_TYPE:
        raise RuntimeError("Unhandled variant type")                  # panic!, read_parsing.rs:452-454
    if out.status[0] != A.HP_LOCAL_OK:
        raise RuntimeError("local realignment failed with job status %d" % out.status[0])
    vt = [int(v.get_type()) for v in variant_calls]
    return out.alleles.tolist(), out.quals.tolist(), read_stats(out.alleles, out.match_class, vt)


def plan_global_realignment(read: AlignedRead, variant_positions, hom_positions):
    """The CIGAR-projection half of global_realignment (read_parsing.rs:672-742): what one job of hp_wfa_align_batch needs.

    variant_positions / hom_positions: Variant::position() of the block's het / hom calls, ascending.
    Returns None when the mapping overlaps no het variant (the short-circuit at :703-712), else a dict with
      ref_start, ref_end        min_position, max_position + 1 of the aligned pairs           (:677-689, :773-774)
      het_lo, het_hi            first_overlap, last_overlap (indices into variant_positions)  (:692-701)
      hom_lo, hom_hi            first_hom_overlap, last_hom_overlap                           (:718-729)
      read_start, read_end      slice of the read that is aligned: read[read_start:read_end]  (:737-741)
    """
    segs = list(read.segments)
    if not segs:
        raise AssertionError("max_position >= min_position")              # the reference asserts at :686
    min_position = min(s[0] for s in segs)
    max_position = max(s[0] + s[2] - 1 for s in segs)

    def overlap(positions):
        idx = [i for i, p in enumerate(positions) if min_position <= p <= max_position]
        return (idx[0], idx[-1] + 1) if idx else None

    het = overlap(variant_positions)
    if het is None:
        return None
    hom = overlap(hom_positions) or (0, 0)
    first = min(segs, key=lambda s: s[0])
    last = max(segs, key=lambda s: s[0] + s[2] - 1)
    return dict(ref_start=min_position, ref_end=max_position + 1, het_lo=het[0], het_hi=het[1], hom_lo=hom[0], hom_hi=hom[1],
                read_start=first[1], read_end=last[1] + last[2])
