"""Host-side mirror of src/data_types/read_segments.rs (AlleleType, ReadSegment) -- plain data handling only.

The scoring methods of the reference (score_partial_haplotype, read_segments.rs:177-206) are NOT re-implemented on
the host: scoring happens inside the CUDA kernels.  This class only carries the clipped allele / quality vectors the
C ABI consumes, with the reference's constructor semantics (clip to first/last set allele, read_segments.rs:40-62).
"""
import enum

import numpy as np


class AlleleType(enum.IntEnum):
    """src/data_types/read_segments.rs:5-16"""
    Reference = 0
    Alternate = 1
    Ambiguous = 2
    NoOverlap = 3


class ReadSegment:
    """src/data_types/read_segments.rs:19-62: alleles / quals clipped to region = [first set, last set + 1)."""

    def __init__(self, read_name, alleles, quals):
        alleles = np.asarray(alleles, dtype=np.uint8)
        quals = np.asarray(quals, dtype=np.uint8)
        if len(alleles) != len(quals):
            raise ValueError("alleles and quals must have the same length (read_segments.rs:41)")
        isset = np.flatnonzero(alleles < AlleleType.Ambiguous)
        if len(isset):
            first, last = int(isset[0]), int(isset[-1]) + 1
        else:
            first = last = len(alleles)
        self.read_name = read_name
        self.alleles = alleles[first:last].copy()
        self.quals = quals[first:last].copy()
        self.start, self.end = first, last

    def region(self):
        return range(self.start, self.end)

    def get_num_set(self):
        """read_segments.rs:151-155"""
        return int((self.alleles < AlleleType.Ambiguous).sum())

    def __eq__(self, o):
        return (self.read_name, self.start, self.end) == (o.read_name, o.start, o.end) and \
            np.array_equal(self.alleles, o.alleles) and np.array_equal(self.quals, o.quals)
