"""Mirror of the parts of src/data_types/variants.rs and src/sequence_alignment.rs that sit on the local-realignment
path, on top of the CUDA path (there is no CPU implementation here: distances come from hp_edit_distance_batch).

  VariantType                                                   variants.rs:8-31
  Variant.new_snv / new_deletion / ... (data carrier only)      variants.rs:96-494
  Variant.add_reference_prefix / add_reference_postfix / truncate_reference_postfix     variants.rs:497-539
  Variant.match_allele / closest_allele / closest_allele_clip   variants.rs:598-641
  edit_distance(v1, v2)                                         sequence_alignment.rs:6-38
"""
import enum

import numpy as np

from . import lib

_CTX = None


def _ctx():
    global _CTX
    if _CTX is None:
        _CTX = lib.Context(device=0)
    return _CTX


class VariantType(enum.IntEnum):
    Snv = 0
    Insertion = 1
    Deletion = 2
    Indel = 3
    SvInsertion = 4
    SvDeletion = 5
    SvDuplication = 6
    SvInversion = 7
    SvBreakend = 8
    TandemRepeat = 9
    Unknown = 10


def edit_distance(v1, v2):
    """sequence_alignment::edit_distance on the GPU."""
    return int(_ctx().edit_distance_batch([(v1, v2)])[0])


class Variant:
    """Data carrier with the reference's accessor names; allele matching runs through the C ABI."""

    def __init__(self, vcf_index, variant_type, position, ref_len, allele0, allele1, index_allele0=0, index_allele1=1):
        if index_allele0 >= index_allele1:
            raise ValueError("index_allele0 must be less than index_allele1")          # VariantError::IndexAlleleOrder
        self.vcf_index = vcf_index
        self.variant_type = VariantType(variant_type)
        self._position = int(position)
        self.ref_len = int(ref_len)
        self.prefix_len = 0
        self.postfix_len = 0
        self.allele0 = bytes(allele0)
        self.allele1 = bytes(allele1)
        self.index_allele0 = index_allele0
        self.index_allele1 = index_allele1
        self._ignored = False

    # accessors (variants.rs:541-591)
    def get_type(self): return self.variant_type
    def position(self): return self._position
    def get_ref_len(self): return self.ref_len
    def get_prefix_len(self): return self.prefix_len
    def get_postfix_len(self): return self.postfix_len
    def get_allele0(self): return self.allele0
    def get_allele1(self): return self.allele1
    def is_ignored(self): return self._ignored
    def set_ignored(self): self._ignored = True
    def get_truncated_allele0(self): return self.allele0[self.prefix_len: len(self.allele0) - self.postfix_len]
    def get_truncated_allele1(self): return self.allele1[self.prefix_len: len(self.allele1) - self.postfix_len]

    def add_reference_prefix(self, prefix):
        prefix = bytes(prefix)
        assert len(prefix) <= self._position - self.prefix_len
        self.allele0 = prefix + self.allele0
        self.allele1 = prefix + self.allele1
        self.prefix_len += len(prefix)

    def add_reference_postfix(self, postfix):
        postfix = bytes(postfix)
        self.allele0 += postfix
        self.allele1 += postfix
        self.postfix_len += len(postfix)

    def truncate_reference_postfix(self, truncate_amount):
        assert truncate_amount <= self.postfix_len
        self.allele0 = self.allele0[: len(self.allele0) - truncate_amount]
        self.allele1 = self.allele1[: len(self.allele1) - truncate_amount]
        self.postfix_len -= truncate_amount

    def match_allele(self, allele):
        allele = bytes(allele)
        return 0 if allele == self.allele0 else (1 if allele == self.allele1 else 2)

    def closest_allele(self, allele):
        return self.closest_allele_clip(allele, 0, 0)

    def closest_allele_clip(self, allele, head_clip, tail_clip):
        """-> (allele chosen 0/1/2, min edit distance, other edit distance)"""
        assert head_clip <= self.prefix_len and tail_clip <= self.postfix_len
        a0 = self.allele0[head_clip: len(self.allele0) - tail_clip]
        a1 = self.allele1[head_clip: len(self.allele1) - tail_clip]
        d0, d1 = (int(x) for x in _ctx().edit_distance_batch([(allele, a0), (allele, a1)]))
        if d0 < d1:
            return 0, d0, d1
        if d0 > d1:
            return 1, d1, d0
        return 2, d0, d1

    def convert_index(self, index):
        if index == 0:
            return self.index_allele0
        if index == 1:
            return self.index_allele1
        if index == 2:
            return 255
        raise ValueError("index must be 0, 1, or 2")


def variant_table(variants):
    """Pack Variant objects into the arrays of hp_variant_table (+ prefix_len / postfix_len) with FULL alleles."""
    keys = ("position", "ref_len", "allele0_off", "allele0_len", "allele1_off", "allele1_len", "index_allele0", "vtype", "ignored",
            "prefix_len", "postfix_len")
    t = {k: [] for k in keys}
    blob, n = [], 0
    for v in variants:
        t["position"].append(v.position()); t["ref_len"].append(v.ref_len)
        t["allele0_off"].append(n); t["allele0_len"].append(len(v.allele0)); blob.append(np.frombuffer(v.allele0, np.uint8)); n += len(v.allele0)
        t["allele1_off"].append(n); t["allele1_len"].append(len(v.allele1)); blob.append(np.frombuffer(v.allele1, np.uint8)); n += len(v.allele1)
        t["index_allele0"].append(v.index_allele0); t["vtype"].append(int(v.variant_type)); t["ignored"].append(1 if v.is_ignored() else 0)
        t["prefix_len"].append(v.prefix_len); t["postfix_len"].append(v.postfix_len)
    t["allele_bytes"] = np.concatenate(blob) if n else np.zeros(1, np.uint8)
    return t
