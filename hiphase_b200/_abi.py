"""ctypes mirror of include/hiphase_b200.h (the C ABI structs) plus numpy <-> struct packing helpers.

The same struct layouts feed the product library (hiphase_b200/csrc/libhiphase_b200.so) and, in tests and the
bench CPU-baseline leg only, the CPU checker under oracle/.
"""
import ctypes as C

import numpy as np

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)

HP_OK = 0
HP_ERR_INVALID_INPUT = -1
HP_ERR_NO_DEVICE = -2
HP_ERR_CUDA = -3
HP_ERR_UNSUPPORTED = -4
HP_ERR_OUT_OF_MEMORY = -5
HP_ERR_INTERNAL = -6

HP_BLOCK_OK = 0
HP_BLOCK_IGNORED_NOT_NOOVERLAP = 1
HP_BLOCK_COST_OVERFLOW = 2
HP_BLOCK_QUEUE_OVERFLOW = 3
HP_BLOCK_ASSERT = 4
HP_BLOCK_TOO_DENSE = 5
HP_BLOCK_INDEX_EXHAUSTED = 6
HP_COMM_ID_BYTES = 128

HP_WFA_OK = 0
HP_WFA_MAX_EDIT_DISTANCE = 1
HP_WFA_SKIPPED = 2
HP_WFA_WORKSPACE_OVERFLOW = 3
HP_WFA_GRAPH_TOO_LARGE = 4

# VariantType (src/data_types/variants.rs:8-31)
VT_SNV, VT_INSERTION, VT_DELETION, VT_INDEL, VT_SV_INSERTION, VT_SV_DELETION = 0, 1, 2, 3, 4, 5
VT_SV_DUPLICATION, VT_SV_INVERSION, VT_SV_BREAKEND, VT_TANDEM_REPEAT, VT_UNKNOWN = 6, 7, 8, 9, 10


class hp_params(C.Structure):
    _fields_ = [("min_queue_size", C.c_uint32), ("queue_increment", C.c_uint32),
                ("wfa_prune_distance", C.c_uint32), ("wfa_max_edit_distance", C.c_uint32)]


def default_params():
    """Defaults of src/cli.rs:186-226."""
    return hp_params(1000, 3, 500, 500)


class hp_block_batch(C.Structure):
    _fields_ = [("n_blocks", C.c_uint32), ("var_off", u64p), ("read_off", u64p), ("read_start", u32p),
                ("read_end", u32p), ("cell_off", u64p), ("alleles", u8p), ("quals", u8p),
                ("ignored", u8p), ("is_snv", u8p)]


class hp_phase_stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("pruned_solutions", "estimated_cost", "actual_cost", "phased_variants",
                                          "phased_snvs", "homozygous_variants", "skipped_variants")]


class hp_astar_counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("evals", "cells", "sum_parent_len", "pops")]


class hp_astar_out(C.Structure):
    _fields_ = [("h1", u8p), ("h2", u8p), ("stats", C.POINTER(hp_phase_stats)), ("status", i32p),
                ("heuristic", u64p), ("counters", C.POINTER(hp_astar_counters))]


class hp_variant_table(C.Structure):
    _fields_ = [("n_variants", C.c_uint32), ("position", i64p), ("ref_len", u32p), ("allele0_off", u64p),
                ("allele0_len", u32p), ("allele1_off", u64p), ("allele1_len", u32p), ("index_allele0", u8p),
                ("vtype", u8p), ("ignored", u8p), ("allele_bytes", u8p), ("n_allele_bytes", C.c_uint64)]


class hp_wfa_batch(C.Structure):
    _fields_ = [("n_jobs", C.c_uint32), ("variants", hp_variant_table), ("reference", u8p),
                ("n_reference", C.c_uint64), ("ref_start", u64p), ("ref_end", u64p), ("het_lo", u32p),
                ("het_hi", u32p), ("hom_lo", u32p), ("hom_hi", u32p), ("read_bytes", u8p), ("read_off", u64p),
                ("row_off", u64p)]


class hp_wfa_counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("bases_compared", "waves_processed", "set_ops", "n_nodes")]


class hp_wfa_out(C.Structure):
    _fields_ = [("status", i32p), ("score", u32p), ("alleles", u8p), ("quals", u8p), ("n_nodes", u32p),
                ("traversed", u64p), ("trav_words", C.c_uint32), ("counters", C.POINTER(hp_wfa_counters))]


class hp_post_out(C.Structure):
    _fields_ = [("span_counts", u32p), ("block_tags", u64p), ("read_haplotag", u8p), ("read_tag", u64p)]


STATS_DTYPE = np.dtype([(n, "<u8") for n, _ in hp_phase_stats._fields_])
COUNTERS_DTYPE = np.dtype([(n, "<u8") for n, _ in hp_astar_counters._fields_])
WFA_COUNTERS_DTYPE = np.dtype([(n, "<u8") for n, _ in hp_wfa_counters._fields_])


def ptr(arr, ctype_ptr):
    """Pointer to a C-contiguous numpy array (or NULL for None)."""
    if arr is None:
        return ctype_ptr()
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(ctype_ptr)


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class BlockBatch:
    """A batch of phase blocks in the reference's u8 layout (numpy side of hp_block_batch)."""

    FIELDS = ("var_off", "read_off", "read_start", "read_end", "cell_off", "alleles", "quals", "ignored", "is_snv")

    def __init__(self, var_off, read_off, read_start, read_end, cell_off, alleles, quals, ignored, is_snv):
        self.var_off = _np(var_off, np.uint64)
        self.read_off = _np(read_off, np.uint64)
        self.read_start = _np(read_start, np.uint32)
        self.read_end = _np(read_end, np.uint32)
        self.cell_off = _np(cell_off, np.uint64)
        self.alleles = _np(alleles, np.uint8)
        self.quals = _np(quals, np.uint8)
        self.ignored = _np(ignored, np.uint8)
        self.is_snv = _np(is_snv, np.uint8)
        self.n_blocks = len(self.var_off) - 1
        assert len(self.read_off) == self.n_blocks + 1
        assert len(self.cell_off) == len(self.read_start) + 1 == len(self.read_end) + 1

    @property
    def n_vars(self):
        return int(self.var_off[-1])

    @property
    def n_reads(self):
        return int(self.read_off[-1])

    @property
    def n_cells(self):
        return int(self.cell_off[-1])

    def as_struct(self):
        return hp_block_batch(self.n_blocks, ptr(self.var_off, u64p), ptr(self.read_off, u64p),
                              ptr(self.read_start, u32p), ptr(self.read_end, u32p), ptr(self.cell_off, u64p),
                              ptr(self.alleles, u8p), ptr(self.quals, u8p), ptr(self.ignored, u8p),
                              ptr(self.is_snv, u8p))

    @staticmethod
    def from_blocks(blocks):
        """blocks: list of dicts {n_var, reads:[(start, alleles_u8, quals_u8)], ignored, is_snv} (clipped reads)."""
        var_off, read_off, rs, re, cell_off = [0], [0], [], [], [0]
        al, ql, ig, sn = [], [], [], []
        for b in blocks:
            n = int(b["n_var"])
            var_off.append(var_off[-1] + n)
            ig.append(_np(b.get("ignored", np.zeros(n)), np.uint8))
            sn.append(_np(b.get("is_snv", np.ones(n)), np.uint8))
            for (start, a, q) in b["reads"]:
                a = _np(a, np.uint8)
                q = _np(q, np.uint8)
                assert len(a) == len(q)
                rs.append(int(start))
                re.append(int(start) + len(a))
                cell_off.append(cell_off[-1] + len(a))
                al.append(a)
                ql.append(q)
            read_off.append(len(rs))
        cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.uint8)
        return BlockBatch(var_off, read_off, rs, re, cell_off, cat(al), cat(ql), cat(ig), cat(sn))

    def block(self, b):
        """Block b back as a dict (for the slow python restatement and brute-force checks)."""
        v0, v1 = int(self.var_off[b]), int(self.var_off[b + 1])
        reads = []
        for r in range(int(self.read_off[b]), int(self.read_off[b + 1])):
            c0, c1 = int(self.cell_off[r]), int(self.cell_off[r + 1])
            reads.append((int(self.read_start[r]), self.alleles[c0:c1].copy(), self.quals[c0:c1].copy()))
        return {"n_var": v1 - v0, "reads": reads, "ignored": self.ignored[v0:v1].copy(), "is_snv": self.is_snv[v0:v1].copy()}

    def select(self, idx):
        return BlockBatch.from_blocks([self.block(int(i)) for i in idx])


class AstarOut:
    """Host-side result buffers of hp_astar_solve_batch."""

    def __init__(self, batch, want_heuristic=False, want_counters=False):
        self.h1 = np.full(batch.n_vars, 255, np.uint8)
        self.h2 = np.full(batch.n_vars, 255, np.uint8)
        self.stats = np.zeros(batch.n_blocks, STATS_DTYPE)
        self.status = np.full(batch.n_blocks, -1, np.int32)
        self.heuristic = np.zeros(batch.n_vars + batch.n_blocks, np.uint64) if want_heuristic else None
        self.counters = np.zeros(batch.n_blocks, COUNTERS_DTYPE) if want_counters else None

    @staticmethod
    def sized(n_vars, n_blocks, alloc=None):
        """Result buffers for n_vars variants / n_blocks blocks; alloc(n, dtype) -> array (e.g. pinned memory)."""
        o = AstarOut.__new__(AstarOut)
        mk = alloc or (lambda n, dt: np.zeros(n, dt))
        o.h1 = mk(n_vars, np.uint8); o.h2 = mk(n_vars, np.uint8)
        o.h1[...] = 255; o.h2[...] = 255
        o.stats = mk(n_blocks, STATS_DTYPE)
        o.status = mk(n_blocks, np.int32); o.status[...] = -1
        o.heuristic = None; o.counters = None
        return o

    def as_struct(self):
        return hp_astar_out(ptr(self.h1, u8p), ptr(self.h2, u8p),
                            self.stats.ctypes.data_as(C.POINTER(hp_phase_stats)), ptr(self.status, i32p),
                            ptr(self.heuristic, u64p),
                            self.counters.ctypes.data_as(C.POINTER(hp_astar_counters)) if self.counters is not None
                            else C.POINTER(hp_astar_counters)())


class WfaBatch:
    """numpy side of hp_wfa_batch.  variants: dict of arrays (see hp_variant_table)."""

    def __init__(self, variants, reference, ref_start, ref_end, het_lo, het_hi, hom_lo, hom_hi, read_bytes, read_off):
        v = variants
        self.position = _np(v["position"], np.int64)
        self.ref_len = _np(v["ref_len"], np.uint32)
        self.allele0_off = _np(v["allele0_off"], np.uint64)
        self.allele0_len = _np(v["allele0_len"], np.uint32)
        self.allele1_off = _np(v["allele1_off"], np.uint64)
        self.allele1_len = _np(v["allele1_len"], np.uint32)
        self.index_allele0 = _np(v["index_allele0"], np.uint8)
        self.vtype = _np(v["vtype"], np.uint8)
        self.ignored = _np(v["ignored"], np.uint8)
        self.allele_bytes = _np(v["allele_bytes"], np.uint8)
        self.reference = _np(reference, np.uint8)
        self.ref_start = _np(ref_start, np.uint64)
        self.ref_end = _np(ref_end, np.uint64)
        self.het_lo = _np(het_lo, np.uint32)
        self.het_hi = _np(het_hi, np.uint32)
        self.hom_lo = _np(hom_lo, np.uint32)
        self.hom_hi = _np(hom_hi, np.uint32)
        self.read_bytes = _np(read_bytes, np.uint8)
        self.read_off = _np(read_off, np.uint64)
        self.n_jobs = len(self.ref_start)
        row_len = (self.het_hi.astype(np.int64) - self.het_lo.astype(np.int64))
        self.row_off = np.concatenate([[0], np.cumsum(row_len)]).astype(np.uint64)

    @property
    def n_variants(self):
        return len(self.position)

    def as_struct(self):
        vt = hp_variant_table(self.n_variants, ptr(self.position, i64p), ptr(self.ref_len, u32p),
                              ptr(self.allele0_off, u64p), ptr(self.allele0_len, u32p), ptr(self.allele1_off, u64p),
                              ptr(self.allele1_len, u32p), ptr(self.index_allele0, u8p), ptr(self.vtype, u8p),
                              ptr(self.ignored, u8p), ptr(self.allele_bytes, u8p), len(self.allele_bytes))
        return hp_wfa_batch(self.n_jobs, vt, ptr(self.reference, u8p), len(self.reference), ptr(self.ref_start, u64p),
                            ptr(self.ref_end, u64p), ptr(self.het_lo, u32p), ptr(self.het_hi, u32p),
                            ptr(self.hom_lo, u32p), ptr(self.hom_hi, u32p), ptr(self.read_bytes, u8p),
                            ptr(self.read_off, u64p), ptr(self.row_off, u64p))


class WfaOut:
    def __init__(self, batch, trav_words=0, want_counters=False):
        n = batch.n_jobs
        self.status = np.full(n, -1, np.int32)
        self.score = np.zeros(n, np.uint32)
        self.alleles = np.full(int(batch.row_off[-1]), 255, np.uint8)
        self.quals = np.full(int(batch.row_off[-1]), 255, np.uint8)
        self.n_nodes = np.zeros(n, np.uint32)
        self.trav_words = int(trav_words)
        self.traversed = np.zeros(n * trav_words, np.uint64) if trav_words else None
        self.counters = np.zeros(n, WFA_COUNTERS_DTYPE) if want_counters else None

    def as_struct(self):
        return hp_wfa_out(ptr(self.status, i32p), ptr(self.score, u32p), ptr(self.alleles, u8p), ptr(self.quals, u8p),
                          ptr(self.n_nodes, u32p), ptr(self.traversed, u64p), self.trav_words,
                          self.counters.ctypes.data_as(C.POINTER(hp_wfa_counters)) if self.counters is not None
                          else C.POINTER(hp_wfa_counters)())


class PostOut:
    """Host-side result buffers of hp_post_solve_batch (span counts, block tags, read haplotags)."""

    def __init__(self, batch):
        self.span_counts = np.zeros(batch.n_vars, np.uint32)
        self.block_tags = np.zeros(batch.n_vars, np.uint64)
        self.read_haplotag = np.full(batch.n_reads, 255, np.uint8)
        self.read_tag = np.zeros(batch.n_reads, np.uint64)

    def as_struct(self):
        return hp_post_out(ptr(self.span_counts, u32p), ptr(self.block_tags, u64p), ptr(self.read_haplotag, u8p), ptr(self.read_tag, u64p))


# ---- local realignment (src/read_parsing.rs:121-503) ------------------------------------------
HP_LOCAL_OK, HP_LOCAL_UNHANDLED_TYPE, HP_LOCAL_BAD_SLICE, HP_LOCAL_ALLELE_TOO_LONG = 0, 1, 2, 3
HP_LOCAL_OVERLAPS, HP_LOCAL_EXACT = 1, 2


class hp_local_batch(C.Structure):
    _fields_ = [("n_jobs", C.c_uint32), ("variants", hp_variant_table), ("prefix_len", u32p), ("postfix_len", u32p),
                ("var_lo", u32p), ("var_hi", u32p), ("read_pos", i64p), ("seg_off", u64p), ("seg_ref_start", i64p),
                ("seg_read_start", u32p), ("seg_len", u32p), ("read_bytes", u8p), ("read_quals", u8p),
                ("read_off", u64p), ("row_off", u64p)]


class hp_local_out(C.Structure):
    _fields_ = [("alleles", u8p), ("quals", u8p), ("match_class", u8p), ("edit_distance", u32p), ("status", i32p)]


class LocalBatch:
    """numpy side of hp_local_batch.  variants: dict of arrays as for WfaBatch (FULL alleles) + prefix_len/postfix_len;
    jobs: per read mapping var_lo/var_hi, read_pos, aligned segments (seg_off CSR), read bytes and base qualities."""

    def __init__(self, variants, var_lo, var_hi, read_pos, seg_off, seg_ref_start, seg_read_start, seg_len,
                 read_bytes, read_quals, read_off):
        v = variants
        self.position = _np(v["position"], np.int64)
        self.ref_len = _np(v["ref_len"], np.uint32)
        self.allele0_off = _np(v["allele0_off"], np.uint64)
        self.allele0_len = _np(v["allele0_len"], np.uint32)
        self.allele1_off = _np(v["allele1_off"], np.uint64)
        self.allele1_len = _np(v["allele1_len"], np.uint32)
        self.index_allele0 = _np(v.get("index_allele0", np.zeros(len(self.position))), np.uint8)
        self.vtype = _np(v["vtype"], np.uint8)
        self.ignored = _np(v["ignored"], np.uint8)
        self.allele_bytes = _np(v["allele_bytes"], np.uint8)
        self.prefix_len = _np(v["prefix_len"], np.uint32)
        self.postfix_len = _np(v["postfix_len"], np.uint32)
        self.var_lo = _np(var_lo, np.uint32)
        self.var_hi = _np(var_hi, np.uint32)
        self.read_pos = _np(read_pos, np.int64)
        self.seg_off = _np(seg_off, np.uint64)
        self.seg_ref_start = _np(seg_ref_start, np.int64)
        self.seg_read_start = _np(seg_read_start, np.uint32)
        self.seg_len = _np(seg_len, np.uint32)
        self.read_bytes = _np(read_bytes, np.uint8)
        self.read_quals = _np(read_quals, np.uint8)
        self.read_off = _np(read_off, np.uint64)
        self.n_jobs = len(self.var_lo)
        row_len = self.var_hi.astype(np.int64) - self.var_lo.astype(np.int64)
        self.row_off = np.concatenate([[0], np.cumsum(row_len)]).astype(np.uint64)

    @property
    def n_variants(self):
        return len(self.position)

    def as_struct(self):
        vt = hp_variant_table(self.n_variants, ptr(self.position, i64p), ptr(self.ref_len, u32p),
                              ptr(self.allele0_off, u64p), ptr(self.allele0_len, u32p), ptr(self.allele1_off, u64p),
                              ptr(self.allele1_len, u32p), ptr(self.index_allele0, u8p), ptr(self.vtype, u8p),
                              ptr(self.ignored, u8p), ptr(self.allele_bytes, u8p), len(self.allele_bytes))
        return hp_local_batch(self.n_jobs, vt, ptr(self.prefix_len, u32p), ptr(self.postfix_len, u32p),
                              ptr(self.var_lo, u32p), ptr(self.var_hi, u32p), ptr(self.read_pos, i64p),
                              ptr(self.seg_off, u64p), ptr(self.seg_ref_start, i64p), ptr(self.seg_read_start, u32p),
                              ptr(self.seg_len, u32p), ptr(self.read_bytes, u8p), ptr(self.read_quals, u8p),
                              ptr(self.read_off, u64p), ptr(self.row_off, u64p))


class LocalOut:
    def __init__(self, batch):
        n = int(batch.row_off[-1])
        self.alleles = np.full(n, 255, np.uint8)
        self.quals = np.full(n, 255, np.uint8)
        self.match_class = np.full(n, 255, np.uint8)
        self.edit_distance = np.zeros(2 * n, np.uint32)
        self.status = np.full(batch.n_jobs, -1, np.int32)

    def as_struct(self):
        return hp_local_out(ptr(self.alleles, u8p), ptr(self.quals, u8p), ptr(self.match_class, u8p),
                            ptr(self.edit_distance, u32p), ptr(self.status, i32p))


# ---- matrix assembly (read_segments.rs:40-121; read_parsing.rs:612-629) -------------------------------------------
HP_GROUP_DROPPED, HP_GROUP_PHASABLE, HP_GROUP_KEPT, HP_GROUP_ASSERT = 0, 1, 2, 3


class hp_rows_batch(C.Structure):
    _fields_ = [("n_blocks", C.c_uint32), ("var_off", u64p), ("group_off", u64p), ("group_row_off", u64p), ("row_start", u32p),
                ("row_cell_off", u64p), ("alleles", u8p), ("quals", u8p), ("min_matched_alleles", C.c_uint32)]


class hp_assembled(C.Structure):
    _fields_ = [("read_off", u64p), ("read_start", u32p), ("read_end", u32p), ("cell_off", u64p), ("alleles", u8p), ("quals", u8p),
                ("cell_capacity", C.c_uint64), ("group_class", u8p), ("group_num_set", u32p), ("n_reads", C.c_uint64),
                ("n_cells", C.c_uint64)]


class RowsBatch:
    """numpy side of hp_rows_batch.  blocks: list of dict(n_var, groups=[[(row_start, alleles, quals), ...], ...])."""

    def __init__(self, blocks, min_matched_alleles=2):
        var_off, group_off, group_row_off, row_start, row_cell_off, al, ql = [0], [0], [0], [], [0], [], []
        for b in blocks:
            var_off.append(var_off[-1] + int(b["n_var"]))
            for grp in b["groups"]:
                for (s, a, q) in grp:
                    a = np.asarray(a, np.uint8); q = np.asarray(q, np.uint8)
                    assert len(a) == len(q)
                    row_start.append(int(s)); al.append(a); ql.append(q)
                    row_cell_off.append(row_cell_off[-1] + len(a))
                group_row_off.append(len(row_start))
            group_off.append(len(group_row_off) - 1)
        self.n_blocks = len(blocks)
        self.var_off = _np(var_off, np.uint64); self.group_off = _np(group_off, np.uint64)
        self.group_row_off = _np(group_row_off, np.uint64); self.row_start = _np(row_start, np.uint32)
        self.row_cell_off = _np(row_cell_off, np.uint64)
        self.alleles = _np(np.concatenate(al) if al else np.zeros(0, np.uint8), np.uint8)
        self.quals = _np(np.concatenate(ql) if ql else np.zeros(0, np.uint8), np.uint8)
        self.min_matched_alleles = int(min_matched_alleles)
        self.n_groups = len(self.group_row_off) - 1
        # capacity: the span every group can touch
        cap = 0
        for g in range(self.n_groups):
            r0, r1 = int(self.group_row_off[g]), int(self.group_row_off[g + 1])
            lens = np.diff(self.row_cell_off[r0:r1 + 1].astype(np.int64))
            if r1 > r0 and lens.sum():
                st = self.row_start[r0:r1].astype(np.int64)
                cap += int((st + lens)[lens > 0].max() - st[lens > 0].min())
        self.cell_capacity = cap

    def as_struct(self):
        return hp_rows_batch(self.n_blocks, ptr(self.var_off, u64p), ptr(self.group_off, u64p), ptr(self.group_row_off, u64p),
                             ptr(self.row_start, u32p), ptr(self.row_cell_off, u64p), ptr(self.alleles, u8p), ptr(self.quals, u8p),
                             self.min_matched_alleles)


class Assembled:
    def __init__(self, rows):
        n = max(rows.n_groups, 1)
        self.read_off = np.zeros(rows.n_blocks + 1, np.uint64)
        self.read_start = np.zeros(n, np.uint32); self.read_end = np.zeros(n, np.uint32)
        self.cell_off = np.zeros(n + 1, np.uint64)
        self.cell_capacity = max(rows.cell_capacity, 1)
        self.alleles = np.full(self.cell_capacity, 255, np.uint8); self.quals = np.full(self.cell_capacity, 255, np.uint8)
        self.group_class = np.full(n, 255, np.uint8); self.group_num_set = np.zeros(n, np.uint32)
        self._s = hp_assembled(ptr(self.read_off, u64p), ptr(self.read_start, u32p), ptr(self.read_end, u32p), ptr(self.cell_off, u64p),
                               ptr(self.alleles, u8p), ptr(self.quals, u8p), self.cell_capacity, ptr(self.group_class, u8p),
                               ptr(self.group_num_set, u32p), 0, 0)

    def as_struct(self):
        return self._s

    def block_batch(self, rows, ignored=None, is_snv=None):
        """The assembled reads as a BlockBatch ready for astar_solve_batch."""
        nr, nc, nv = int(self._s.n_reads), int(self._s.n_cells), int(rows.var_off[-1])
        return BlockBatch(rows.var_off, self.read_off, self.read_start[:nr], self.read_end[:nr], self.cell_off[:nr + 1],
                          self.alleles[:nc], self.quals[:nc], np.zeros(nv, np.uint8) if ignored is None else ignored,
                          np.ones(nv, np.uint8) if is_snv is None else is_snv)


# ---- realignment pipeline (read_parsing.rs:545-629) ------------------------------------------------------------------
HP_MAP_GLOBAL, HP_MAP_LOCAL_FAILED, HP_MAP_LOCAL_DISABLED, HP_MAP_SKIPPED = 0, 1, 2, 3


class hp_realign_batch(C.Structure):
    _fields_ = [("n_blocks", C.c_uint32), ("map_off", u64p), ("map_group", u32p), ("n_groups", u32p), ("var_off", u64p),
                ("wfa_het_base", u32p), ("wfa", hp_wfa_batch), ("local", hp_local_batch), ("global_failure_minimum", C.c_uint32),
                ("global_failure_ratio", C.c_double), ("min_matched_alleles", C.c_uint32)]


class hp_realign_out(C.Structure):
    _fields_ = [("map_mode", u8p), ("map_score", u32p), ("block_disabled_at", u32p), ("block_failures", u32p),
                ("block_parsed", u32p), ("assembled", hp_assembled)]


class RealignBatch:
    """numpy side of hp_realign_batch: one WfaBatch job and one LocalBatch job per mapping, mappings grouped by block in BAM
    order, read-name groups per block."""

    def __init__(self, wfa, local, map_off, map_group, n_groups, var_off, wfa_het_base, global_failure_minimum=50,
                 global_failure_ratio=0.5, min_matched_alleles=2):
        self.wfa, self.local = wfa, local
        self.map_off = _np(map_off, np.uint64); self.map_group = _np(map_group, np.uint32)
        self.n_groups = _np(n_groups, np.uint32); self.var_off = _np(var_off, np.uint64)
        self.wfa_het_base = _np(wfa_het_base, np.uint32)
        self.global_failure_minimum = int(global_failure_minimum)
        self.global_failure_ratio = float(global_failure_ratio)
        self.min_matched_alleles = int(min_matched_alleles)
        self.n_blocks = len(self.n_groups)
        self.n_maps = int(self.map_off[-1])

    def as_struct(self):
        return hp_realign_batch(self.n_blocks, ptr(self.map_off, u64p), ptr(self.map_group, u32p), ptr(self.n_groups, u32p),
                                ptr(self.var_off, u64p), ptr(self.wfa_het_base, u32p), self.wfa.as_struct(), self.local.as_struct(),
                                self.global_failure_minimum, self.global_failure_ratio, self.min_matched_alleles)


class RealignOut:
    def __init__(self, batch):
        nm, nb = max(batch.n_maps, 1), batch.n_blocks
        self.map_mode = np.full(nm, 255, np.uint8); self.map_score = np.zeros(nm, np.uint32)
        self.block_disabled_at = np.zeros(nb, np.uint32); self.block_failures = np.zeros(nb, np.uint32)
        self.block_parsed = np.zeros(nb, np.uint32)
        ng = max(int(batch.n_groups.sum()), 1)
        nvar = np.diff(batch.var_off.astype(np.int64))
        self.cell_capacity = max(int((batch.n_groups.astype(np.int64) * nvar).sum()), 1)
        self.read_off = np.zeros(nb + 1, np.uint64)
        self.read_start = np.zeros(ng, np.uint32); self.read_end = np.zeros(ng, np.uint32)
        self.cell_off = np.zeros(ng + 1, np.uint64)
        self.alleles = np.full(self.cell_capacity, 255, np.uint8); self.quals = np.full(self.cell_capacity, 255, np.uint8)
        self.group_class = np.full(ng, 255, np.uint8); self.group_num_set = np.zeros(ng, np.uint32)
        self._s = hp_realign_out(ptr(self.map_mode, u8p), ptr(self.map_score, u32p), ptr(self.block_disabled_at, u32p),
                                 ptr(self.block_failures, u32p), ptr(self.block_parsed, u32p),
                                 hp_assembled(ptr(self.read_off, u64p), ptr(self.read_start, u32p), ptr(self.read_end, u32p),
                                              ptr(self.cell_off, u64p), ptr(self.alleles, u8p), ptr(self.quals, u8p), self.cell_capacity,
                                              ptr(self.group_class, u8p), ptr(self.group_num_set, u32p), 0, 0))
        self._batch = batch

    def as_struct(self):
        return self._s

    def block_batch(self, ignored=None, is_snv=None):
        """The assembled reads as a BlockBatch ready for astar_solve_batch."""
        a = self._s.assembled
        nr, nc, nv = int(a.n_reads), int(a.n_cells), int(self._batch.var_off[-1])
        return BlockBatch(self._batch.var_off, self.read_off, self.read_start[:nr], self.read_end[:nr], self.cell_off[:nr + 1],
                          self.alleles[:nc], self.quals[:nc], np.zeros(nv, np.uint8) if ignored is None else ignored,
                          np.ones(nv, np.uint8) if is_snv is None else is_snv)


# ---- CIGAR projection of global realignment (read_parsing.rs:672-742) ---------------------------------------------------
class hp_plan_batch(C.Structure):
    _fields_ = [("n_maps", C.c_uint32), ("map_block", u32p), ("seg_off", u64p), ("seg_ref_start", i64p), ("seg_read_start", u32p),
                ("seg_len", u32p), ("n_blocks", C.c_uint32), ("het_first", u32p), ("het_pos", i64p), ("hom_first", u32p), ("hom_pos", i64p)]


class hp_plan_out(C.Structure):
    _fields_ = [("ref_start", u64p), ("ref_end", u64p), ("het_lo", u32p), ("het_hi", u32p), ("hom_lo", u32p), ("hom_hi", u32p),
                ("read_start", u32p), ("read_end", u32p)]


class PlanBatch:
    def __init__(self, map_block, seg_off, seg_ref_start, seg_read_start, seg_len, het_first, het_pos, hom_first, hom_pos):
        self.map_block = _np(map_block, np.uint32); self.seg_off = _np(seg_off, np.uint64)
        self.seg_ref_start = _np(seg_ref_start, np.int64); self.seg_read_start = _np(seg_read_start, np.uint32)
        self.seg_len = _np(seg_len, np.uint32)
        self.het_first = _np(het_first, np.uint32); self.het_pos = _np(het_pos, np.int64)
        self.hom_first = _np(hom_first, np.uint32); self.hom_pos = _np(hom_pos, np.int64)
        self.n_maps, self.n_blocks = len(self.map_block), len(self.het_first) - 1

    def as_struct(self):
        return hp_plan_batch(self.n_maps, ptr(self.map_block, u32p), ptr(self.seg_off, u64p), ptr(self.seg_ref_start, i64p),
                             ptr(self.seg_read_start, u32p), ptr(self.seg_len, u32p), self.n_blocks, ptr(self.het_first, u32p),
                             ptr(self.het_pos, i64p), ptr(self.hom_first, u32p), ptr(self.hom_pos, i64p))


class PlanOut:
    FIELDS = ("ref_start", "ref_end", "het_lo", "het_hi", "hom_lo", "hom_hi", "read_start", "read_end")

    def __init__(self, batch):
        n = max(batch.n_maps, 1)
        self.ref_start = np.zeros(n, np.uint64); self.ref_end = np.zeros(n, np.uint64)
        for f in self.FIELDS[2:]:
            setattr(self, f, np.zeros(n, np.uint32))

    def as_struct(self):
        return hp_plan_out(ptr(self.ref_start, u64p), ptr(self.ref_end, u64p), *[ptr(getattr(self, f), u32p) for f in self.FIELDS[2:]])
