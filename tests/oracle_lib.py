"""ctypes loader for the CPU oracle (oracle/libhp_oracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

from hiphase_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "libhp_oracle.so")
        if not os.path.exists(path) or os.path.exists("/root/reference"):
            # in the build container always make sure it is fresh; on the GPU box use the prebuilt file
            try:
                build_oracle()
            except Exception:
                if not os.path.exists(path):
                    raise
        L = C.CDLL(path)
        L.hpo_score_partial.restype = C.c_uint64
        L.hpo_score_partial.argtypes = [C.c_uint64, C.c_uint64, A.u8p, A.u8p, A.u8p, C.c_uint64, C.c_uint64]
        L.hpo_graph_new.restype = C.c_void_p
        L.hpo_graph_new.argtypes = [C.c_uint64]
        L.hpo_graph_free.argtypes = [C.c_void_p]
        L.hpo_graph_add_node.restype = C.c_int64
        L.hpo_graph_add_node.argtypes = [C.c_void_p, A.u8p, C.c_uint64, A.u64p, C.c_uint32]
        L.hpo_graph_num_nodes.restype = C.c_uint64
        L.hpo_graph_num_nodes.argtypes = [C.c_void_p]
        L.hpo_graph_from_job.restype = C.c_void_p
        L.hpo_graph_from_job.argtypes = [C.POINTER(A.hp_wfa_batch), C.c_uint32, C.c_uint64]
        L.hpo_graph_allele_map.restype = C.c_uint64
        L.hpo_graph_allele_map.argtypes = [C.c_void_p, A.u64p, C.c_uint64]
        L.hpo_graph_sizes.argtypes = [C.c_void_p, A.u64p, A.u64p, A.u64p]
        L.hpo_graph_flatten.argtypes = [C.c_void_p, A.u8p, A.u64p, A.u32p, A.u64p]
        L.hpo_graph_edit_distance.argtypes = [C.c_void_p, A.u8p, C.c_uint64, C.c_uint64, A.u64p, A.u64p, A.u64p,
                                              C.POINTER(A.hp_wfa_counters), C.c_uint64]
        L.hpo_astar_solve_batch.argtypes = [C.POINTER(A.hp_params), C.POINTER(A.hp_block_batch),
                                            C.POINTER(A.hp_astar_out), C.c_int]
        L.hpo_astar_subsolver.argtypes = [C.POINTER(A.hp_params), C.POINTER(A.hp_block_batch), C.c_uint64, C.c_uint64,
                                          A.u64p, A.u64p, A.u64p]
        L.hpo_astar_node_path.argtypes = [C.POINTER(A.hp_block_batch), A.u64p, C.c_uint32, A.u8p, A.u8p, A.u64p,
                                          A.u64p, A.u64p]
        L.hpo_tracker_script.argtypes = [C.c_uint32, C.c_uint32, A.u8p, A.u32p, A.u64p]
        L.hpo_collapse.argtypes = [C.c_uint32, C.c_uint64, A.u8p, A.u8p, A.u8p, A.u8p, A.u64p, A.u64p]
        L.hpo_read_segment_region.argtypes = [A.u8p, C.c_uint64, A.u64p, A.u64p]
        L.hpo_post_solve_batch.argtypes = [C.POINTER(A.hp_block_batch), A.i64p, A.u8p, A.u8p, C.POINTER(A.hp_post_out)]
        L.hpo_wfa_align_batch.argtypes = [C.POINTER(A.hp_params), C.POINTER(A.hp_wfa_batch), C.POINTER(A.hp_wfa_out),
                                          C.c_int]
        L.hpo_edit_distance.restype = C.c_uint64
        L.hpo_edit_distance.argtypes = [A.u8p, C.c_uint64, A.u8p, C.c_uint64]
        L.hpo_match_allele.restype = C.c_uint8
        L.hpo_match_allele.argtypes = [C.POINTER(A.hp_local_batch), C.c_uint32, A.u8p, C.c_uint64]
        L.hpo_closest_allele_clip.restype = C.c_uint8
        L.hpo_closest_allele_clip.argtypes = [C.POINTER(A.hp_local_batch), C.c_uint32, A.u8p, C.c_uint64, C.c_uint64,
                                              C.c_uint64, A.u64p, A.u64p]
        L.hpo_local_realign_batch.argtypes = [C.POINTER(A.hp_local_batch), C.POINTER(A.hp_local_out)]
        L.hpo_assemble_blocks.argtypes = [C.POINTER(A.hp_rows_batch), C.POINTER(A.hp_assembled)]
        _LIB = L
    return _LIB


def u8(x):
    return np.ascontiguousarray(x, dtype=np.uint8)


def astar_solve(batch, params=None, threads=1, want_heuristic=True, want_counters=True):
    """Oracle astar_solver over a BlockBatch -> AstarOut (numpy)."""
    params = params or A.default_params()
    out = A.AstarOut(batch, want_heuristic, want_counters)
    bs, os_ = batch.as_struct(), out.as_struct()
    out.failures = lib().hpo_astar_solve_batch(C.byref(params), C.byref(bs), C.byref(os_), int(threads))
    return out


def region(alleles):
    a = u8(alleles)
    s, e = C.c_uint64(), C.c_uint64()
    lib().hpo_read_segment_region(A.ptr(a, A.u8p), len(a), C.byref(s), C.byref(e))
    return s.value, e.value


def score_partial(start, alleles, quals, hap, offset):
    a, q, h = u8(alleles), u8(quals), u8(hap)
    return lib().hpo_score_partial(start, start + len(a), A.ptr(a, A.u8p), A.ptr(q, A.u8p), A.ptr(h, A.u8p), len(h), offset)


def collapse(rows_alleles, rows_quals):
    a, q = u8(rows_alleles), u8(rows_quals)
    k, n = a.shape
    oa, oq = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    s, e = C.c_uint64(), C.c_uint64()
    rc = lib().hpo_collapse(k, n, A.ptr(a, A.u8p), A.ptr(q, A.u8p), A.ptr(oa, A.u8p), A.ptr(oq, A.u8p), C.byref(s), C.byref(e))
    assert rc == 0
    return oa, oq, (s.value, e.value)


class Graph:
    """Oracle WFAGraph handle (src/wfa_graph.rs:61-68)."""

    def __init__(self, max_ed=1000, handle=None):
        self.h = handle if handle is not None else lib().hpo_graph_new(max_ed)

    def __del__(self):
        if getattr(self, "h", None):
            lib().hpo_graph_free(self.h)
            self.h = None

    def add_node(self, seq, parents):
        s = u8(list(seq))
        p = np.ascontiguousarray(parents, dtype=np.uint64)
        return lib().hpo_graph_add_node(self.h, A.ptr(s, A.u8p), len(s), A.ptr(p, A.u64p), len(p))

    def num_nodes(self):
        return lib().hpo_graph_num_nodes(self.h)

    def edit_distance(self, read, prune=None, shuffle_seed=0, counters=False):
        r = u8(list(read))
        n = self.num_nodes()
        score, ntrav = C.c_uint64(), C.c_uint64()
        trav = np.zeros(max(n, 1), np.uint64)
        ctr = A.hp_wfa_counters()
        st = lib().hpo_graph_edit_distance(self.h, A.ptr(r, A.u8p), len(r), (1 << 64) - 1 if prune is None else prune,
                                           C.byref(score), A.ptr(trav, A.u64p), C.byref(ntrav), C.byref(ctr), shuffle_seed)
        res = (st, score.value, [int(x) for x in trav[:ntrav.value]])
        return res + (ctr,) if counters else res

    def allele_map(self):
        cap = 4 * max(self.num_nodes(), 1) + 16
        t = np.zeros(3 * cap, np.uint64)
        n = lib().hpo_graph_allele_map(self.h, A.ptr(t, A.u64p), cap)
        m = {}
        for i in range(n):
            m.setdefault(int(t[3 * i]), []).append((int(t[3 * i + 1]), int(t[3 * i + 2])))
        return m

    def flatten(self):
        nn, ns, npar = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().hpo_graph_sizes(self.h, C.byref(nn), C.byref(ns), C.byref(npar))
        seq = np.zeros(max(ns.value, 1), np.uint8)
        seq_off = np.zeros(nn.value + 1, np.uint64)
        par = np.zeros(max(npar.value, 1), np.uint32)
        par_off = np.zeros(nn.value + 1, np.uint64)
        lib().hpo_graph_flatten(self.h, A.ptr(seq, A.u8p), A.ptr(seq_off, A.u64p), A.ptr(par, A.u32p), A.ptr(par_off, A.u64p))
        return seq, seq_off, par, par_off

    @staticmethod
    def from_job(wfa_batch, job, max_ed=1000):
        bs = wfa_batch.as_struct()
        h = lib().hpo_graph_from_job(C.byref(bs), job, max_ed)
        if not h:
            return None
        return Graph(handle=h)


def wfa_align(batch, params=None, threads=1, trav_words=0, want_counters=False):
    params = params or A.default_params()
    out = A.WfaOut(batch, trav_words, want_counters)
    bs, os_ = batch.as_struct(), out.as_struct()
    out.failures = lib().hpo_wfa_align_batch(C.byref(params), C.byref(bs), C.byref(os_), int(threads))
    return out


def post_solve(batch, var_pos, h1, h2):
    var_pos = np.ascontiguousarray(var_pos, np.int64); h1 = u8(h1); h2 = u8(h2)
    out = A.PostOut(batch)
    bs, os_ = batch.as_struct(), out.as_struct()
    out.rc = lib().hpo_post_solve_batch(C.byref(bs), A.ptr(var_pos, A.i64p), A.ptr(h1, A.u8p), A.ptr(h2, A.u8p), C.byref(os_))
    return out


def edit_distance(a, b):
    a, b = u8(list(a)), u8(list(b))
    return int(lib().hpo_edit_distance(A.ptr(a, A.u8p), len(a), A.ptr(b, A.u8p), len(b)))


def match_allele(local_batch, variant, seq):
    s = u8(list(seq)); bs = local_batch.as_struct()
    return int(lib().hpo_match_allele(C.byref(bs), variant, A.ptr(s, A.u8p), len(s)))


def closest_allele_clip(local_batch, variant, seq, head=0, tail=0):
    """(allele, min distance, other distance) like Variant::closest_allele_clip (variants.rs:624-641)."""
    s = u8(list(seq)); bs = local_batch.as_struct()
    d0, d1 = C.c_uint64(), C.c_uint64()
    a = int(lib().hpo_closest_allele_clip(C.byref(bs), variant, A.ptr(s, A.u8p), len(s), head, tail, C.byref(d0), C.byref(d1)))
    return a, min(d0.value, d1.value), max(d0.value, d1.value)


def local_realign(batch):
    out = A.LocalOut(batch)
    bs, os_ = batch.as_struct(), out.as_struct()
    out.rc = lib().hpo_local_realign_batch(C.byref(bs), C.byref(os_))
    return out


def assemble_blocks(rows):
    out = A.Assembled(rows)
    rs = rows.as_struct()
    out.rc = lib().hpo_assemble_blocks(C.byref(rs), C.byref(out.as_struct()))
    return out


def realign_block_batch(batch, params=None):
    """The read loop of load_full_read_segments (read_parsing.rs:545-629), one mapping at a time."""
    params = params or A.default_params()
    out = A.RealignOut(batch)
    bs = batch.as_struct()
    out.rc = lib().hpo_realign_block_batch(C.byref(params), C.byref(bs), C.byref(out.as_struct()))
    return out


def wfa_plan_batch(batch):
    out = A.PlanOut(batch)
    bs = batch.as_struct()
    out.rc = lib().hpo_wfa_plan_batch(C.byref(bs), C.byref(out.as_struct()))
    return out
