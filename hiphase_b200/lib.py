"""Loader for the product library hiphase_b200/csrc/libhiphase_b200.so (the C ABI of include/hiphase_b200.h).

There is no CPU fallback: if the shared library is missing, or no B200 is visible, every call fails loudly.
"""
import ctypes as C
import os
import subprocess

from . import _abi as A

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
# HP_B200_LIB: kernel-development override (A/B builds of the same ABI); the product path is the in-tree .so
LIB_PATH = os.environ.get("HP_B200_LIB") or os.path.join(CSRC, "libhiphase_b200.so")

EXPORTS = ("hp_abi_version", "hp_build_info", "hp_default_params", "hp_ctx_create", "hp_ctx_destroy", "hp_last_error",
           "hp_astar_solve_batch", "hp_astar_solve_device", "hp_astar_solve_one", "hp_launch_count",
           "hp_last_kernel_ms", "hp_wfa_align_batch", "hp_wfa_graph_align", "hp_post_solve_batch",
           "hp_local_realign_batch", "hp_edit_distance_batch", "hp_assemble_blocks", "hp_pack_write_blocks", "hp_pack_open",
           "hp_pack_get_blocks", "hp_pack_close", "hp_pack_last_error", "hp_write_phase_stats",
           "hp_ctx_set_lanes", "hp_astar_submit", "hp_astar_poll", "hp_astar_wait", "hp_host_alloc", "hp_host_free",
           "hp_host_register", "hp_host_unregister", "hp_block_costs", "hp_lpt_partition", "hp_comm_unique_id",
           "hp_comm_init", "hp_comm_destroy", "hp_comm_allgather", "hp_comm_gather_results", "hp_realign_block_batch", "hp_wfa_plan_batch",
           "hp_service_create", "hp_service_solve_one", "hp_service_counters", "hp_service_destroy")

_LIB = None


class HiPhaseB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("hiphase_b200 error %d: %s" % (code, msg))
        self.code = code


def build(force=False):
    """Compile the CUDA extension for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.run(["make", "-s", "-j8", "-C", CSRC, "all"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise HiPhaseB200Error(A.HP_ERR_INTERNAL, "CUDA extension not built: %s is missing (run __graft_entry__.build())" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.hp_abi_version.restype = C.c_int
        L.hp_default_params.argtypes = [C.POINTER(A.hp_params)]
        L.hp_ctx_create.argtypes = [C.POINTER(A.hp_params), C.c_int, C.POINTER(C.c_void_p)]
        L.hp_ctx_destroy.argtypes = [C.c_void_p]
        L.hp_last_error.restype = C.c_char_p
        L.hp_build_info.restype = C.c_char_p
        L.hp_last_error.argtypes = [C.c_void_p]
        L.hp_astar_solve_batch.argtypes = [C.c_void_p, C.POINTER(A.hp_block_batch), C.POINTER(A.hp_astar_out)]
        L.hp_astar_solve_device.argtypes = [C.c_void_p, C.POINTER(A.hp_block_batch), C.c_uint64, C.c_uint64, C.c_uint64,
                                            C.c_uint32, C.POINTER(A.hp_astar_out), C.c_void_p]
        L.hp_astar_solve_one.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, A.u32p, A.u32p, A.u64p, A.u8p, A.u8p, A.u8p,
                                         A.u8p, A.u8p, A.u8p, C.POINTER(A.hp_phase_stats)]
        L.hp_launch_count.restype = C.c_uint64
        L.hp_launch_count.argtypes = [C.c_void_p]
        L.hp_last_kernel_ms.restype = C.c_float
        L.hp_last_kernel_ms.argtypes = [C.c_void_p]
        L.hp_wfa_align_batch.argtypes = [C.c_void_p, C.POINTER(A.hp_wfa_batch), C.POINTER(A.hp_wfa_out)]
        L.hp_post_solve_batch.argtypes = [C.c_void_p, C.POINTER(A.hp_block_batch), A.i64p, A.u8p, A.u8p, C.POINTER(A.hp_post_out)]
        L.hp_wfa_graph_align.argtypes = [C.c_void_p, C.c_uint32, A.u8p, A.u64p, A.u32p, A.u64p, A.u8p, C.c_uint64,
                                         C.c_uint64, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), A.u64p]
        L.hp_local_realign_batch.argtypes = [C.c_void_p, C.POINTER(A.hp_local_batch), C.POINTER(A.hp_local_out)]
        L.hp_edit_distance_batch.argtypes = [C.c_void_p, C.c_uint32, A.u8p, C.c_uint64, A.u64p, A.u32p, A.u64p, A.u32p, A.u32p]
        L.hp_assemble_blocks.argtypes = [C.c_void_p, C.POINTER(A.hp_rows_batch), C.POINTER(A.hp_assembled)]
        L.hp_pack_write_blocks.argtypes = [C.c_char_p, C.POINTER(A.hp_block_batch), A.i64p]
        L.hp_pack_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        L.hp_pack_get_blocks.argtypes = [C.c_void_p, C.POINTER(A.hp_block_batch), C.POINTER(A.i64p)]
        L.hp_pack_close.argtypes = [C.c_void_p]
        L.hp_pack_last_error.restype = C.c_char_p
        L.hp_write_phase_stats.argtypes = [C.c_char_p, C.POINTER(A.hp_block_batch), A.i64p, C.POINTER(A.hp_astar_out), C.c_uint64]
        L.hp_ctx_set_lanes.argtypes = [C.c_void_p, C.c_int]
        L.hp_astar_submit.argtypes = [C.c_void_p, C.POINTER(A.hp_block_batch), C.POINTER(A.hp_astar_out), C.POINTER(C.c_void_p)]
        L.hp_astar_poll.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        L.hp_astar_wait.argtypes = [C.c_void_p, C.c_void_p]
        L.hp_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.hp_host_free.argtypes = [C.c_void_p]
        L.hp_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.hp_host_unregister.argtypes = [C.c_void_p]
        L.hp_block_costs.argtypes = [C.c_uint64, A.u32p, A.u64p, A.u64p]
        L.hp_lpt_partition.argtypes = [A.u64p, C.c_uint64, C.c_uint32, A.u32p]
        L.hp_comm_unique_id.argtypes = [A.u8p]
        L.hp_comm_init.argtypes = [C.c_void_p, A.u8p, C.c_int, C.c_int]
        L.hp_comm_destroy.argtypes = [C.c_void_p]
        L.hp_comm_allgather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.hp_service_create.argtypes = [C.POINTER(A.hp_params), C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
        L.hp_service_solve_one.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, A.u32p, A.u32p, A.u64p, A.u8p, A.u8p, A.u8p, A.u8p, A.u8p, A.u8p,
                                           C.POINTER(A.hp_phase_stats)]
        L.hp_service_counters.argtypes = [C.c_void_p, A.u64p, A.u64p]
        L.hp_service_destroy.argtypes = [C.c_void_p]
        L.hp_service_destroy.restype = None
        L.hp_wfa_plan_batch.argtypes = [C.c_void_p, C.POINTER(A.hp_plan_batch), C.POINTER(A.hp_plan_out)]
        L.hp_realign_block_batch.argtypes = [C.c_void_p, C.POINTER(A.hp_realign_batch), C.POINTER(A.hp_realign_out)]
        L.hp_comm_gather_results.argtypes = [C.c_void_p, C.c_uint64, A.u64p, A.u64p, C.POINTER(A.hp_astar_out), C.c_uint64,
                                             A.u64p, C.c_int, C.POINTER(A.hp_astar_out)]
        _LIB = L
    return _LIB


# ---- host-only helpers of the sharding layer (no GPU needed) ------------------------------------------------------
def block_costs(n_var, n_cells):
    """hp_block_costs: serial-chain cost model of each block (cells * min(N, 40) + N)."""
    import numpy as np
    n_var = np.ascontiguousarray(n_var, np.uint32); n_cells = np.ascontiguousarray(n_cells, np.uint64)
    cost = np.zeros(len(n_var), np.uint64)
    rc = lib().hp_block_costs(len(n_var), A.ptr(n_var, A.u32p), A.ptr(n_cells, A.u64p), A.ptr(cost, A.u64p))
    if rc != A.HP_OK:
        raise HiPhaseB200Error(rc, "hp_block_costs")
    return cost


def lpt_partition(costs, n_shards):
    """hp_lpt_partition -> shard_of[n] (uint32)."""
    import numpy as np
    costs = np.ascontiguousarray(costs, np.uint64)
    shard_of = np.zeros(len(costs), np.uint32)
    rc = lib().hp_lpt_partition(A.ptr(costs, A.u64p), len(costs), int(n_shards), A.ptr(shard_of, A.u32p))
    if rc != A.HP_OK:
        raise HiPhaseB200Error(rc, "hp_lpt_partition")
    return shard_of


def comm_unique_id():
    """hp_comm_unique_id: the 128-byte NCCL id rank 0 hands to every rank (any out-of-band channel)."""
    import numpy as np
    uid = np.zeros(A.HP_COMM_ID_BYTES, np.uint8)
    rc = lib().hp_comm_unique_id(A.ptr(uid, A.u8p))
    if rc != A.HP_OK:
        raise HiPhaseB200Error(rc, (lib().hp_last_error(None) or b"").decode())
    return uid


class PinnedArena:
    """Pinned host memory from hp_host_alloc, handed out as numpy views (kept alive by the arena)."""

    def __init__(self):
        self._ptrs = []

    def alloc(self, nbytes):
        import numpy as np
        p = C.c_void_p()
        rc = lib().hp_host_alloc(C.byref(p), int(max(nbytes, 1)))
        if rc != A.HP_OK:
            raise HiPhaseB200Error(rc, "hp_host_alloc(%d)" % nbytes)
        self._ptrs.append(p)
        return np.ctypeslib.as_array(C.cast(p, A.u8p), shape=(int(max(nbytes, 1)),))[:nbytes]

    def empty(self, n, dtype):
        import numpy as np
        dt = np.dtype(dtype)
        return self.alloc(int(n) * dt.itemsize).view(dt)

    def copy(self, arr):
        out = self.empty(arr.size, arr.dtype)
        out[...] = arr.reshape(-1)
        return out

    def close(self):
        for p in self._ptrs:
            lib().hp_host_free(p)
        self._ptrs = []


class BlockService:
    """hp_service: the blocking, thread-safe one-block call (the reference's astar_solver call shape) on top of batched launches."""

    def __init__(self, params=None, device=0, max_batch_blocks=0, linger_us=0):
        self._h = C.c_void_p()
        rc = lib().hp_service_create(C.byref(params) if params is not None else None, int(device), int(max_batch_blocks), int(linger_us), C.byref(self._h))
        if rc != A.HP_OK:
            raise HiPhaseB200Error(rc, (lib().hp_last_error(None) or b"").decode())

    def solve_one(self, block_batch, b=0):
        """Block b of a BlockBatch -> (h1, h2, stats record); callable from many threads at once."""
        import numpy as np
        v0, v1 = int(block_batch.var_off[b]), int(block_batch.var_off[b + 1])
        r0, r1 = int(block_batch.read_off[b]), int(block_batch.read_off[b + 1])
        h1 = np.zeros(v1 - v0, np.uint8); h2 = np.zeros(v1 - v0, np.uint8)
        st = A.hp_phase_stats()
        rs = np.ascontiguousarray(block_batch.read_start[r0:r1]); re = np.ascontiguousarray(block_batch.read_end[r0:r1])
        co = np.ascontiguousarray(block_batch.cell_off[r0:r1 + 1])
        ig = np.ascontiguousarray(block_batch.ignored[v0:v1]); sn = np.ascontiguousarray(block_batch.is_snv[v0:v1])
        rc = lib().hp_service_solve_one(self._h, v1 - v0, r1 - r0, A.ptr(rs, A.u32p), A.ptr(re, A.u32p), A.ptr(co, A.u64p),
                                        A.ptr(block_batch.alleles, A.u8p), A.ptr(block_batch.quals, A.u8p), A.ptr(ig, A.u8p), A.ptr(sn, A.u8p),
                                        A.ptr(h1, A.u8p), A.ptr(h2, A.u8p), C.byref(st))
        if rc != A.HP_OK:
            raise HiPhaseB200Error(rc, "hp_service_solve_one")
        return h1, h2, st

    def counters(self):
        import numpy as np
        a = np.zeros(1, np.uint64); b = np.zeros(1, np.uint64)
        lib().hp_service_counters(self._h, A.ptr(a, A.u64p), A.ptr(b, A.u64p))
        return int(a[0]), int(b[0])

    def close(self):
        if self._h:
            lib().hp_service_destroy(self._h)
            self._h = C.c_void_p()


class Context:
    """hp_ctx handle.  One per thread (src/main.rs:385-408 runs one solve_block per worker)."""

    def __init__(self, params=None, device=-1):
        self.params = params or A.default_params()
        self._h = C.c_void_p()
        rc = lib().hp_ctx_create(C.byref(self.params), int(device), C.byref(self._h))
        if rc != A.HP_OK:
            raise HiPhaseB200Error(rc, (lib().hp_last_error(None) or b"").decode())

    def close(self):
        if self._h:
            lib().hp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != A.HP_OK:
            raise HiPhaseB200Error(rc, (lib().hp_last_error(self._h) or b"").decode())

    @property
    def handle(self):
        return self._h

    def set_team(self, team):
        """Profiling / test aid: force the speculative team size of the A* kernel (0 = automatic, 1, 2 or 4)."""
        self.check(lib().hp_debug_set_team(self._h, int(team)))

    def set_wfa_build_mode(self, mode):
        """Test / A-B aid: bit 0 = build the WFA graphs on the host, bit 1 = start without a workspace hint."""
        self.check(lib().hp_debug_wfa_build_mode(self._h, int(mode)))

    def set_wfa_filter(self, on):
        """Test / A-B aid: the piece filter that proves MaxEditDistance for hopeless reads without aligning them."""
        self.check(lib().hp_debug_wfa_filter(self._h, int(bool(on))))

    def wfa_filtered(self):
        """Reads the piece filter answered in the last wfa_align_batch call."""
        lib().hp_debug_wfa_filtered.restype = C.c_uint32
        return int(lib().hp_debug_wfa_filtered(self._h))

    def launch_count(self):
        return int(lib().hp_launch_count(self._h))

    def last_kernel_ms(self):
        return float(lib().hp_last_kernel_ms(self._h))

    # ---- A* ----
    def astar_solve_batch(self, batch, want_heuristic=False, want_counters=False):
        """Host buffers in / out (H2D + kernels + D2H): the end-to-end call.  Returns an AstarOut."""
        out = A.AstarOut(batch, want_heuristic, want_counters)
        bs, os_ = batch.as_struct(), out.as_struct()
        self.check(lib().hp_astar_solve_batch(self._h, C.byref(bs), C.byref(os_)))
        return out

    def set_lanes(self, lanes):
        """Number of batches that may be in flight on the device at once (hp_ctx_set_lanes)."""
        self.check(lib().hp_ctx_set_lanes(self._h, int(lanes)))

    def astar_submit(self, batch, out=None, want_heuristic=False, want_counters=False):
        """hp_astar_submit: enqueue one host batch, return a job handle (keeps batch / out alive until waited)."""
        out = out if out is not None else A.AstarOut(batch, want_heuristic, want_counters)
        bs, os_ = batch.as_struct(), out.as_struct()
        job = C.c_void_p()
        self.check(lib().hp_astar_submit(self._h, C.byref(bs), C.byref(os_), C.byref(job)))
        return (job, batch, out)

    def astar_poll(self, handle):
        done = C.c_int(0)
        self.check(lib().hp_astar_poll(self._h, handle[0], C.byref(done)))
        return bool(done.value)

    def astar_wait(self, handle):
        """hp_astar_wait: block until the job's results are in its AstarOut; returns it."""
        self.check(lib().hp_astar_wait(self._h, handle[0]))
        return handle[2]

    # ---- multi-GPU hand-off (NCCL communicator owned by the context) ----
    def comm_init(self, unique_id, rank, world):
        import numpy as np
        uid = np.ascontiguousarray(unique_id, np.uint8) if unique_id is not None else None
        self.check(lib().hp_comm_init(self._h, A.ptr(uid, A.u8p), int(rank), int(world)))

    def comm_allgather(self, send, recv):
        self.check(lib().hp_comm_allgather(self._h, send.ctypes.data_as(C.c_void_p), recv.ctypes.data_as(C.c_void_p), send.nbytes))

    def comm_gather_results(self, local_ids, local_batch, local_out, all_var_off, all_out=None, root=-1, rank=0):
        """hp_comm_gather_results: the results of all blocks ordered by global block index, on every rank (root < 0) or on
        rank `root` only (the other ranks get None)."""
        import numpy as np
        local_ids = np.ascontiguousarray(local_ids, np.uint64)
        all_var_off = np.ascontiguousarray(all_var_off, np.uint64)
        n_total = len(all_var_off) - 1
        receive = root < 0 or root == rank
        if all_out is None and receive:
            all_out = A.AstarOut.sized(int(all_var_off[-1]), n_total)
        ls = local_out.as_struct()
        as_ = all_out.as_struct() if receive else None
        self.check(lib().hp_comm_gather_results(self._h, len(local_ids), A.ptr(local_ids, A.u64p), A.ptr(local_batch.var_off, A.u64p),
                                                C.byref(ls), n_total, A.ptr(all_var_off, A.u64p), int(root),
                                                C.byref(as_) if receive else None))
        return all_out if receive else None

    # ---- post-solve (span counts, block tags, haplotags) ----
    def post_solve_batch(self, batch, var_pos, h1, h2):
        import numpy as np
        var_pos = np.ascontiguousarray(var_pos, np.int64); h1 = np.ascontiguousarray(h1, np.uint8); h2 = np.ascontiguousarray(h2, np.uint8)
        out = A.PostOut(batch)
        bs, os_ = batch.as_struct(), out.as_struct()
        self.check(lib().hp_post_solve_batch(self._h, C.byref(bs), A.ptr(var_pos, A.i64p), A.ptr(h1, A.u8p), A.ptr(h2, A.u8p), C.byref(os_)))
        return out

    # ---- graph-WFA ----
    def wfa_align_batch(self, batch, trav_words=0, want_counters=False):
        """Host buffers in / out: graph construction + alignment + allele/qual rows.  Returns a WfaOut."""
        out = A.WfaOut(batch, trav_words, want_counters)
        bs, os_ = batch.as_struct(), out.as_struct()
        self.check(lib().hp_wfa_align_batch(self._h, C.byref(bs), C.byref(os_)))
        return out

    def wfa_graph_align(self, seq, seq_off, parent_idx, parent_off, read, prune_distance=None, max_edit_distance=1000):
        """Pre-built graph (WFAGraph::add_node order) vs one read -> (status, score, sorted traversed node ids)."""
        import numpy as np
        seq = np.ascontiguousarray(seq, np.uint8); seq_off = np.ascontiguousarray(seq_off, np.uint64)
        parent_idx = np.ascontiguousarray(parent_idx, np.uint32); parent_off = np.ascontiguousarray(parent_off, np.uint64)
        read = np.ascontiguousarray(read, np.uint8)
        n = len(seq_off) - 1
        trav = np.zeros((n + 63) // 64, np.uint64)
        st, sc = C.c_int32(-1), C.c_uint32(0)
        self.check(lib().hp_wfa_graph_align(self._h, n, A.ptr(seq, A.u8p) if len(seq) else A.u8p(), A.ptr(seq_off, A.u64p),
                                            A.ptr(parent_idx, A.u32p) if len(parent_idx) else A.u32p(), A.ptr(parent_off, A.u64p),
                                            A.ptr(read, A.u8p) if len(read) else A.u8p(), len(read),
                                            (1 << 64) - 1 if prune_distance is None else prune_distance, max_edit_distance,
                                            C.byref(st), C.byref(sc), A.ptr(trav, A.u64p)))
        nodes = [i for i in range(n) if (int(trav[i // 64]) >> (i % 64)) & 1]
        return st.value, sc.value, nodes

    # ---- local realignment ----
    def local_realign_batch(self, batch):
        """Host buffers in / out: local_realignment (read_parsing.rs:121-503) for a batch of read mappings.  Returns a LocalOut."""
        out = A.LocalOut(batch)
        bs, os_ = batch.as_struct(), out.as_struct()
        self.check(lib().hp_local_realign_batch(self._h, C.byref(bs), C.byref(os_)))
        return out

    def edit_distance_batch(self, pairs):
        """sequence_alignment::edit_distance (sequence_alignment.rs:6-38) for a list of (a, b) byte sequences."""
        import numpy as np
        seqs, a_off, a_len, b_off, b_len, pos = [], [], [], [], [], 0
        for x, y in pairs:
            for off, ln, z in ((a_off, a_len, x), (b_off, b_len, y)):
                z = np.asarray(list(z), np.uint8)
                off.append(pos); ln.append(len(z)); seqs.append(z); pos += len(z)
        blob = np.concatenate(seqs) if pos else np.zeros(1, np.uint8)
        a_off = np.asarray(a_off, np.uint64); b_off = np.asarray(b_off, np.uint64)
        a_len = np.asarray(a_len, np.uint32); b_len = np.asarray(b_len, np.uint32)
        dist = np.zeros(len(pairs), np.uint32)
        self.check(lib().hp_edit_distance_batch(self._h, len(pairs), A.ptr(blob, A.u8p), pos, A.ptr(a_off, A.u64p), A.ptr(a_len, A.u32p),
                                                A.ptr(b_off, A.u64p), A.ptr(b_len, A.u32p), A.ptr(dist, A.u32p)))
        return dist

    # ---- matrix assembly ----
    def assemble_blocks(self, rows):
        """ReadSegment::new + collapse + the min-matched-alleles filter for a RowsBatch.  Returns an Assembled."""
        out = A.Assembled(rows)
        rs = rows.as_struct()
        self.check(lib().hp_assemble_blocks(self._h, C.byref(rs), C.byref(out.as_struct())))
        return out

    # ---- realignment pipeline (WFA -> local fallback -> switch-off rule -> assembly) ----
    def realign_block_batch(self, batch):
        """hp_realign_block_batch for a RealignBatch.  Returns a RealignOut (block_batch() feeds astar_solve_batch)."""
        out = A.RealignOut(batch)
        bs = batch.as_struct()
        self.check(lib().hp_realign_block_batch(self._h, C.byref(bs), C.byref(out.as_struct())))
        return out

    def wfa_plan_batch(self, batch):
        """hp_wfa_plan_batch: the CIGAR projection of global realignment for a PlanBatch.  Returns a PlanOut."""
        out = A.PlanOut(batch)
        bs = batch.as_struct()
        self.check(lib().hp_wfa_plan_batch(self._h, C.byref(bs), C.byref(out.as_struct())))
        return out

    def astar_solve_device(self, dev_batch_struct, n_vars, n_reads, n_cells, max_block_vars, dev_out_struct, stream):
        self.check(lib().hp_astar_solve_device(self._h, C.byref(dev_batch_struct), n_vars, n_reads, n_cells,
                                               max_block_vars, C.byref(dev_out_struct), C.c_void_p(stream)))


# ---- packed phase-block container (host only) -------------------------------------------------
RUNNER_PATH = os.path.join(CSRC, "hp_phase_blocks")


def pack_write_blocks(path, batch, var_pos=None):
    """hp_pack_write_blocks: a BlockBatch (+ optional variant positions) -> HPB200 file."""
    import numpy as np
    bs = batch.as_struct()
    vp = None if var_pos is None else np.ascontiguousarray(var_pos, np.int64)
    rc = lib().hp_pack_write_blocks(os.fsencode(path), C.byref(bs), A.ptr(vp, A.i64p))
    if rc != A.HP_OK:
        raise HiPhaseB200Error(rc, (lib().hp_pack_last_error() or b"").decode())


def pack_read_blocks(path):
    """hp_pack_open + hp_pack_get_blocks -> (BlockBatch, var_pos or None); the arrays are copied out of the container."""
    import numpy as np
    h = C.c_void_p()
    rc = lib().hp_pack_open(os.fsencode(path), C.byref(h))
    if rc != A.HP_OK:
        raise HiPhaseB200Error(rc, (lib().hp_pack_last_error() or b"").decode())
    try:
        bs, vp = A.hp_block_batch(), A.i64p()
        lib().hp_pack_get_blocks(h, C.byref(bs), C.byref(vp))
        nb = bs.n_blocks

        def arr(p, n, dt):
            return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)
        var_off = arr(bs.var_off, nb + 1, np.uint64); read_off = arr(bs.read_off, nb + 1, np.uint64)
        nv, nr = int(var_off[-1]), int(read_off[-1])
        cell_off = arr(bs.cell_off, nr + 1, np.uint64)
        nc = int(cell_off[-1])
        batch = A.BlockBatch(var_off, read_off, arr(bs.read_start, nr, np.uint32), arr(bs.read_end, nr, np.uint32), cell_off,
                             arr(bs.alleles, nc, np.uint8), arr(bs.quals, nc, np.uint8), arr(bs.ignored, nv, np.uint8), arr(bs.is_snv, nv, np.uint8))
        var_pos = arr(vp, nv, np.int64) if vp else None
        return batch, var_pos
    finally:
        lib().hp_pack_close(h)


def write_phase_stats(path, batch, out, var_pos=None, first_block_index=0):
    """hp_write_phase_stats: the solver-side columns of HiPhase's --stats-file, one row per block."""
    import numpy as np
    bs, os_ = batch.as_struct(), out.as_struct()
    vp = None if var_pos is None else np.ascontiguousarray(var_pos, np.int64)
    rc = lib().hp_write_phase_stats(os.fsencode(path), C.byref(bs), A.ptr(vp, A.i64p), C.byref(os_), int(first_block_index))
    if rc != A.HP_OK:
        raise HiPhaseB200Error(rc, (lib().hp_pack_last_error() or b"").decode())
