// hp_host.h -- host-side context shared by the API translation units.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hiphase_b200.h"
#include "hp_device.cuh"

namespace hp {

// Grow-only device buffer.
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes);
    void release();
};

// Grow-only pinned host buffer (staging of the result hand-off).
struct HostPin {
    void* ptr = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes) {
        if (bytes <= cap) return true;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        if (cudaHostAlloc(&ptr, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); ptr = nullptr; return false; }
        cap = want;
        return true;
    }
    void release() { if (ptr) cudaFreeHost(ptr); ptr = nullptr; cap = 0; }
};

// pageable host memory (hp_api.cu)
struct CopySeg { uint8_t* dst; const uint8_t* src; size_t bytes; };
bool host_pageable(const void* p);
void parallel_copy(const std::vector<CopySeg>& segs);
bool upload_large(HostPin& pin, void* dst, const void* src, size_t bytes, cudaStream_t st);

// kernel launchers (astar_kernels.cu)
size_t astar_smem_bytes(uint32_t sub_capl, int team);
int astar_max_team();
int astar_warps_per_sm(bool dense = false);
uint64_t astar_slab_bytes(uint32_t qcap, uint32_t hap_words, uint32_t sub_capl);
cudaError_t launch_astar_prep(PrepArgs pa, uint32_t max_block_vars, cudaStream_t stream);
cudaError_t launch_astar_solve(const AstarArgs& a, int n_ctas, int team, cudaStream_t* streams /* [3]: one per class */, bool dense = false);
int astar_max_ctas_per_sm(uint32_t sub_capl);
cudaError_t launch_score_planes_debug(const AstarArgs& a, uint64_t h1, uint64_t h2, int offset, int L, uint32_t* s1, uint32_t* s2,
                                      cudaStream_t stream);

}  // namespace hp

// One A* lane: a stream and a private set of workspaces.  Batches in flight on different lanes overlap on the device.
struct AstarLane {
    cudaStream_t stream = nullptr, aux[2] = {nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join[2] = {nullptr, nullptr}, ev_done = nullptr, ev_in = nullptr;
    bool used = false;                       // ev_done has been recorded at least once
    bool timing_pending = false;
    hp::DevBuf meta, rmeta, planes, act_off, act_cur, act_idx, col, order, heur, ticket, stage_in, stage_out, dbg;
    // Pageable caller buffers (an ordinary Vec): inputs are gathered into pin_in by a few host threads and go to the device in
    // one copy; results land in pin_out and are handed to the caller's arrays by hp_astar_wait.  (A cudaMemcpyAsync to or from
    // pageable memory is staged by the driver and blocks the calling thread until the stream gets there, which would make
    // hp_astar_submit wait for the kernels.)
    hp::HostPin pin_in, pin_out;
    bool out_staged = false;                 // the batch in flight writes its results to pin_out
    size_t out_off[6] = {0, 0, 0, 0, 0, 0};  // h1, h2, stats, status, heuristic, counters inside pin_out
    struct hp_astar_job* job = nullptr;      // job in flight on this lane (streaming entry)
};

// A job of the streaming entry (hp_astar_submit .. hp_astar_wait).
struct hp_astar_job {
    int lane = 0;
    hp_block_batch batch;                    // the caller's host batch (pointers stay the caller's)
    hp_astar_out out;                        // the caller's host outputs
    uint32_t max_n = 0;
};

struct hp_ctx {
    hp_params params;
    int device = 0;
    int sm_count = 0;
    uint32_t sub_capl = 0;
    uint32_t qcap = 0;
    int force_team = 0;                     // 0 = choose from the batch size; 1/2/4 = fixed team size
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing_pending = false;
    float last_ms = 0.f;
    uint64_t launches = 0;
    std::string err;
    // A* lanes (created on demand) and the slab pool they share: main-queue slabs are acquired per CTA on the device
    std::vector<AstarLane*> lanes;
    int n_lanes = 4;
    int next_lane = 0;
    int last_lane = 0;
    hp::DevBuf slabs, slab_busy;
    uint32_t n_slabs = 0;
    int max_ctas_per_sm = 16;               // solver CTAs that can be resident on one SM (occupancy query)
    uint64_t slab_bytes = 0;
    uint32_t slab_hap_words = 0, slab_qcap = 0;
    bool want_dbg = false;
    uint32_t dbg_blocks = 0;
    // staging / scratch of the other entry points (WFA, local realignment, assembly, post-solve)
    hp::DevBuf ticket, stage_in, stage_out;
    hp::DevBuf ed_scratch;                  // local realignment / edit distance: per-warp delta rows of multi-panel comparisons
    // WFA workspaces
    hp::DevBuf wfa_ws, wfa_in, wfa_out, wfa_graph;
    bool wfa_no_filter = false;             // test aid: never short-circuit hopeless reads (piece filter off)
    uint32_t wfa_filtered = 0;              // reads the piece filter answered in the last hp_wfa_align_batch call
    uint32_t wfa_filtered_now = 0;          // ... of the run in flight
    bool wfa_dbg_times = false;             // profiling aid: the counting variant reports per-job start / end times instead of set_ops / n_nodes
    bool wfa_no_hint = false;               // test aid: start with an unsized graph workspace (exercises the regrow path)
    bool wfa_host_build = false;            // debug / A-B aid: build the graphs on the host instead of on the device
    uint64_t wfa_layout[4] = {0, 0, 0, 0};  // {workspace pointer, slab bytes, table cap, warps} the hash keys were last zeroed for
    uint32_t wfa_table_cap = 1u << 15;      // (node, diagonal) hash slots per warp; grown 8x on overflow
    std::vector<int32_t> wfa_h_status;
    std::vector<uint32_t> wfa_h_score, wfa_h_nodes;
    std::vector<uint8_t> wfa_h_alleles, wfa_h_quals;
    std::vector<uint64_t> wfa_h_trav, wfa_h_ctr;
    // NCCL communicator (hp_comm_*), resolved at run time from libnccl.so.2
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    hp::DevBuf comm_send, comm_recv;
    hp::HostPin pin_send, pin_recv;
    // hp_realign_block_batch: pinned scratch for the bases / qualities of the reads gathered for local realignment (grow-only:
    // fresh pageable vectors cost ~10 ms of page faults per call and a staged copy)
    hp::HostPin realign_rb, realign_rq;
    // pinned staging for large pageable inputs of the synchronous entries (read bases, reference, base qualities)
    hp::HostPin pin_reads, pin_ref, pin_quals;
};

namespace hp {
int fail(hp_ctx* ctx, int code, const std::string& msg);
}
