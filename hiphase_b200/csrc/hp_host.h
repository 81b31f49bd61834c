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

// kernel launchers (astar_kernels.cu)
size_t astar_smem_bytes(uint32_t sub_capl, int team);
int astar_max_team();
int astar_warps_per_sm();
uint64_t astar_slab_bytes(uint32_t qcap, uint32_t hap_words, uint32_t sub_capl);
cudaError_t launch_astar_prep(const PrepArgs& pa, cudaStream_t stream);
cudaError_t launch_astar_solve(const AstarArgs& a, int n_ctas, int team, cudaStream_t stream);

}  // namespace hp

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
    // A* workspaces
    bool want_dbg = false;
    uint32_t dbg_blocks = 0;
    hp::DevBuf dbg;
    hp::DevBuf meta, rmeta, planes, act_off, act_cur, act_idx, col, order, heur, ticket, slabs, stage_in, stage_out;
    // WFA workspaces
    hp::DevBuf wfa_ws, wfa_in, wfa_out, wfa_graph;
    bool wfa_no_hint = false;               // test aid: start with an unsized graph workspace (exercises the regrow path)
    bool wfa_host_build = false;            // debug / A-B aid: build the graphs on the host instead of on the device
    uint32_t wfa_table_cap = 1u << 15;      // (node, diagonal) hash slots per warp; grown 8x on overflow
    std::vector<int32_t> wfa_h_status;
    std::vector<uint32_t> wfa_h_score, wfa_h_nodes;
    std::vector<uint8_t> wfa_h_alleles, wfa_h_quals;
    std::vector<uint64_t> wfa_h_trav, wfa_h_ctr;
};
