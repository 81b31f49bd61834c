// hp_api.cu -- host side of the C ABI (include/hiphase_b200.h): context, workspaces, H2D/D2H, kernel launches.
// There is deliberately no CPU code path for the algorithms in this file.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <thread>
#include <atomic>

#include <cuda_runtime.h>

#include "../../include/hiphase_b200.h"
#include "hp_device.cuh"
#include "hp_host.h"

namespace hp {

thread_local std::string g_create_error;

bool DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return true;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&ptr, want) != cudaSuccess) { cudaGetLastError(); ptr = nullptr; return false; }
    cap = want;
    return true;
}
void DevBuf::release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }

__global__ void iota_kernel(uint32_t* p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// Processing order: blocks are grouped by score-vector class (max reads per column <= 32 / <= 64 / more) and, inside
// a class, binned largest-first (LPT) by floor(log2(n_cells * min(n_var, 40))) so the longest serial chains start
// first and the tail of the persistent kernel stays short.
__device__ __forceinline__ int order_bin(const BlkMeta& m) {
    const uint64_t cost = (uint64_t)m.n_cells * min(m.n_var, 40u) + m.n_var;
    const int cls = m.max_act <= 32 ? 0 : (m.max_act <= 64 ? 1 : 2);
    return cls * 64 + (63 - (63 - __clzll((long long)(cost | 1ull))));
}
__global__ void order_count_kernel(const BlkMeta* meta, uint32_t n, uint32_t* bins) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&bins[order_bin(meta[i])], 1u);
}
__global__ void order_scan_kernel(uint32_t* bins, uint32_t* class_info) {
    uint32_t run = 0;
    for (int c = 0; c < 3; c++) {
        class_info[4 + c] = run;
        for (int i = 0; i < 64; i++) { uint32_t x = bins[c * 64 + i]; bins[c * 64 + i] = run; run += x; }
        class_info[c] = run - class_info[4 + c];
    }
}
__global__ void order_fill_kernel(const BlkMeta* meta, uint32_t n, uint32_t* bins, uint32_t* order) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[atomicAdd(&bins[order_bin(meta[i])], 1u)] = i;
}

int fail(hp_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}

#define HP_CUDA(ctx, call)                                                                             \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            cudaGetLastError();                                                                        \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? HP_ERR_OUT_OF_MEMORY : HP_ERR_CUDA,     \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                           \
        }                                                                                              \
    } while (0)

static uint32_t sub_capl_for(const hp_params& p) {
    const uint64_t max_visits = (uint64_t)p.min_queue_size / 10 + (uint64_t)p.queue_increment * HP_MAX_SEGMENT;
    const uint64_t live = 1 + 3 * max_visits;
    return (uint32_t)((live + 31) / 32 + 1);
}

// ---- lanes ----------------------------------------------------------------------------------------------------
static int lane_get(hp_ctx* ctx, int idx, AstarLane** out) {
    while ((int)ctx->lanes.size() <= idx) ctx->lanes.push_back(nullptr);
    if (!ctx->lanes[idx]) {
        AstarLane* l = new AstarLane();
        bool ok = cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&l->aux[0], cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&l->aux[1], cudaStreamNonBlocking) == cudaSuccess &&
                  cudaEventCreate(&l->ev0) == cudaSuccess && cudaEventCreate(&l->ev1) == cudaSuccess &&
                  cudaEventCreateWithFlags(&l->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&l->ev_join[0], cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&l->ev_join[1], cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&l->ev_done, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&l->ev_in, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { cudaGetLastError(); delete l; return fail(ctx, HP_ERR_CUDA, "lane stream/event creation failed"); }
        ctx->lanes[idx] = l;
    }
    *out = ctx->lanes[idx];
    return HP_OK;
}

static void lane_destroy(AstarLane* l) {
    if (!l) return;
    for (DevBuf* b : {&l->meta, &l->rmeta, &l->planes, &l->act_off, &l->act_cur, &l->act_idx, &l->col, &l->order, &l->heur,
                      &l->ticket, &l->stage_in, &l->stage_out, &l->dbg})
        b->release();
    l->pin_in.release(); l->pin_out.release();
    for (cudaEvent_t e : {l->ev0, l->ev1, l->ev_fork, l->ev_join[0], l->ev_join[1], l->ev_done, l->ev_in}) if (e) cudaEventDestroy(e);
    for (cudaStream_t st : {l->stream, l->aux[0], l->aux[1]}) if (st) cudaStreamDestroy(st);
    delete l;
}

// Waits until no lane has device work in flight (needed before the slab pool is re-sized).
static int lanes_quiesce(hp_ctx* ctx) {
    for (AstarLane* l : ctx->lanes)
        if (l && l->used) HP_CUDA(ctx, cudaEventSynchronize(l->ev_done));
    return HP_OK;
}

// The slab pool must hold slabs of at least this geometry.  n_min = CTAs that may ask for a slab at the same time.
static int slab_pool_reserve(hp_ctx* ctx, uint32_t qcap, uint32_t hap_words, uint32_t n_min) {
    const uint64_t bytes = astar_slab_bytes(qcap, hap_words, ctx->sub_capl);
    if (ctx->slab_qcap == qcap && ctx->slab_hap_words >= hap_words && ctx->n_slabs >= n_min) return HP_OK;
    int rc = lanes_quiesce(ctx);
    if (rc != HP_OK) return rc;
    const bool same = ctx->slab_qcap == qcap;
    const uint32_t hw = std::max(hap_words, same ? ctx->slab_hap_words : 0u);
    const uint64_t sb = std::max(bytes, astar_slab_bytes(qcap, hw, ctx->sub_capl));
    const uint32_t n = same ? std::max(n_min, ctx->n_slabs) : n_min;
    if (!ctx->slabs.reserve(sb * (uint64_t)n) || !ctx->slab_busy.reserve(4ull * n))
        return fail(ctx, HP_ERR_OUT_OF_MEMORY, "queue slab pool allocation failed");
    HP_CUDA(ctx, cudaMemset(ctx->slab_busy.ptr, 0, 4ull * n));
    ctx->n_slabs = n; ctx->slab_bytes = sb; ctx->slab_hap_words = hw; ctx->slab_qcap = qcap;
    return HP_OK;
}

// Enqueues prep + ordering + the three class kernels of one batch (device pointers) on `stream`, using the workspaces
// of `lane`.  Returns without synchronising.  max_ctas > 0 caps the grid (retry path with huge slabs).
int astar_device(hp_ctx* ctx, AstarLane* L, const hp_block_batch* batch, uint64_t n_vars, uint64_t n_reads, uint64_t n_cells,
                 uint32_t max_block_vars, hp_astar_out* out, cudaStream_t stream, int max_ctas, int busy_lanes) {
    const uint32_t nb = batch->n_blocks;
    if (nb == 0) return HP_OK;
    if (n_cells >= (1ull << 32) || n_reads >= (1ull << 32) || n_cells / 64 + n_reads >= (1ull << 32))
        return fail(ctx, HP_ERR_UNSUPPORTED, "batch too large for 32-bit indices: split it");
    if (max_block_vars == 0) return fail(ctx, HP_ERR_INVALID_INPUT, "max_block_vars must be > 0");

    // the lane's workspaces may still be in use by the previous batch enqueued on another stream
    if (L->used) HP_CUDA(ctx, cudaStreamWaitEvent(stream, L->ev_done, 0));

    const uint64_t n_words = n_cells / 64 + n_reads + 1;
    const uint64_t n_vb = n_vars + nb;
    if (!L->meta.reserve(sizeof(BlkMeta) * (size_t)nb) || !L->rmeta.reserve(sizeof(ReadMeta) * (size_t)(n_reads + 1)) ||
        !L->planes.reserve(8ull * HP_PLANE_STRIDE * n_words) || !L->act_off.reserve(4 * n_vb) ||
        !L->act_cur.reserve(4 * n_vb) || !L->act_idx.reserve(4 * (n_cells + 1)) || !L->col.reserve(4 * (n_cells + 130)) || !L->order.reserve(4ull * nb) ||
        !L->heur.reserve(4 * n_vb) || !L->ticket.reserve(512 + 4 * 192))
        return fail(ctx, HP_ERR_OUT_OF_MEMORY, "workspace allocation failed");

    // Share of the device this launch may occupy.  A batch alone on the device takes every resident warp; with other
    // batches in flight each launch takes capacity x over / (busy lanes + 1) warps, so the launches are co-resident and the
    // serial chain of one launch's slowest block runs beside the bulk of the others (a launch that holds the whole grid
    // would keep the next launch's CTAs queued behind it until its bulk has drained).
    // Throughput regime (other batches in flight and at least four blocks per warp of this launch's share): the dense build of
    // the kernels, 20 warps per SM; a launch whose length is its slowest block's chain keeps the 128-register build (C2 steps
    // of 1000 short blocks over 555 warps lose 20 % with the dense build, C3 steps of 10 000 blocks gain 6 %:
    // profiles/r2m_dense_ab.txt).
    bool dense = busy_lanes > 0 && !out->counters;
    bool force_dense = false;
    if (const char* e = getenv("HP_DBG_DENSE")) {                        // A/B and profiling aid: 0 = never, 2 = also for a launch alone
        dense = dense && atoi(e) != 0;
        if (atoi(e) == 2 && !out->counters) dense = force_dense = true;
    }
    int share_warps = 0, team = 1;
    for (;;) {
        int warps_per_sm = astar_warps_per_sm(dense);
        if (const char* e = getenv("HP_DBG_WARPS_PER_SM")) warps_per_sm = std::max(1, std::min(warps_per_sm, atoi(e)));   // occupancy experiments
        const int resident_warps = ctx->sm_count * warps_per_sm;
        share_warps = resident_warps;
        if (busy_lanes > 0) {
            double over = 1.5;              // measured on C3 (profiles/r2c_sweep.txt): 1.0-1.5 equal within noise, 2.0 slower
            if (const char* e = getenv("HP_DBG_OVERSUB")) over = std::max(0.25, atof(e));
            share_warps = (int)std::min<double>(resident_warps, std::max(1.0, resident_warps * over / (busy_lanes + 1)));
        }
        // team size: use otherwise idle warps for speculative sub-solves (exact, see astar_kernels.cu)
        team = 1;
        // (up to 2x oversubscription of the share still pays: measured on C2, 1000 blocks alone -> team 4)
        while (team * 2 <= astar_max_team() && (uint64_t)nb * team * 2 <= 2ull * (uint64_t)share_warps) team *= 2;
        if (ctx->force_team > 0) team = std::min(ctx->force_team, astar_max_team());
        if (dense && !force_dense && (team > 1 || (uint64_t)nb < 4ull * (uint64_t)share_warps)) { dense = false; continue; }   // latency regime after all
        break;
    }
    int n_ctas = (int)std::min<uint64_t>(nb, (uint64_t)std::max(1, share_warps / team));
    if (max_ctas > 0) n_ctas = std::min(n_ctas, max_ctas);
    const uint32_t hap_words = (max_block_vars + 63) / 64;
    // every CTA that can be resident at once (any mix of launches) finds a free slab
    const uint32_t pool_min = max_ctas > 0 ? (uint32_t)n_ctas : (uint32_t)(ctx->sm_count * ctx->max_ctas_per_sm);
    int rc = slab_pool_reserve(ctx, ctx->qcap, hap_words, pool_min);
    if (rc != HP_OK) return rc;

    HP_CUDA(ctx, cudaMemsetAsync(L->act_off.ptr, 0, 4 * n_vb, stream));
    HP_CUDA(ctx, cudaMemsetAsync(L->act_cur.ptr, 0, 4 * n_vb, stream));
    HP_CUDA(ctx, cudaMemsetAsync(L->ticket.ptr, 0, 512 + 4 * 192, stream));

    PrepArgs pa;
    pa.n_blocks = nb; pa.var_off = batch->var_off; pa.read_off = batch->read_off; pa.read_start = batch->read_start;
    pa.read_end = batch->read_end; pa.cell_off = batch->cell_off; pa.alleles = batch->alleles; pa.quals = batch->quals;
    pa.ignored = batch->ignored;
    pa.meta = (BlkMeta*)L->meta.ptr; pa.rmeta = (ReadMeta*)L->rmeta.ptr; pa.planes = (uint64_t*)L->planes.ptr;
    pa.act_off = (uint32_t*)L->act_off.ptr; pa.act_cur = (uint32_t*)L->act_cur.ptr; pa.act_idx = (uint32_t*)L->act_idx.ptr; pa.col = (uint32_t*)L->col.ptr;
    pa.n_cells_total = n_cells; pa.smem_cap = 0;
    HP_CUDA(ctx, launch_astar_prep(pa, max_block_vars, stream));
    ctx->launches++;

    uint32_t* class_info = (uint32_t*)((uint8_t*)L->ticket.ptr + 256);
    uint32_t* bins = (uint32_t*)((uint8_t*)L->ticket.ptr + 512);
    const int tb = 256, gb = (nb + tb - 1) / tb;
    order_count_kernel<<<gb, tb, 0, stream>>>(pa.meta, nb, bins);
    order_scan_kernel<<<1, 1, 0, stream>>>(bins, class_info);
    order_fill_kernel<<<gb, tb, 0, stream>>>(pa.meta, nb, bins, (uint32_t*)L->order.ptr);
    HP_CUDA(ctx, cudaGetLastError());
    ctx->launches += 3;

    AstarArgs a;
    a.n_blocks = nb;
    a.alleles = batch->alleles; a.quals = batch->quals; a.ignored = batch->ignored; a.is_snv = batch->is_snv;
    a.meta = pa.meta; a.rmeta = pa.rmeta; a.planes = pa.planes; a.act_off = pa.act_off; a.act_idx = pa.act_idx; a.col = pa.col;
    a.order = (uint32_t*)L->order.ptr; a.class_info = class_info;
    a.heur = (uint32_t*)L->heur.ptr; a.ticket = (uint32_t*)L->ticket.ptr;
    a.slabs = (uint8_t*)ctx->slabs.ptr; a.slab_bytes = ctx->slab_bytes; a.qcap = ctx->qcap; a.hap_words = ctx->slab_hap_words;
    a.slab_busy = (uint32_t*)ctx->slab_busy.ptr; a.n_slabs = ctx->n_slabs;
    a.slab_seed = (uint32_t)((ctx->launches * 977u) % ctx->n_slabs);
    a.min_queue_size = ctx->params.min_queue_size; a.queue_increment = ctx->params.queue_increment;
    a.sub_capl = ctx->sub_capl;
    a.out_h1 = out->h1; a.out_h2 = out->h2; a.out_stats = (uint64_t*)out->stats; a.out_status = out->status;
    a.out_heur = out->heuristic; a.out_counters = (uint64_t*)out->counters;
    a.dbg_cycles = nullptr;
    if (out->counters && ctx->want_dbg) {
        if (!L->dbg.reserve(128ull * nb)) return fail(ctx, HP_ERR_OUT_OF_MEMORY, "debug buffer allocation failed");
        a.dbg_cycles = (uint64_t*)L->dbg.ptr; ctx->dbg_blocks = nb;
    }

    // the three score-vector classes run side by side (aux streams fork from / join `stream`)
    HP_CUDA(ctx, cudaEventRecord(L->ev0, stream));
    HP_CUDA(ctx, cudaEventRecord(L->ev_fork, stream));
    cudaStream_t cls_stream[3] = {stream, L->aux[0], L->aux[1]};
    for (int c = 1; c < 3; c++) HP_CUDA(ctx, cudaStreamWaitEvent(cls_stream[c], L->ev_fork, 0));
    HP_CUDA(ctx, launch_astar_solve(a, n_ctas, team, cls_stream, dense));
    for (int c = 1; c < 3; c++) {
        HP_CUDA(ctx, cudaEventRecord(L->ev_join[c - 1], cls_stream[c]));
        HP_CUDA(ctx, cudaStreamWaitEvent(stream, L->ev_join[c - 1], 0));
    }
    HP_CUDA(ctx, cudaEventRecord(L->ev1, stream));
    ctx->launches += 3;
    L->timing_pending = true;
    return HP_OK;
}

// Marks the end of everything enqueued for this lane on `stream` (kernels and result copies).
// true when the driver would have to stage copies to / from this host pointer (ordinary malloc / Vec memory)
bool host_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// Copies the segments {dst, src, bytes} with a few threads (a single memcpy stream is ~10 GB/s; the H2D engine takes 50).
void parallel_copy(const std::vector<CopySeg>& segs) {
    constexpr size_t kPiece = 4u << 20;
    std::vector<CopySeg> pieces;
    size_t total = 0;
    for (const CopySeg& s : segs)
        for (size_t o = 0; o < s.bytes; o += kPiece) { pieces.push_back({s.dst + o, s.src + o, std::min(kPiece, s.bytes - o)}); total += pieces.back().bytes; }
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nt = (unsigned)std::min<size_t>(std::min<size_t>(8, hw), std::max<size_t>(1, total / (8u << 20)));
    std::atomic<size_t> next{0};
    auto work = [&] { for (size_t i; (i = next.fetch_add(1)) < pieces.size();) memcpy(pieces[i].dst, pieces[i].src, pieces[i].bytes); };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (std::thread& t : th) t.join();
}

// Host -> device copy of a large array: pageable sources go through `pin` (filled by parallel_copy) so the copy engine runs
// at its own speed; pinned sources and small arrays are copied directly.  `pin` must stay untouched until the stream is synchronised.
bool upload_large(HostPin& pin, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return true;
    if (bytes < (4u << 20) || !host_pageable(src)) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (!pin.reserve(bytes)) return false;
    parallel_copy({CopySeg{(uint8_t*)pin.ptr, (const uint8_t*)src, bytes}});
    return cudaMemcpyAsync(dst, pin.ptr, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
}

static int lane_mark_done(hp_ctx* ctx, AstarLane* L, cudaStream_t stream) {
    HP_CUDA(ctx, cudaEventRecord(L->ev_done, stream));
    L->used = true;
    return HP_OK;
}

}  // namespace hp

using namespace hp;

extern "C" {

int hp_abi_version(void) { return HP_ABI_VERSION; }

#define HP_STR2(x) #x
#define HP_STR(x) HP_STR2(x)
const char* hp_build_info(void) {
    return "hiphase_b200 ABI " HP_STR(HP_ABI_VERSION) "; nvcc " HP_STR(__CUDACC_VER_MAJOR__) "." HP_STR(__CUDACC_VER_MINOR__) "." HP_STR(__CUDACC_VER_BUILD__)
           "; -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo; compiled " __DATE__ " " __TIME__;
}

void hp_default_params(hp_params* p) {
    p->min_queue_size = 1000; p->queue_increment = 3; p->wfa_prune_distance = 500; p->wfa_max_edit_distance = 500;
}

const char* hp_last_error(const hp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hp_ctx_create(const hp_params* params, int device, hp_ctx** out_ctx) {
    // lanes are streams: give them their own hardware queues (no effect once the process has initialised CUDA)
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    if (!out_ctx) return HP_ERR_INVALID_INPUT;
    *out_ctx = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, HP_ERR_NO_DEVICE, "no CUDA device: hiphase_b200 has no CPU fallback");
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n_dev) return fail(nullptr, HP_ERR_NO_DEVICE, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, HP_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return fail(nullptr, HP_ERR_NO_DEVICE, "kernels are built for sm_100a (Blackwell B200) only");
    hp_params p;
    if (params) p = *params; else hp_default_params(&p);
    // astar_subsolver visits min_queue_size/10 + queue_increment*size nodes (astar_phaser.rs:333); zero visits
    // underflows next_expected-1 in the reference.
    if ((uint64_t)p.min_queue_size / 10 + p.queue_increment == 0)
        return fail(nullptr, HP_ERR_UNSUPPORTED, "min_queue_size/10 + queue_increment must be > 0");
    const uint64_t max_visits = (uint64_t)p.min_queue_size / 10 + (uint64_t)p.queue_increment * HP_MAX_SEGMENT;
    if (4 * max_visits + 2 >= (1u << 20)) return fail(nullptr, HP_ERR_UNSUPPORTED, "sub-solver visit budget exceeds the 20-bit node index");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, HP_ERR_CUDA, "cudaSetDevice failed");
    const uint32_t capl = sub_capl_for(p);
    if (astar_smem_bytes(capl, astar_max_team()) > (size_t)prop.sharedMemPerBlockOptin)
        return fail(nullptr, HP_ERR_UNSUPPORTED, "sub-solver queue does not fit shared memory for these parameters");

    hp_ctx* ctx = new hp_ctx();
    ctx->params = p; ctx->device = device; ctx->sm_count = prop.multiProcessorCount; ctx->sub_capl = capl;
    ctx->max_ctas_per_sm = astar_max_ctas_per_sm(capl);
    uint64_t q = std::max<uint64_t>(16ull * p.min_queue_size + 384, 4096);
    ctx->qcap = (uint32_t)std::min<uint64_t>((q + 31) & ~31ull, 1u << 24);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, HP_ERR_CUDA, "stream/event creation failed");
    }
    *out_ctx = ctx;
    return HP_OK;
}

void hp_ctx_destroy(hp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    hp_comm_destroy(ctx);
    for (AstarLane* l : ctx->lanes) {
        if (l) { if (l->used) cudaEventSynchronize(l->ev_done); delete l->job; lane_destroy(l); }
    }
    for (DevBuf* b : {&ctx->ticket, &ctx->slabs, &ctx->slab_busy, &ctx->stage_in, &ctx->stage_out, &ctx->wfa_ws, &ctx->wfa_in,
                      &ctx->wfa_out, &ctx->wfa_graph, &ctx->comm_send, &ctx->comm_recv, &ctx->ed_scratch})
        b->release();
    ctx->pin_send.release(); ctx->pin_recv.release();
    ctx->realign_rb.release(); ctx->realign_rq.release();
    ctx->pin_reads.release(); ctx->pin_ref.release(); ctx->pin_quals.release();
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

uint64_t hp_launch_count(const hp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int hp_ctx_set_lanes(hp_ctx* ctx, int lanes) {
    if (!ctx || lanes < 1 || lanes > 16) return HP_ERR_INVALID_INPUT;
    ctx->n_lanes = lanes;
    if (ctx->next_lane >= lanes) ctx->next_lane = 0;
    return HP_OK;
}

// Not part of the public header: per-block phase cycles of the last counting run (profiling aid for bench/profiles).
int hp_debug_set_team(hp_ctx* ctx, int team) { if (!ctx) return HP_ERR_INVALID_INPUT; ctx->force_team = team; return HP_OK; }
// bit 0: build the WFA graphs on the host (A/B aid); bit 1: no workspace hint (exercises the regrow path)
int hp_debug_wfa_build_mode(hp_ctx* ctx, int mode) { if (!ctx) return HP_ERR_INVALID_INPUT; ctx->wfa_host_build = (mode & 1) != 0; ctx->wfa_no_hint = (mode & 2) != 0; ctx->wfa_dbg_times = (mode & 4) != 0; return HP_OK; }
// Test aid (not in the public header): scores every read of block 0 of the batch last solved on this context with the device's
// bit-plane scorer (score_planes) against haplotypes h1 / h2 (bit j = allele at haplotype position j) over
// [offset, offset + len): the device counterpart of ReadSegment::score_partial_haplotype (read_segments.rs:177-206).
int hp_debug_score_partial(hp_ctx* ctx, uint64_t h1, uint64_t h2, uint32_t offset, uint32_t len, uint32_t n_reads,
                           uint32_t* s1, uint32_t* s2) {
    if (!ctx || !s1 || !s2 || len > 64 || n_reads == 0) return HP_ERR_INVALID_INPUT;
    AstarLane* L = (size_t)ctx->last_lane < ctx->lanes.size() ? ctx->lanes[ctx->last_lane] : nullptr;
    if (!L || !L->meta.ptr || !L->planes.ptr) return fail(ctx, HP_ERR_INVALID_INPUT, "no solved batch on this context");
    HP_CUDA(ctx, cudaSetDevice(ctx->device));
    HP_CUDA(ctx, cudaStreamSynchronize(L->stream));
    if (!ctx->stage_out.reserve(8ull * n_reads)) return fail(ctx, HP_ERR_OUT_OF_MEMORY, "debug buffer allocation failed");
    AstarArgs a{};
    a.meta = (const BlkMeta*)L->meta.ptr; a.rmeta = (const ReadMeta*)L->rmeta.ptr; a.planes = (const uint64_t*)L->planes.ptr;
    uint32_t* d1 = (uint32_t*)ctx->stage_out.ptr; uint32_t* d2 = d1 + n_reads;
    HP_CUDA(ctx, cudaMemsetAsync(d1, 0xff, 8ull * n_reads, L->stream));
    HP_CUDA(ctx, launch_score_planes_debug(a, h1, h2, (int)offset, (int)len, d1, d2, L->stream));
    HP_CUDA(ctx, cudaMemcpyAsync(s1, d1, 4ull * n_reads, cudaMemcpyDeviceToHost, L->stream));
    HP_CUDA(ctx, cudaMemcpyAsync(s2, d2, 4ull * n_reads, cudaMemcpyDeviceToHost, L->stream));
    HP_CUDA(ctx, cudaStreamSynchronize(L->stream));
    return HP_OK;
}
// Piece filter of the graph-WFA kernel: switch (test / A-B aid) and the number of reads it answered in the last batch call.
int hp_debug_wfa_filter(hp_ctx* ctx, int on) { if (!ctx) return HP_ERR_INVALID_INPUT; ctx->wfa_no_filter = on == 0; return HP_OK; }
uint32_t hp_debug_wfa_filtered(const hp_ctx* ctx) { return ctx ? ctx->wfa_filtered : 0; }
int hp_debug_enable_block_cycles(hp_ctx* ctx, int on) { if (!ctx) return HP_ERR_INVALID_INPUT; ctx->want_dbg = on != 0; return HP_OK; }
int hp_debug_read_block_cycles(hp_ctx* ctx, uint64_t* out, uint32_t n_blocks) {
    if (!ctx || !out || n_blocks > ctx->dbg_blocks) return HP_ERR_INVALID_INPUT;
    AstarLane* L = (size_t)ctx->last_lane < ctx->lanes.size() ? ctx->lanes[ctx->last_lane] : nullptr;
    if (!L || !L->dbg.ptr) return HP_ERR_INVALID_INPUT;
    if (cudaMemcpy(out, L->dbg.ptr, 128ull * n_blocks, cudaMemcpyDeviceToHost) != cudaSuccess) return HP_ERR_CUDA;
    return HP_OK;
}

// Device time of the solver kernels of the last A* call (or of the last WFA / local kernel timed through ctx->ev0/ev1).
float hp_last_kernel_ms(const hp_ctx* cctx) {
    hp_ctx* ctx = const_cast<hp_ctx*>(cctx);
    if (!ctx) return 0.f;
    AstarLane* L = (size_t)ctx->last_lane < ctx->lanes.size() ? ctx->lanes[ctx->last_lane] : nullptr;
    if (L && L->timing_pending) {
        if (cudaEventSynchronize(L->ev1) == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, L->ev0, L->ev1) == cudaSuccess) ctx->last_ms = ms;
        }
        cudaGetLastError();
        L->timing_pending = false;
    }
    if (ctx->timing_pending) {
        if (cudaEventSynchronize(ctx->ev1) == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ms = ms;
        }
        cudaGetLastError();
        ctx->timing_pending = false;
    }
    return ctx->last_ms;
}

// Lane policy: the first lane that holds no job and has no device work in flight (so strictly sequential callers stay on
// lane 0 and its workspaces); when every lane is busy on the device, lanes are taken in rotation (the new batch queues
// behind that lane's previous one).
static int take_lane(hp_ctx* ctx, AstarLane** L, int* idx, int* busy) {
    int pick = -1, n_busy = 0;
    for (int i = 0; i < ctx->n_lanes; i++) {
        AstarLane* l = (size_t)i < ctx->lanes.size() ? ctx->lanes[i] : nullptr;
        bool idle = !l || !l->used;
        if (l && l->used) {
            const cudaError_t e = cudaEventQuery(l->ev_done);
            if (e == cudaSuccess) idle = true; else cudaGetLastError();
        }
        if (!idle) n_busy++;
        if (pick < 0 && idle && !(l && l->job)) pick = i;
    }
    if (pick < 0) {
        for (int k = 0; k < ctx->n_lanes; k++) {
            const int i = (ctx->next_lane + k) % ctx->n_lanes;
            if (!ctx->lanes[i]->job) { pick = i; break; }
        }
        if (pick < 0) return fail(ctx, HP_ERR_INVALID_INPUT, "every lane holds an unfinished job: call hp_astar_wait on the oldest first");
        ctx->next_lane = (pick + 1) % ctx->n_lanes;
        n_busy--;                       // the picked lane's batch precedes this one on the device
    }
    int rc = lane_get(ctx, pick, L);
    if (rc != HP_OK) return rc;
    *idx = pick; ctx->last_lane = pick; *busy = std::max(0, n_busy);
    return HP_OK;
}

int hp_astar_solve_device(hp_ctx* ctx, const hp_block_batch* batch, uint64_t n_vars, uint64_t n_reads, uint64_t n_cells,
                          uint32_t max_block_vars, hp_astar_out* out, void* stream) {
    if (!ctx || !batch || !out) return HP_ERR_INVALID_INPUT;
    HP_CUDA(ctx, cudaSetDevice(ctx->device));
    AstarLane* L; int li;
    int busy = 0;
    int rc = take_lane(ctx, &L, &li, &busy);
    if (rc != HP_OK) return rc;
    if (L->job) return fail(ctx, HP_ERR_INVALID_INPUT, "lane busy with a submitted job: wait for it first");
    // The kernels run on the lane's own streams (one hardware queue each); the caller's stream only orders them: the lane
    // starts after what the caller has enqueued so far, and the caller's stream continues once the lane is done.
    HP_CUDA(ctx, cudaEventRecord(L->ev_in, (cudaStream_t)stream));
    HP_CUDA(ctx, cudaStreamWaitEvent(L->stream, L->ev_in, 0));
    rc = astar_device(ctx, L, batch, n_vars, n_reads, n_cells, max_block_vars, out, L->stream, 0, busy);
    if (rc != HP_OK) return rc;
    rc = lane_mark_done(ctx, L, L->stream);
    if (rc != HP_OK) return rc;
    HP_CUDA(ctx, cudaStreamWaitEvent((cudaStream_t)stream, L->ev_done, 0));
    return HP_OK;
}

static int validate_host_batch(hp_ctx* ctx, const hp_block_batch* b, uint32_t* max_n) {
    if (!b->var_off || !b->read_off || !b->cell_off) return fail(ctx, HP_ERR_INVALID_INPUT, "null offset array");
    uint32_t mx = 0;
    if (b->var_off[0] != 0 || b->read_off[0] != 0 || b->cell_off[0] != 0) return fail(ctx, HP_ERR_INVALID_INPUT, "offset arrays must start at 0");
    for (uint32_t i = 0; i < b->n_blocks; i++) {
        if (b->var_off[i + 1] < b->var_off[i] || b->read_off[i + 1] < b->read_off[i])
            return fail(ctx, HP_ERR_INVALID_INPUT, "offset arrays must be non-decreasing");
        const uint64_t n = b->var_off[i + 1] - b->var_off[i];
        if (n == 0) return fail(ctx, HP_ERR_INVALID_INPUT, "empty phase block (astar_solver is never called with 0 variants, phaser.rs:415-434)");
        if (n > (1u << 24)) return fail(ctx, HP_ERR_UNSUPPORTED, "phase block with more than 2^24 variants");
        mx = std::max<uint32_t>(mx, (uint32_t)n);
    }
    const uint64_t nr = b->read_off[b->n_blocks];
    for (uint64_t r = 0; r < nr; r++)
        if (b->cell_off[r + 1] < b->cell_off[r]) return fail(ctx, HP_ERR_INVALID_INPUT, "cell_off must be non-decreasing");
    *max_n = mx;
    return HP_OK;
}

// Carves `bytes` (256-aligned) out of a staging buffer.
static uint8_t* carve(uint8_t*& p, size_t bytes) { uint8_t* r = p; p += (bytes + 255) & ~(size_t)255; return r; }

// H2D + kernels + D2H of one host batch, all asynchronous on the lane's stream (no synchronisation).  The lane's previous
// job has been waited for (hp_astar_wait / the retry's own synchronisation), so its pinned staging is free.
static int astar_host_enqueue(hp_ctx* ctx, AstarLane* L, const hp_block_batch* b, hp_astar_out* out, uint32_t max_n, int max_ctas, int busy_lanes) {
    const uint32_t nb = b->n_blocks;
    const uint64_t n_vars = b->var_off[nb], n_reads = b->read_off[nb], n_cells = b->cell_off[n_reads];
    cudaStream_t st = L->stream;
    // the staging buffers may still feed the lane's previous batch
    if (L->used) HP_CUDA(ctx, cudaStreamWaitEvent(st, L->ev_done, 0));
    // ---- H2D ----
    const size_t in_bytes = 256 * 12 + 8 * (nb + 1) * 2 + 4 * n_reads * 2 + 8 * (n_reads + 1) + n_cells * 2 + n_vars * 2;
    if (in_bytes > L->stage_in.cap && L->used) HP_CUDA(ctx, cudaEventSynchronize(L->ev_done));   // re-allocation frees the old buffer
    if (!L->stage_in.reserve(in_bytes)) return fail(ctx, HP_ERR_OUT_OF_MEMORY, "input staging allocation failed");
    const bool stage_inputs = host_pageable(n_cells ? (const void*)b->alleles : (const void*)b->var_off);
    if (stage_inputs && !L->pin_in.reserve(in_bytes)) return fail(ctx, HP_ERR_OUT_OF_MEMORY, "pinned input staging allocation failed");
    uint8_t* p = (uint8_t*)L->stage_in.ptr;
    uint8_t* const p0 = p;
    std::vector<CopySeg> segs;
    hp_block_batch d = *b;
#define HP_UP(field, type, count)                                                                                  \
    do {                                                                                                           \
        uint8_t* dst = carve(p, sizeof(type) * (size_t)(count));                                                   \
        if ((count) > 0) {                                                                                         \
            if (stage_inputs) segs.push_back({(uint8_t*)L->pin_in.ptr + (dst - p0), (const uint8_t*)b->field, sizeof(type) * (size_t)(count)}); \
            else HP_CUDA(ctx, cudaMemcpyAsync(dst, b->field, sizeof(type) * (size_t)(count), cudaMemcpyHostToDevice, st)); \
        }                                                                                                          \
        d.field = (const type*)dst;                                                                                \
    } while (0)
    HP_UP(var_off, uint64_t, nb + 1); HP_UP(read_off, uint64_t, nb + 1);
    HP_UP(read_start, uint32_t, n_reads); HP_UP(read_end, uint32_t, n_reads); HP_UP(cell_off, uint64_t, n_reads + 1);
    HP_UP(alleles, uint8_t, n_cells); HP_UP(quals, uint8_t, n_cells);
    HP_UP(ignored, uint8_t, n_vars); HP_UP(is_snv, uint8_t, n_vars);
#undef HP_UP
    if (stage_inputs) {
        parallel_copy(segs);
        HP_CUDA(ctx, cudaMemcpyAsync(p0, L->pin_in.ptr, (size_t)(p - p0), cudaMemcpyHostToDevice, st));
    }
    // ---- device outputs ----
    const size_t out_bytes = 256 * 8 + n_vars * 2 + sizeof(hp_phase_stats) * (size_t)nb + 4ull * nb +
                             (out->heuristic ? 8 * (n_vars + nb) : 0) + (out->counters ? sizeof(hp_astar_counters) * (size_t)nb : 0);
    if (out_bytes > L->stage_out.cap && L->used) HP_CUDA(ctx, cudaEventSynchronize(L->ev_done));
    if (!L->stage_out.reserve(out_bytes)) return fail(ctx, HP_ERR_OUT_OF_MEMORY, "output staging allocation failed");
    uint8_t* q = (uint8_t*)L->stage_out.ptr;
    uint8_t* const q0 = q;
    hp_astar_out dout;
    dout.h1 = carve(q, n_vars); dout.h2 = carve(q, n_vars);
    dout.stats = (hp_phase_stats*)carve(q, sizeof(hp_phase_stats) * (size_t)nb);
    dout.status = (int32_t*)carve(q, 4ull * nb);
    dout.heuristic = out->heuristic ? (uint64_t*)carve(q, 8 * (n_vars + nb)) : nullptr;
    dout.counters = out->counters ? (hp_astar_counters*)carve(q, sizeof(hp_astar_counters) * (size_t)nb) : nullptr;

    int rc = astar_device(ctx, L, &d, n_vars, n_reads, n_cells, max_n, &dout, st, max_ctas, busy_lanes);
    if (rc != HP_OK) return rc;
    // ---- D2H ----
    L->out_staged = host_pageable(out->h1);
    if (L->out_staged) {
        if (!L->pin_out.reserve(out_bytes)) return fail(ctx, HP_ERR_OUT_OF_MEMORY, "pinned output staging allocation failed");
        L->out_off[0] = (size_t)((uint8_t*)dout.h1 - q0); L->out_off[1] = (size_t)((uint8_t*)dout.h2 - q0);
        L->out_off[2] = (size_t)((uint8_t*)dout.stats - q0); L->out_off[3] = (size_t)((uint8_t*)dout.status - q0);
        L->out_off[4] = dout.heuristic ? (size_t)((uint8_t*)dout.heuristic - q0) : 0;
        L->out_off[5] = dout.counters ? (size_t)((uint8_t*)dout.counters - q0) : 0;
        HP_CUDA(ctx, cudaMemcpyAsync(L->pin_out.ptr, q0, (size_t)(q - q0), cudaMemcpyDeviceToHost, st));
        return lane_mark_done(ctx, L, st);
    }
    HP_CUDA(ctx, cudaMemcpyAsync(out->h1, dout.h1, n_vars, cudaMemcpyDeviceToHost, st));
    HP_CUDA(ctx, cudaMemcpyAsync(out->h2, dout.h2, n_vars, cudaMemcpyDeviceToHost, st));
    HP_CUDA(ctx, cudaMemcpyAsync(out->stats, dout.stats, sizeof(hp_phase_stats) * (size_t)nb, cudaMemcpyDeviceToHost, st));
    HP_CUDA(ctx, cudaMemcpyAsync(out->status, dout.status, 4ull * nb, cudaMemcpyDeviceToHost, st));
    if (out->heuristic) HP_CUDA(ctx, cudaMemcpyAsync(out->heuristic, dout.heuristic, 8 * (n_vars + nb), cudaMemcpyDeviceToHost, st));
    if (out->counters) HP_CUDA(ctx, cudaMemcpyAsync(out->counters, dout.counters, sizeof(hp_astar_counters) * (size_t)nb, cudaMemcpyDeviceToHost, st));
    return lane_mark_done(ctx, L, st);
}

// After the lane's stream has been synchronised: results that landed in the pinned staging go to the caller's arrays.
static void astar_host_deliver(AstarLane* L, const hp_block_batch* b, hp_astar_out* out) {
    if (!L->out_staged) return;
    L->out_staged = false;
    const uint32_t nb = b->n_blocks;
    const uint64_t n_vars = b->var_off[nb];
    const uint8_t* s = (const uint8_t*)L->pin_out.ptr;
    std::vector<CopySeg> segs;
    segs.push_back({out->h1, s + L->out_off[0], (size_t)n_vars});
    segs.push_back({out->h2, s + L->out_off[1], (size_t)n_vars});
    segs.push_back({(uint8_t*)out->stats, s + L->out_off[2], sizeof(hp_phase_stats) * (size_t)nb});
    segs.push_back({(uint8_t*)out->status, s + L->out_off[3], 4ull * nb});
    if (out->heuristic) segs.push_back({(uint8_t*)out->heuristic, s + L->out_off[4], 8 * (size_t)(n_vars + nb)});
    if (out->counters) segs.push_back({(uint8_t*)out->counters, s + L->out_off[5], sizeof(hp_astar_counters) * (size_t)nb});
    parallel_copy(segs);
}

// Blocks whose main queue outgrew its slab are re-run (synchronously, on the same lane) with a 4x larger slab and
// fewer resident CTAs.  Called once the first pass of the batch has landed in *out.
static int astar_host_retry(hp_ctx* ctx, AstarLane* L, const hp_block_batch* b, hp_astar_out* out) {
    const uint32_t qcap0 = ctx->qcap;
    for (int attempt = 0; attempt < 4; attempt++) {
        std::vector<uint32_t> redo;
        for (uint32_t i = 0; i < b->n_blocks; i++) if (out->status[i] == HP_BLOCK_QUEUE_OVERFLOW) redo.push_back(i);
        if (redo.empty()) break;
        if (ctx->qcap >= (1u << 24)) break;
        ctx->qcap = std::min<uint32_t>(ctx->qcap * 4, 1u << 24);
        // gather the sub-batch on the host
        std::vector<uint64_t> var_off{0}, read_off{0}, cell_off{0};
        std::vector<uint32_t> rs, re;
        std::vector<uint8_t> al, ql, ig, sn;
        uint32_t sub_max = 0;
        for (uint32_t i : redo) {
            const uint64_t v0 = b->var_off[i], v1 = b->var_off[i + 1];
            sub_max = std::max<uint32_t>(sub_max, (uint32_t)(v1 - v0));
            ig.insert(ig.end(), b->ignored + v0, b->ignored + v1);
            sn.insert(sn.end(), b->is_snv + v0, b->is_snv + v1);
            var_off.push_back(var_off.back() + (v1 - v0));
            for (uint64_t r = b->read_off[i]; r < b->read_off[i + 1]; r++) {
                rs.push_back(b->read_start[r]); re.push_back(b->read_end[r]);
                al.insert(al.end(), b->alleles + b->cell_off[r], b->alleles + b->cell_off[r + 1]);
                ql.insert(ql.end(), b->quals + b->cell_off[r], b->quals + b->cell_off[r + 1]);
                cell_off.push_back(cell_off.back() + (b->cell_off[r + 1] - b->cell_off[r]));
            }
            read_off.push_back(rs.size());
        }
        hp_block_batch sb;
        sb.n_blocks = (uint32_t)redo.size();
        sb.var_off = var_off.data(); sb.read_off = read_off.data(); sb.read_start = rs.data(); sb.read_end = re.data();
        sb.cell_off = cell_off.data(); sb.alleles = al.data(); sb.quals = ql.data(); sb.ignored = ig.data(); sb.is_snv = sn.data();
        const uint64_t nv = var_off.back();
        std::vector<uint8_t> h1(nv), h2(nv);
        std::vector<hp_phase_stats> stats(redo.size());
        std::vector<int32_t> status(redo.size());
        std::vector<uint64_t> heur(out->heuristic ? nv + redo.size() : 0);
        std::vector<hp_astar_counters> ctr(out->counters ? redo.size() : 0);
        hp_astar_out so;
        so.h1 = h1.data(); so.h2 = h2.data(); so.stats = stats.data(); so.status = status.data();
        so.heuristic = out->heuristic ? heur.data() : nullptr; so.counters = out->counters ? ctr.data() : nullptr;
        // keep the slab arena under ~24 GB
        const uint64_t slab = astar_slab_bytes(ctx->qcap, (sub_max + 63) / 64, ctx->sub_capl);
        int max_ctas = (int)std::max<uint64_t>(1, std::min<uint64_t>(ctx->sm_count, (24ull << 30) / std::max<uint64_t>(slab, 1)));
        int rc = astar_host_enqueue(ctx, L, &sb, &so, sub_max, max_ctas, 0);
        if (rc == HP_OK && cudaStreamSynchronize(L->stream) != cudaSuccess) { cudaGetLastError(); rc = fail(ctx, HP_ERR_CUDA, "retry pass failed on the device"); }
        if (rc != HP_OK) { ctx->qcap = qcap0; L->out_staged = false; return rc; }
        astar_host_deliver(L, &sb, &so);
        for (size_t k = 0; k < redo.size(); k++) {
            const uint32_t i = redo[k];
            const uint64_t v0 = b->var_off[i], n = b->var_off[i + 1] - v0;
            memcpy(out->h1 + v0, h1.data() + var_off[k], n); memcpy(out->h2 + v0, h2.data() + var_off[k], n);
            out->stats[i] = stats[k]; out->status[i] = status[k];
            if (out->heuristic) memcpy(out->heuristic + v0 + i, heur.data() + var_off[k] + k, 8 * (n + 1));
            if (out->counters) out->counters[i] = ctr[k];
        }
    }
    ctx->qcap = qcap0;
    return HP_OK;
}

int hp_astar_submit(hp_ctx* ctx, const hp_block_batch* b, hp_astar_out* out, hp_astar_job** job) {
    if (!ctx || !b || !out || !job || !out->h1 || !out->h2 || !out->stats || !out->status) return HP_ERR_INVALID_INPUT;
    *job = nullptr;
    HP_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t max_n = 0;
    if (b->n_blocks != 0) {
        int rc = validate_host_batch(ctx, b, &max_n);
        if (rc != HP_OK) return rc;
    }
    AstarLane* L; int li;
    int busy = 0;
    int rc = take_lane(ctx, &L, &li, &busy);
    if (rc != HP_OK) return rc;
    if (L->job) return fail(ctx, HP_ERR_INVALID_INPUT, "every lane holds an unfinished job: call hp_astar_wait on the oldest first");
    hp_astar_job* j = new hp_astar_job();
    j->lane = li; j->batch = *b; j->out = *out; j->max_n = max_n;
    if (b->n_blocks != 0) {
        rc = astar_host_enqueue(ctx, L, b, out, max_n, 0, busy);
        if (rc != HP_OK) { delete j; return rc; }
    }
    L->job = j;
    *job = j;
    return HP_OK;
}

int hp_astar_poll(hp_ctx* ctx, hp_astar_job* job, int* done) {
    if (!ctx || !job || !done || (size_t)job->lane >= ctx->lanes.size() || ctx->lanes[job->lane]->job != job) return HP_ERR_INVALID_INPUT;
    AstarLane* L = ctx->lanes[job->lane];
    if (job->batch.n_blocks == 0 || !L->used) { *done = 1; return HP_OK; }
    const cudaError_t e = cudaEventQuery(L->ev_done);
    if (e == cudaSuccess) { *done = 1; return HP_OK; }
    if (e == cudaErrorNotReady) { cudaGetLastError(); *done = 0; return HP_OK; }
    cudaGetLastError();
    return fail(ctx, HP_ERR_CUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(e));
}

int hp_astar_wait(hp_ctx* ctx, hp_astar_job* job) {
    if (!ctx || !job || (size_t)job->lane >= ctx->lanes.size() || ctx->lanes[job->lane]->job != job) return HP_ERR_INVALID_INPUT;
    AstarLane* L = ctx->lanes[job->lane];
    int rc = HP_OK;
    if (job->batch.n_blocks != 0) {
        if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamSynchronize(L->stream) != cudaSuccess) {
            cudaGetLastError();
            L->out_staged = false;
            rc = fail(ctx, HP_ERR_CUDA, "device work of the job failed");
        } else {
            ctx->last_lane = job->lane;
            astar_host_deliver(L, &job->batch, &job->out);
            rc = astar_host_retry(ctx, L, &job->batch, &job->out);
        }
    }
    L->job = nullptr;
    delete job;
    return rc;
}

int hp_astar_solve_batch(hp_ctx* ctx, const hp_block_batch* b, hp_astar_out* out) {
    hp_astar_job* job = nullptr;
    int rc = hp_astar_submit(ctx, b, out, &job);
    if (rc != HP_OK) return rc;
    return hp_astar_wait(ctx, job);
}

int hp_astar_solve_one(hp_ctx* ctx, uint32_t n_var, uint32_t n_reads, const uint32_t* read_start, const uint32_t* read_end,
                       const uint64_t* cell_off, const uint8_t* alleles, const uint8_t* quals, const uint8_t* ignored,
                       const uint8_t* is_snv, uint8_t* h1, uint8_t* h2, hp_phase_stats* stats) {
    if (!ctx) return HP_ERR_INVALID_INPUT;
    uint64_t var_off[2] = {0, n_var}, read_off[2] = {0, n_reads};
    hp_block_batch b;
    b.n_blocks = 1; b.var_off = var_off; b.read_off = read_off; b.read_start = read_start; b.read_end = read_end;
    b.cell_off = cell_off; b.alleles = alleles; b.quals = quals; b.ignored = ignored; b.is_snv = is_snv;
    int32_t status = -1;
    hp_astar_out o;
    o.h1 = h1; o.h2 = h2; o.stats = stats; o.status = &status; o.heuristic = nullptr; o.counters = nullptr;
    int rc = hp_astar_solve_batch(ctx, &b, &o);
    if (rc != HP_OK) return rc;
    if (status != HP_BLOCK_OK) {
        // the reference panics here (astar_phaser.rs:439, 529, 631)
        return fail(ctx, HP_ERR_INVALID_INPUT, "block rejected with per-block status " + std::to_string(status));
    }
    return HP_OK;
}

}  // extern "C"
