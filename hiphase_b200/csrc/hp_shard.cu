// hp_shard.cu -- sharding phase blocks over the GPUs of one box and pinned host memory.
//
// Phase blocks share nothing (the reference runs them as independent pool jobs, src/main.rs:385-408), so the multi-GPU
// path has no data-path collective: a cost-sorted deal of whole blocks before the solve (hp_block_costs,
// hp_lpt_partition), one process per GPU, and one result hand-off after it (hp_comm_gather_results: ncclAllGather over
// NVLink / NVSwitch), re-ordered by block index the way OrderedVcfWriter re-orders the reference's worker results
// (src/writers/ordered_vcf_writer.rs:158-170).  NCCL is resolved at run time (dlopen of libnccl.so.2: the copy torch has
// already loaded when the caller is a torch.distributed process), so the library itself has no link-time dependency.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <cuda_runtime.h>

#include "../../include/hiphase_b200.h"
#include "hp_host.h"

using namespace hp;

namespace {

// ---- the few NCCL entry points used, declared here so no NCCL header is needed ----
struct NcclUniqueId { char internal[HP_COMM_ID_BYTES]; };
typedef void* NcclComm;
struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string err;
};
constexpr int kNcclUint8 = 1;   // ncclUint8

NcclApi& nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { api.err = "libnccl.so.2 not found (dlopen)"; return api; }
    api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(h, "ncclCommInitRank");
    api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(h, "ncclAllGather");
    api.CommDestroy = (int (*)(NcclComm))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
    if (!api.ok) api.err = "libnccl.so.2 lacks an expected symbol";
    return api;
}

int nccl_fail(hp_ctx* ctx, const char* what, int rc) {
    return fail(ctx, HP_ERR_CUDA, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "NCCL error"));
}

#define HP_CUDA_S(ctx, call)                                                                          \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            cudaGetLastError();                                                                       \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? HP_ERR_OUT_OF_MEMORY : HP_ERR_CUDA,    \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                          \
        }                                                                                             \
    } while (0)

// device-side all-gather of `bytes` per rank: send (device) -> recv (device, world * bytes)
int allgather_dev(hp_ctx* ctx, const void* send, void* recv, uint64_t bytes) {
    if (ctx->comm_world == 1) {
        if (send != recv) HP_CUDA_S(ctx, cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return HP_OK;
    }
    const int rc = nccl().AllGather(send, recv, (size_t)bytes, kNcclUint8, (NcclComm)ctx->nccl_comm, ctx->stream);
    if (rc != 0) return nccl_fail(ctx, "ncclAllGather", rc);
    return HP_OK;
}

}  // namespace

extern "C" {

// ---- pinned host memory ---------------------------------------------------------------------------------------
int hp_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return HP_ERR_INVALID_INPUT;
    *ptr = nullptr;
    if (cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return HP_ERR_OUT_OF_MEMORY; }
    return HP_OK;
}
int hp_host_free(void* ptr) {
    if (!ptr) return HP_OK;
    if (cudaFreeHost(ptr) != cudaSuccess) { cudaGetLastError(); return HP_ERR_CUDA; }
    return HP_OK;
}
int hp_host_register(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return HP_ERR_INVALID_INPUT;
    if (cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return HP_ERR_CUDA; }
    return HP_OK;
}
int hp_host_unregister(void* ptr) {
    if (!ptr) return HP_ERR_INVALID_INPUT;
    if (cudaHostUnregister(ptr) != cudaSuccess) { cudaGetLastError(); return HP_ERR_CUDA; }
    return HP_OK;
}

// ---- partitioning (host only) -----------------------------------------------------------------------------------
int hp_block_costs(uint64_t n_blocks, const uint32_t* n_var, const uint64_t* n_cells, uint64_t* cost) {
    if ((n_blocks && (!n_var || !n_cells || !cost))) return HP_ERR_INVALID_INPUT;
    for (uint64_t i = 0; i < n_blocks; i++) cost[i] = n_cells[i] * std::min<uint64_t>(n_var[i], HP_MAX_SEGMENT) + n_var[i];
    return HP_OK;
}

int hp_lpt_partition(const uint64_t* cost, uint64_t n_blocks, uint32_t n_shards, uint32_t* shard_of) {
    if (n_shards == 0 || (n_blocks && (!cost || !shard_of))) return HP_ERR_INVALID_INPUT;
    std::vector<uint64_t> order(n_blocks);
    std::iota(order.begin(), order.end(), 0ull);
    std::stable_sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return cost[a] > cost[b]; });
    // least loaded shard first, ties to the lower shard index
    typedef std::pair<uint64_t, uint32_t> Load;
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (uint32_t s = 0; s < n_shards; s++) heap.push(Load(0, s));
    for (uint64_t i : order) {
        Load l = heap.top(); heap.pop();
        shard_of[i] = l.second;
        l.first += cost[i];
        heap.push(l);
    }
    return HP_OK;
}

// ---- NCCL communicator -------------------------------------------------------------------------------------------
int hp_comm_unique_id(uint8_t id[HP_COMM_ID_BYTES]) {
    if (!id) return HP_ERR_INVALID_INPUT;
    if (!nccl().ok) return fail(nullptr, HP_ERR_UNSUPPORTED, nccl().err);
    NcclUniqueId u;
    const int rc = nccl().GetUniqueId(&u);
    if (rc != 0) return nccl_fail(nullptr, "ncclGetUniqueId", rc);
    memcpy(id, u.internal, HP_COMM_ID_BYTES);
    return HP_OK;
}

int hp_comm_init(hp_ctx* ctx, const uint8_t id[HP_COMM_ID_BYTES], int rank, int world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return HP_ERR_INVALID_INPUT;
    hp_comm_destroy(ctx);
    ctx->comm_rank = rank; ctx->comm_world = world;
    if (world == 1) return HP_OK;
    if (!id) return HP_ERR_INVALID_INPUT;
    if (!nccl().ok) return fail(ctx, HP_ERR_UNSUPPORTED, nccl().err);
    HP_CUDA_S(ctx, cudaSetDevice(ctx->device));
    NcclUniqueId u;
    memcpy(u.internal, id, HP_COMM_ID_BYTES);
    NcclComm comm = nullptr;
    const int rc = nccl().CommInitRank(&comm, world, u, rank);
    if (rc != 0) return nccl_fail(ctx, "ncclCommInitRank", rc);
    ctx->nccl_comm = comm;
    return HP_OK;
}

int hp_comm_destroy(hp_ctx* ctx) {
    if (!ctx) return HP_ERR_INVALID_INPUT;
    if (ctx->nccl_comm && nccl().ok) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        nccl().CommDestroy((NcclComm)ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr; ctx->comm_world = 1; ctx->comm_rank = 0;
    return HP_OK;
}

int hp_comm_allgather(hp_ctx* ctx, const void* send, void* recv, uint64_t bytes) {
    if (!ctx || (bytes && (!send || !recv))) return HP_ERR_INVALID_INPUT;
    if (ctx->comm_world > 1 && !ctx->nccl_comm) return fail(ctx, HP_ERR_INVALID_INPUT, "hp_comm_init has not been called");
    if (bytes == 0) return HP_OK;
    HP_CUDA_S(ctx, cudaSetDevice(ctx->device));
    const uint64_t W = (uint64_t)ctx->comm_world;
    if (!ctx->comm_send.reserve(bytes) || !ctx->comm_recv.reserve(bytes * W)) return fail(ctx, HP_ERR_OUT_OF_MEMORY, "gather staging allocation failed");
    HP_CUDA_S(ctx, cudaMemcpyAsync(ctx->comm_send.ptr, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
    int rc = allgather_dev(ctx, ctx->comm_send.ptr, ctx->comm_recv.ptr, bytes);
    if (rc != HP_OK) return rc;
    HP_CUDA_S(ctx, cudaMemcpyAsync(recv, ctx->comm_recv.ptr, bytes * W, cudaMemcpyDeviceToHost, ctx->stream));
    HP_CUDA_S(ctx, cudaStreamSynchronize(ctx->stream));
    return HP_OK;
}

// fixed-stride record of the hand-off: {global block index, status, PhaseStats}
struct GatherRecord {
    uint64_t index;
    int64_t status;
    hp_phase_stats stats;
};

int hp_comm_gather_results(hp_ctx* ctx, uint64_t n_local, const uint64_t* local_ids, const uint64_t* local_var_off,
                           const hp_astar_out* lo, uint64_t n_total, const uint64_t* all_var_off, int root, hp_astar_out* ao) {
    if (!ctx || !all_var_off) return HP_ERR_INVALID_INPUT;
    const bool receive = root < 0 || root == ctx->comm_rank;
    if (receive && (!ao || !ao->h1 || !ao->h2 || !ao->stats || !ao->status)) return HP_ERR_INVALID_INPUT;
    if (n_local && (!local_ids || !local_var_off || !lo || !lo->h1 || !lo->h2 || !lo->stats || !lo->status)) return HP_ERR_INVALID_INPUT;
    if (ctx->comm_world > 1 && !ctx->nccl_comm) return fail(ctx, HP_ERR_INVALID_INPUT, "hp_comm_init has not been called");
    HP_CUDA_S(ctx, cudaSetDevice(ctx->device));
    const uint64_t W = (uint64_t)ctx->comm_world;
    const uint64_t nv_local = n_local ? local_var_off[n_local] : 0;
    // 1. shard sizes
    uint64_t mine[2] = {n_local, nv_local};
    std::vector<uint64_t> sizes(2 * W);
    int rc = hp_comm_allgather(ctx, mine, sizes.data(), sizeof(mine));
    if (rc != HP_OK) return rc;
    uint64_t max_n = 0, max_v = 0, sum_n = 0;
    for (uint64_t r = 0; r < W; r++) { max_n = std::max(max_n, sizes[2 * r]); max_v = std::max(max_v, sizes[2 * r + 1]); sum_n += sizes[2 * r]; }
    if (sum_n != n_total) return fail(ctx, HP_ERR_INVALID_INPUT, "the shards do not add up to n_total blocks");
    if (n_total == 0) return HP_OK;
    // 2. one message per rank: max_n fixed-stride records, then h1 and h2 (each padded to max_v bytes, 16-aligned)
    const uint64_t rec_bytes = sizeof(GatherRecord) * max_n;
    const uint64_t hv = (max_v + 15) & ~15ull;
    const uint64_t msg = rec_bytes + 2 * hv;
    if (!ctx->pin_send.reserve(msg) || (receive && !ctx->pin_recv.reserve(msg * W)) || !ctx->comm_send.reserve(msg) || !ctx->comm_recv.reserve(msg * W))
        return fail(ctx, HP_ERR_OUT_OF_MEMORY, "gather staging allocation failed");
    GatherRecord* send_rec = (GatherRecord*)ctx->pin_send.ptr;
    uint8_t* send_h = (uint8_t*)ctx->pin_send.ptr + rec_bytes;
    for (uint64_t i = 0; i < n_local; i++) {
        if (local_ids[i] >= n_total || all_var_off[local_ids[i] + 1] - all_var_off[local_ids[i]] != local_var_off[i + 1] - local_var_off[i])
            return fail(ctx, HP_ERR_INVALID_INPUT, "local block does not match the global variant offsets");
        send_rec[i].index = local_ids[i]; send_rec[i].status = lo->status[i]; send_rec[i].stats = lo->stats[i];
    }
    for (uint64_t i = n_local; i < max_n; i++) { send_rec[i].index = ~0ull; send_rec[i].status = -1; memset(&send_rec[i].stats, 0, sizeof(hp_phase_stats)); }
    if (nv_local) parallel_copy({CopySeg{send_h, lo->h1, (size_t)nv_local}, CopySeg{send_h + hv, lo->h2, (size_t)nv_local}});
    HP_CUDA_S(ctx, cudaMemcpyAsync(ctx->comm_send.ptr, ctx->pin_send.ptr, msg, cudaMemcpyHostToDevice, ctx->stream));
    rc = allgather_dev(ctx, ctx->comm_send.ptr, ctx->comm_recv.ptr, msg);
    if (rc != HP_OK) return rc;
    if (receive) HP_CUDA_S(ctx, cudaMemcpyAsync(ctx->pin_recv.ptr, ctx->comm_recv.ptr, msg * W, cudaMemcpyDeviceToHost, ctx->stream));
    HP_CUDA_S(ctx, cudaStreamSynchronize(ctx->stream));
    if (!receive) return HP_OK;
    // 3. re-order by global block index: a serial pass checks every record and finds where its haplotype bytes start, then a
    //    few host threads copy the blocks into place (200 000 blocks = three small copies each: ~40 ms on one thread)
    std::vector<uint8_t> seen(n_total, 0);
    struct Src { const GatherRecord* rec; const uint8_t* h1; const uint8_t* h2; };
    std::vector<Src> src;
    src.reserve(n_total);
    for (uint64_t r = 0; r < W; r++) {
        uint64_t v = 0;
        const uint8_t* base = (const uint8_t*)ctx->pin_recv.ptr + r * msg;
        const GatherRecord* rec = (const GatherRecord*)base;
        const uint8_t* h1 = base + rec_bytes;
        const uint8_t* h2 = h1 + hv;
        for (uint64_t i = 0; i < sizes[2 * r]; i++) {
            const GatherRecord& g = rec[i];
            if (g.index >= n_total || seen[g.index]) return fail(ctx, HP_ERR_INVALID_INPUT, "block index gathered twice or out of range");
            seen[g.index] = 1;
            const uint64_t n = all_var_off[g.index + 1] - all_var_off[g.index];
            if (v + n > sizes[2 * r + 1]) return fail(ctx, HP_ERR_INVALID_INPUT, "gathered haplotype bytes do not match the variant offsets");
            src.push_back({&g, h1 + v, h2 + v});
            v += n;
        }
    }
    auto place = [&](size_t lo_i, size_t hi_i) {
        for (size_t k = lo_i; k < hi_i; k++) {
            const GatherRecord& g = *src[k].rec;
            const uint64_t o = all_var_off[g.index], n = all_var_off[g.index + 1] - o;
            memcpy(ao->h1 + o, src[k].h1, n); memcpy(ao->h2 + o, src[k].h2, n);
            ao->stats[g.index] = g.stats; ao->status[g.index] = (int32_t)g.status;
        }
    };
    const size_t nt = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(8, std::max(1u, std::thread::hardware_concurrency())), src.size() / 8192));
    std::vector<std::thread> th;
    for (size_t t = 1; t < nt; t++) th.emplace_back(place, src.size() * t / nt, src.size() * (t + 1) / nt);
    place(0, src.size() / nt);
    for (std::thread& t : th) t.join();
    return HP_OK;
}

}  // extern "C"
