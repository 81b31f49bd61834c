// astar_kernels.cu -- sm_100a kernels for the A* phasing hot path.
//
// Replaces (results bit-identical): calculate_astar_heuristic / astar_subsolver / astar_solver
// (src/astar_phaser.rs:246-633) and their inner loop ReadSegment::score_partial_haplotype
// (src/data_types/read_segments.rs:177-206).
//
// Design:
//   * astar_prep_kernel  -- one CTA per phase block: validates the block, bit-packs every read into 64-variant
//     word records {allele bit, non-binary bit, 8 quality bit-planes of qual/gcd}, and builds the per-variant
//     active-read lists (the reference's interval-tree stabbing query, astar_phaser.rs:92, precomputed).
//   * astar_solve_kernel -- persistent, ONE WARP PER PHASE BLOCK pulled from an atomic ticket.  The reference's
//     pop sequence is a strict total order on (cost, -hets, node_index), so the warp replays exactly that order;
//     parallelism is inside one expansion: lanes = active reads, each scoring both parent haplotypes with
//     AND/XOR/POPC over the bit planes (mismatch cost = gcd * sum_b 2^b popc(mask & plane_b)), the four children
//     derived from the parent scores plus the new column, then redux.sync adds.  The sub-solver queue
//     (<= 1 + 3*(100+3*40) nodes, 40-bit haplotypes) lives in shared memory as 32 lane-owned stripes with the
//     stripe minimum cached in registers (pop = 2 redux.sync + ballot); the main queue lives in a per-warp slab
//     in HBM/L2 with full-length haplotype records.
#include <algorithm>

#include "hp_device.cuh"
#include "../../include/hiphase_b200.h"

namespace hp {

// =============================================================================================================
// prep kernel
// =============================================================================================================

__device__ __forceinline__ uint32_t gcd_u32(uint32_t a, uint32_t b) {
    while (b) { uint32_t t = a % b; a = b; b = t; }
    return a;
}

constexpr int kPrepThreads = 128;            // = reads per tile (one read per thread per tile)
constexpr uint32_t kStageBytes = 4096;       // per array (alleles / quals) and stage
constexpr uint32_t kPrepStages = 2;

// ---- bulk-async (TMA) staging primitives: cp.async.bulk global -> shared, completion on an mbarrier ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One CTA per phase block.  The block's cells (alleles / quals, the reference's u8 layout) are streamed twice through a
// two-stage shared-memory ring by the bulk-async copy engine (cp.async.bulk + mbarrier, 128-read tiles, 16-byte aligned
// supersets of the tile's cell range), the per-variant coverage counters and active-list cursors live in shared memory
// (blocks with more variants than the launch's cap fall back to the global arrays), and only the products leave the SM:
// read metadata, bit-plane word records, active lists and column records.
__global__ void __launch_bounds__(kPrepThreads) astar_prep_kernel(PrepArgs a) {
    extern __shared__ __align__(16) uint8_t prep_dyn[];
    const uint32_t b = blockIdx.x;
    if (b >= a.n_blocks) return;
    const uint64_t v0 = a.var_off[b], v1 = a.var_off[b + 1];
    const uint64_t r0 = a.read_off[b], r1 = a.read_off[b + 1];
    const uint64_t c0 = a.cell_off[r0], c1 = a.cell_off[r1];
    const uint32_t N = (uint32_t)(v1 - v0);
    const uint32_t R = (uint32_t)(r1 - r0);
    const int tid = threadIdx.x;

    __shared__ uint32_t s_presence[8];
    __shared__ unsigned long long s_qsum;
    __shared__ int s_status;
    __shared__ uint32_t s_maxspan;
    __shared__ uint8_t s_div[256];
    __shared__ uint32_t s_partial[kPrepThreads];
    __shared__ uint32_t s_g, s_planes, s_maxact;
    __shared__ __align__(8) uint64_t s_bar[kPrepStages];

    // dynamic shared memory: [stages x {alleles, quals} x kStageBytes] [cnt: (cap + 1) u32] [cur: (cap + 1) u16, packed]
    uint8_t* const stage = prep_dyn;
    uint32_t* const cnt_s = (uint32_t*)(prep_dyn + kPrepStages * 2 * kStageBytes);
    uint32_t* const cur_s = cnt_s + (a.smem_cap + 1u);
    const bool in_smem = N <= a.smem_cap;
    // bulk copies need 16-byte aligned global addresses: true for the library's staging buffers, checked for callers' pointers
    const bool aligned = ((((uintptr_t)a.alleles) | ((uintptr_t)a.quals)) & 15u) == 0;
    const uint64_t bulk_limit = a.n_cells_total & ~15ull;            // no bulk copy may read past the arrays

    if (tid < 8) s_presence[tid] = 0;
    if (tid == 0) {
        s_qsum = 0; s_status = HP_BLOCK_OK; s_maxspan = 0; s_maxact = 0;
        for (uint32_t k = 0; k < kPrepStages; k++) mbar_init(&s_bar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t* cnt = in_smem ? cnt_s : a.act_off + v0 + b;            // N + 1 entries
    uint32_t* cur = a.act_cur + v0 + b;                              // (global fallback only)
    if (in_smem) {
        for (uint32_t i = tid; i <= N; i += kPrepThreads) cnt_s[i] = 0;
        for (uint32_t i = tid; i <= N / 2; i += kPrepThreads) cur_s[i] = 0;
    }
    __syncthreads();

    // ---- tile pipeline ----------------------------------------------------------------------------------------------
    const uint32_t n_tiles = (R + kPrepThreads - 1) / kPrepThreads;
    uint32_t phase_bits = 0;                                         // bit s = parity the next wait on stage s expects
    struct Tile { uint64_t as; uint32_t bulk, span; bool staged; };
    auto tile_of = [&](uint32_t t) {
        Tile ti;
        const uint64_t ra = r0 + (uint64_t)t * kPrepThreads, rb = min(ra + (uint64_t)kPrepThreads, r1);
        const uint64_t cs = a.cell_off[ra], ce = a.cell_off[rb];
        ti.as = cs & ~15ull;
        const uint64_t ae = (ce + 15ull) & ~15ull;
        ti.span = (uint32_t)min(ae - ti.as, (uint64_t)0xffffffffu);
        ti.staged = aligned && ce >= cs && (ae - ti.as) <= kStageBytes;
        const uint64_t be = min(ae, bulk_limit);
        ti.bulk = (ti.staged && be > ti.as) ? (uint32_t)(be - ti.as) : 0u;
        return ti;
    };
    auto issue = [&](uint32_t t) {                                   // every thread calls it; thread 0 drives the copy engine
        const Tile ti = tile_of(t);
        if (!ti.staged) return;
        const uint32_t s = t % kPrepStages;
        uint8_t* sa = stage + (size_t)s * 2 * kStageBytes;
        uint8_t* sq = sa + kStageBytes;
        if (tid == 0 && ti.bulk) {
            fence_proxy_async();                                     // the stage was read through the generic proxy before
            mbar_expect_tx(&s_bar[s], 2u * ti.bulk);
            bulk_g2s(sa, a.alleles + ti.as, ti.bulk, &s_bar[s]);
            bulk_g2s(sq, a.quals + ti.as, ti.bulk, &s_bar[s]);
        }
        // the last (< 16 byte) piece of the arrays cannot be bulk-copied: plain loads
        const uint64_t tail0 = ti.as + ti.bulk, tail1 = min(ti.as + (uint64_t)ti.span, a.n_cells_total);
        if (tail0 + (uint64_t)tid < tail1 && tid < 16) {
            sa[ti.bulk + tid] = a.alleles[tail0 + tid];
            sq[ti.bulk + tid] = a.quals[tail0 + tid];
        }
    };
    auto wait_tile = [&](uint32_t t, const Tile& ti) {
        const uint32_t s = t % kPrepStages;
        if (ti.bulk) {
            while (!mbar_try_wait(&s_bar[s], (phase_bits >> s) & 1u)) {}
            phase_bits ^= 1u << s;
        }
        __syncthreads();                                             // the plain-load tail, and a common starting line
    };

    // ---- pass 1: validate, count coverage, collect the set of quality values ----
    {
        uint32_t pres[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned long long qsum = 0;
        int status = HP_BLOCK_OK;
        uint32_t maxspan = 0;
        if (N == 0) status = HP_BLOCK_ASSERT;
        if (n_tiles) issue(0);
        for (uint32_t t = 0; t < n_tiles; t++) {
            if (t + 1 < n_tiles) issue(t + 1);
            const Tile ti = tile_of(t);
            wait_tile(t, ti);
            const uint64_t r = r0 + (uint64_t)t * kPrepThreads + tid;
            if (r < r1) {
                const uint32_t s = a.read_start[r], e = a.read_end[r];
                const uint64_t c = a.cell_off[r];
                if (e < s || e > N || a.cell_off[r + 1] - c != (uint64_t)(e - s) ||
                    (ti.staged && (c < ti.as || c + (e - s) > ti.as + ti.span))) status = HP_BLOCK_ASSERT;
                else {
                    maxspan = max(maxspan, e - s);
                    ReadMeta rm;
                    rm.start = s; rm.end = e; rm.word_idx = (uint32_t)(c / 64 + r); rm.cell_rel = (uint32_t)(c - c0);
                    a.rmeta[r] = rm;
                    const uint8_t* sa = stage + (size_t)(t % kPrepStages) * 2 * kStageBytes;
                    const uint8_t* pa = ti.staged ? sa + (c - ti.as) : a.alleles + c;
                    const uint8_t* pq = ti.staged ? sa + kStageBytes + (c - ti.as) : a.quals + c;
                    for (uint32_t i = 0; i < e - s; i++) {
                        const uint8_t al = pa[i];
                        const uint8_t q = pq[i];
                        const uint32_t p = s + i;
                        if (al > 3) status = HP_BLOCK_ASSERT;
                        if (a.ignored[v0 + p]) {
                            if (al != HP_ALLELE_NOOVERLAP && status == HP_BLOCK_OK) status = HP_BLOCK_IGNORED_NOT_NOOVERLAP;
                        } else {
                            pres[q >> 5] |= 1u << (q & 31);
                            qsum += q;
                        }
                        atomicAdd(&cnt[p], 1u);
                    }
                }
            }
            __syncthreads();                                         // stage t % 2 is free again
        }
        for (int k = 0; k < 8; k++) if (pres[k]) atomicOr(&s_presence[k], pres[k]);
        if (qsum) atomicAdd(&s_qsum, qsum);
        if (status != HP_BLOCK_OK) atomicMax(&s_status, status);
        if (maxspan) atomicMax(&s_maxspan, maxspan);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t g = 0, mx = 0;
        for (uint32_t q = 1; q < 256; q++)
            if (s_presence[q >> 5] >> (q & 31) & 1u) { g = gcd_u32(g, q); mx = q; }
        if (g == 0) g = 1;
        uint32_t np = 0;
        for (uint32_t t = mx / g; t; t >>= 1) np++;
        s_g = g; s_planes = np;
        if (s_qsum >= (1ull << 31) && s_status == HP_BLOCK_OK) s_status = HP_BLOCK_COST_OVERFLOW;
    }
    __syncthreads();
    const uint32_t g = s_g;
    for (int q = tid; q < 256; q += kPrepThreads) s_div[q] = (uint8_t)(q / g);

    // ---- exclusive scan of the coverage counts -> act_off ----
    const uint32_t chunk = (N + 1 + kPrepThreads - 1) / kPrepThreads;
    const uint32_t lo = min(N + 1, tid * chunk), hi = min(N + 1, lo + chunk);
    {
        uint32_t sum = 0, mx = 0;
        for (uint32_t i = lo; i < hi; i++) { const uint32_t x = in_smem ? cnt_s[i] : __ldcg(&cnt[i]); sum += x; mx = max(mx, x); }
        s_partial[tid] = sum;
        if (mx) atomicMax(&s_maxact, mx);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int t = 0; t < kPrepThreads; t++) { uint32_t x = s_partial[t]; s_partial[t] = run; run += x; }
    }
    __syncthreads();
    {
        uint32_t run = s_partial[tid];
        if (in_smem) for (uint32_t i = lo; i < hi; i++) { uint32_t x = cnt_s[i]; cnt_s[i] = run; run += x; }
        else for (uint32_t i = lo; i < hi; i++) { uint32_t x = __ldcg(&cnt[i]); __stcg(&cnt[i], run); run += x; }
    }
    __syncthreads();
    if (in_smem) {
        uint32_t* aoff = a.act_off + v0 + b;
        for (uint32_t i = tid; i <= N; i += kPrepThreads) aoff[i] = cnt_s[i];
    }

    if (tid == 0 && s_maxact >= 0xffffu && s_status == HP_BLOCK_OK) s_status = HP_BLOCK_TOO_DENSE;
    __syncthreads();
    // ---- pass 2: bit planes, active lists and column records ----
    if (s_status == HP_BLOCK_OK) {
        if (n_tiles) issue(0);
        for (uint32_t t = 0; t < n_tiles; t++) {
            if (t + 1 < n_tiles) issue(t + 1);
            const Tile ti = tile_of(t);
            wait_tile(t, ti);
            const uint64_t r = r0 + (uint64_t)t * kPrepThreads + tid;
            if (r < r1) {
                const uint32_t s = a.read_start[r], e = a.read_end[r];
                const uint64_t c = a.cell_off[r];
                const uint8_t* sa = stage + (size_t)(t % kPrepStages) * 2 * kStageBytes;
                const uint8_t* pa = ti.staged ? sa + (c - ti.as) : a.alleles + c;
                const uint8_t* pq = ti.staged ? sa + kStageBytes + (c - ti.as) : a.quals + c;
                uint64_t* rec = a.planes + (uint64_t)(c / 64 + r) * HP_PLANE_STRIDE;
                uint64_t w[HP_PLANE_STRIDE];
#pragma unroll
                for (int k = 0; k < (int)HP_PLANE_STRIDE; k++) w[k] = 0;
                uint32_t prev_slot = 0xffffu;        // slot of this read in the previous column's active list
                for (uint32_t i = 0; i < e - s; i++) {
                    const uint8_t al = pa[i];
                    const uint8_t qraw = pq[i];
                    const uint32_t p = s + i;
                    // quality of an ignored column can never be charged (the haplotype is Ambiguous there): drop it
                    const uint32_t q = a.ignored[v0 + p] ? 0u : (uint32_t)s_div[qraw];
                    const uint64_t bit = 1ull << (i & 63);
                    if (al & 1) w[0] |= bit;          // allele bit (meaningful for 0/1; 3 sets it too but nb masks it)
                    if (al >= 2) w[1] |= bit;         // non-binary: mismatches both 0 and 1
#pragma unroll
                    for (int k = 0; k < 8; k++) if (q >> k & 1u) w[2 + k] |= bit;
                    uint32_t slot, off;
                    if (in_smem) {
                        const uint32_t sh = (p & 1u) * 16u;
                        slot = (atomicAdd(&cur_s[p >> 1], 1u << sh) >> sh) & 0xffffu;     // maxact < 0xffff: no carry across halves
                        off = cnt_s[p];
                    } else {
                        slot = atomicAdd(&cur[p], 1u);
                        off = __ldcg(&cnt[p]);
                    }
                    const uint64_t at = c0 + off + slot;
                    a.act_idx[at] = (uint32_t)(r - r0);
                    // column record: qual | allele<<8 | ends<<10 | carry<<16 (carry = slot in column p-1, 0xffff = read starts here)
                    a.col[at] = (uint32_t)qraw | ((uint32_t)al << 8) | ((i + 1 == e - s) ? (1u << 10) : 0u) | (prev_slot << 16);
                    prev_slot = slot;
                    if ((i & 63) == 63 || i + 1 == e - s) {
#pragma unroll
                        for (int k = 0; k < (int)HP_PLANE_STRIDE; k++) { rec[k] = w[k]; w[k] = 0; }
                        rec += HP_PLANE_STRIDE;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        BlkMeta m;
        m.var_base = v0; m.read_base = r0; m.cell_base = c0;
        m.n_var = N; m.n_reads = R; m.n_cells = (uint32_t)(c1 - c0);
        m.qgcd = g; m.n_planes = s_planes; m.status = s_status; m.max_span = s_maxspan; m.max_act = s_maxact;
        a.meta[b] = m;
    }
}

// =============================================================================================================
// solver kernel
// =============================================================================================================

// ---- warp reductions -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t wsum(uint32_t x) { return __reduce_add_sync(HP_FULL_MASK, x); }
__device__ __forceinline__ uint32_t wmin(uint32_t x) { return __reduce_min_sync(HP_FULL_MASK, x); }
__device__ __forceinline__ uint64_t wmin64(uint64_t x) {
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    const uint32_t mhi = wmin(hi);
    const uint32_t mlo = wmin(hi == mhi ? lo : 0xffffffffu);
    return ((uint64_t)mhi << 32) | mlo;
}

// ---- bit-plane scoring (the from-scratch path) --------------------------------------------------------------
// Scores one read against both haplotypes of a node over read coordinates [max(0,o), o+L):
//   o     = read coordinate of haplotype position 0   (problem_offset - read.start, may be negative)
//   L     = haplotype length
//   HapFn = functor(which, bitpos) -> 64 haplotype bits starting at haplotype position bitpos (may be < 0)
// This is ReadSegment::score_partial_haplotype (read_segments.rs:177-206) for h1 and h2 at once:
// cost = qgcd * sum_b 2^b * popc(range & ((hap ^ allele_bits) | non_binary) & quality_plane_b).
template <class HapFn>
__device__ __forceinline__ void score_planes(const AstarArgs& a, const BlkMeta& m, const ReadMeta rm, int o, int L,
                                             HapFn hap, uint32_t& s1, uint32_t& s2) {
    s1 = 0; s2 = 0;
    const int c_lo = max(0, o), c_hi = o + L;
    if (c_hi <= c_lo) return;
    const int k_lo = c_lo >> 6, k_hi = (c_hi - 1) >> 6;
    const uint64_t* rec = a.planes + (uint64_t)rm.word_idx * HP_PLANE_STRIDE;
    for (int k = k_lo; k <= k_hi; k++) {
        const uint64_t* w = rec + (uint64_t)k * HP_PLANE_STRIDE;
        const uint64_t range = bit_range(max(0, c_lo - 64 * k), min(64, c_hi - 64 * k));
        const uint64_t ab = __ldg(w + 0), nb = __ldg(w + 1);
        const int i0 = 64 * k - o;                    // haplotype position of this word's bit 0
        const uint64_t m1 = range & ((hap(0, i0) ^ ab) | nb);
        const uint64_t m2 = range & ((hap(1, i0) ^ ab) | nb);
        for (uint32_t bpl = 0; bpl < m.n_planes; bpl++) {
            const uint64_t q = __ldg(w + 2 + bpl);
            s1 += (uint32_t)__popcll(m1 & q) << bpl;
            s2 += (uint32_t)__popcll(m2 & q) << bpl;
        }
    }
    s1 *= m.qgcd; s2 *= m.qgcd;
}

// ---- incremental expansion ------------------------------------------------------------------------------------
// Candidate children of one expansion, fixed slots (creation order of astar_phaser.rs:367-372 / 535-540):
//   0 = (0|1), 1 = (1|0), 2 = (0/0) [also the single (2,2) child of an ignored variant], 3 = (1/1)
// For the read in active-list slot j of column p the parent holds (s1, s2) = its cost against h1 / h2 so far;
// the four children only need A0 = s1+q0, A1 = s1+q1, B0 = s2+q0, B1 = s2+q1 where q0/q1 is the column's
// quality if the read's allele mismatches 0/1.  Those four values per slot stay in registers ("cache"): when a
// child of this expansion is popped next, its (s1, s2) vector is one shuffle away (slot -> slot of the previous
// column through the column record's carry index).
template <int K>
struct ExpCache {
    uint32_t a0[K > 0 ? K : 1], a1[K > 0 ? K : 1], b0[K > 0 ? K : 1], b1[K > 0 ? K : 1], w[K > 0 ? K : 1];
    uint32_t first_idx;    // node index of the first child of the cached expansion (0xffffffff = empty)
    uint32_t present;      // 4-bit mask of candidates that were created
};

enum VecSrc { SRC_ROOT = 0, SRC_CACHE = 1, SRC_PLANES = 2 };

template <int K, bool kCount, class HapFn>
__device__ __forceinline__ void expand(const AstarArgs& a, const BlkMeta& m, uint32_t lane, const uint32_t* colp,
                                       const uint32_t* aidxp, const ReadMeta* rmeta, uint32_t A, uint32_t p, int off,
                                       int L, bool bad_col, bool ident, int src, uint32_t x1, uint32_t x2, HapFn hap,
                                       ExpCache<K>& cache, uint32_t (&tot)[4], uint32_t (&fro)[4], uint64_t& cells) {
    uint32_t at[4] = {0, 0, 0, 0}, af[4] = {0, 0, 0, 0};
    if constexpr (K == 0) {
        // generic path: any number of active reads, always from the bit planes
        for (uint32_t j = lane; j < A; j += 32) {
            const uint32_t c = __ldg(colp + j);
            const ReadMeta rm = rmeta[__ldg(aidxp + j)];
            uint32_t s1 = 0, s2 = 0;
            if (src != SRC_ROOT) score_planes(a, m, rm, off - (int)rm.start, L, hap, s1, s2);
            const uint32_t q = c & 0xffu, al = (c >> 8) & 3u;
            const bool ends = (c >> 10) & 1u;
            const uint32_t q0 = (!bad_col && al != 0u) ? q : 0u, q1 = (!bad_col && al != 1u) ? q : 0u;
            const uint32_t A0 = s1 + q0, A1 = s1 + q1, B0 = s2 + q0, B1 = s2 + q1;
            const uint32_t c01 = min(A0, B1), c10 = min(A1, B0), c00 = min(A0, B0), c11 = min(A1, B1);
            at[0] += c01; at[1] += c10; at[2] += c00; at[3] += c11;
            if (ends) { af[0] += c01; af[1] += c10; af[2] += c00; af[3] += c11; }
            if (kCount) cells += (uint64_t)(p + 1 - (uint32_t)max((int)rm.start, off));
        }
    } else {
        const ExpCache<K> old = cache;      // the gathers read the previous expansion while this one is written
#pragma unroll
        for (int k = 0; k < K; k++) {
            const uint32_t j = lane + 32u * k;
            const bool valid = j < A;
            const uint32_t c = valid ? __ldg(colp + j) : 0xffff0000u;
            uint32_t s1 = 0, s2 = 0, wv = 0;
            if (src == SRC_CACHE) {
                const uint32_t carry = c >> 16;
                const uint32_t sl = carry & 31u;
                uint32_t v1 = __shfl_sync(HP_FULL_MASK, x1 ? old.a1[0] : old.a0[0], sl);
                uint32_t v2 = __shfl_sync(HP_FULL_MASK, x2 ? old.b1[0] : old.b0[0], sl);
                uint32_t vw = kCount ? __shfl_sync(HP_FULL_MASK, old.w[0], sl) : 0u;
                if constexpr (K > 1) {
#pragma unroll
                    for (int kk = 1; kk < K; kk++) {
                        const uint32_t u1 = __shfl_sync(HP_FULL_MASK, x1 ? old.a1[kk] : old.a0[kk], sl);
                        const uint32_t u2 = __shfl_sync(HP_FULL_MASK, x2 ? old.b1[kk] : old.b0[kk], sl);
                        const uint32_t uw = kCount ? __shfl_sync(HP_FULL_MASK, old.w[kk], sl) : 0u;
                        if ((carry >> 5) == (uint32_t)kk) { v1 = u1; v2 = u2; vw = uw; }
                    }
                }
                if (carry != 0xffffu) { s1 = v1; s2 = v2; wv = vw; }
            } else if (src == SRC_PLANES) {
                if (valid) {
                    const ReadMeta rm = rmeta[__ldg(aidxp + j)];
                    score_planes(a, m, rm, off - (int)rm.start, L, hap, s1, s2);
                    wv = p - (uint32_t)max((int)rm.start, off);
                }
            }
            const uint32_t q = c & 0xffu, al = (c >> 8) & 3u;
            const bool ends = (c >> 10) & 1u;
            const uint32_t q0 = (!bad_col && al != 0u) ? q : 0u, q1 = (!bad_col && al != 1u) ? q : 0u;
            const uint32_t A0 = s1 + q0, A1 = s1 + q1, B0 = s2 + q0, B1 = s2 + q1;
            const uint32_t c01 = min(A0, B1), c10 = min(A1, B0), c00 = min(A0, B0), c11 = min(A1, B1);
            at[0] += c01; at[1] += c10; at[2] += c00; at[3] += c11;
            if (ends) { af[0] += c01; af[1] += c10; af[2] += c00; af[3] += c11; }
            cache.a0[k] = A0; cache.a1[k] = A1; cache.b0[k] = B0; cache.b1[k] = B1;
            if (kCount) { cache.w[k] = wv + 1; if (valid) cells += wv + 1; }
        }
    }
    // only the candidates that will be created are reduced
    tot[2] = wsum(at[2]); fro[2] = wsum(af[2]);
    if (!bad_col) {
        tot[0] = wsum(at[0]); fro[0] = wsum(af[0]);
        tot[3] = wsum(at[3]); fro[3] = wsum(af[3]);
        if (!ident) { tot[1] = wsum(at[1]); fro[1] = wsum(af[1]); }
    }
}

// present-candidate mask of an expansion
__device__ __forceinline__ uint32_t present_mask(bool bad_col, bool ident) { return bad_col ? 0x4u : (ident ? 0xdu : 0xfu); }
// candidate slot of the d-th created child of an expansion with that mask
__device__ __forceinline__ uint32_t slot_of_ordinal(uint32_t present, uint32_t d) {
    return present == 0xfu ? d : (present == 0xdu ? d + (d > 0u ? 1u : 0u) : 2u);
}

// ---- sub-solver key: [total:32][63-hets:6][node_index:20][len:6] -------------------------------------------
__device__ __forceinline__ uint64_t mk64(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; }

// One sub-solver queue entry (32 B = two 128-bit shared-memory transactions).
struct __align__(16) SubEntry {
    uint64_t key, h1, h2;
    uint32_t frozen, tag;
};

#ifndef HP_DEAD_POOL
#define HP_DEAD_POOL 1
#endif
constexpr bool kDeadPool = HP_DEAD_POOL != 0;
#ifndef HP_SWEEP_MIN_DEAD
#define HP_SWEEP_MIN_DEAD 256
#endif
constexpr uint32_t kSweepMinDead = HP_SWEEP_MIN_DEAD;     // dead entries are swept into the pool once this many sit in the stripes   // dead entries are swept into the pool only while the queue is this large
#ifndef HP_SUB_SERIAL_SCAN
#define HP_SUB_SERIAL_SCAN 4
#endif
constexpr uint32_t kSubSerialScan = HP_SUB_SERIAL_SCAN;   // stripes up to this long are rescanned by their owner lane alone
// Per-warp global region behind the shared-memory part of the sub-solver queue: the spill part of the AoS stripes (generic
// path) or, for the fast path, one 16-byte payload per node index (node indices reach 1 + 4*max_visits <= 43*capl + 16).
__host__ __device__ inline uint64_t sub_spill_bytes(uint32_t capl, uint32_t capl_s) {
    const uint64_t aos = (uint64_t)32 * (capl - capl_s) * 32;
    const uint64_t pay = (uint64_t)16 * (43ull * capl + 16);
    return ((aos > pay ? aos : pay) + 255) & ~255ull;
}
constexpr uint32_t kFreeStack = 192;   // free main-queue record slots kept in shared memory
// Sub-solver score vectors as packed u16x2 (a read's cost against h1 in the low half, against h2 in the high half: a window of
// <= 63 variants of quality <= 255 stays below 2^16), so the per-slot child arithmetic runs on the packed-halfword min
// (VIMNMX.U16x2, the DPX family of sm_90+ / sm_100) and PRMT, and the carry into the next column moves two words, not four.
#ifndef HP_SUB_PACKED
#define HP_SUB_PACKED 1
#endif

struct WarpCtx {
    SubEntry* sq;           // this warp's sub-solver queue: first capl_s entries of every stripe, in shared memory
    SubEntry* sq_spill;     // ... and the rest of every stripe in the warp's global slab (rarely touched)
    uint32_t* hring;        // H[] ring buffer, 64 entries, shared by the warps of a team
    uint32_t capl;          // stripe capacity (entries per lane)
    uint32_t capl_s;        // of which in shared memory
    uint32_t lane;
    uint32_t h_floor;       // speculation: H[] entries below this index are not known yet and read as H[h_floor]
    uint64_t* mq_hi;        // main queue: the first mq_cap_s keys of every stripe live in the CTA's shared memory
    uint32_t* mq_idx;       //   (the sub-solver queues of the whole team are idle while warp 0 runs the main loop)
    uint32_t* mq_len;       //   key = (hi, idx); len = length | identical flag << 31; rec = record slot of the node
    uint32_t* mq_rec;
    uint32_t mq_cap_s;
    uint32_t* free_stack;   // small stack of free record slots in shared memory (overflow: slab free list)
    uint32_t* free_ctr;     // shared-memory cursors into the slab free list / the dead pool while a sweep runs
    uint32_t* pool_ctr;
    uint32_t* lencnt_s;     // PQueueHapTracker::length_counts in shared memory (blocks with <= lencnt_cap_s variants)
    uint32_t lencnt_cap_s;
    // counters
    uint64_t evals, sum_lp, pops, cells;
    long long ts_pop, ts_seat, ts_score, ts_rest, ts_popa, ts_popb;   // counting variant: sub-solver phase cycles of this warp
    uint64_t ns_real, ns_planes, ns_exp;
    int status;
    // entry i of stripe `stripe`
    __device__ __forceinline__ SubEntry* ent(uint32_t stripe, uint32_t i) const {
        return i < capl_s ? sq + stripe * capl_s + i : sq_spill + stripe * (capl - capl_s) + (i - capl_s);
    }
    __device__ __forceinline__ uint32_t H(uint32_t idx) const { return hring[max(idx, h_floor) & 63u]; }
};

// tag 0xff: h1/h2 are the node's own haplotypes; tag < 4: h1/h2 are the PARENT's and tag is the node's candidate
// slot -- its allele bit is applied when the node is popped (saves the per-sibling bit fiddling in the dive loop).
__device__ __forceinline__ void sub_store(SubEntry* e, uint64_t key, uint64_t h1, uint64_t h2, uint32_t frozen, uint32_t tag = 0xffu) {
    uint4* p = reinterpret_cast<uint4*>(e);
    p[0] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)h1, (uint32_t)(h1 >> 32));
    p[1] = make_uint4((uint32_t)h2, (uint32_t)(h2 >> 32), frozen, tag);
}

// Removes entry `pos` of stripe `owner` (swap with the stripe's last entry) and recomputes that stripe's cached minimum:
// by the owner lane alone for short stripes, by the whole warp (one entry per lane, then a warp minimum) for longer ones.
__device__ __forceinline__ void sub_remove_rescan(const WarpCtx& w, int owner, uint32_t pos, uint32_t& cnt, uint64_t& ckey, uint32_t& cpos) {
    const uint32_t lane = w.lane;
    const uint32_t cnt_o = __shfl_sync(HP_FULL_MASK, cnt, owner) - 1u;
    if ((int)lane == owner) {
        cnt--;
        if (pos != cnt) {
            const uint4* sp = reinterpret_cast<const uint4*>(w.ent(lane, cnt));
            uint4* dp = reinterpret_cast<uint4*>(w.ent(lane, pos));
            dp[0] = sp[0]; dp[1] = sp[1];
        }
    }
    if (cnt_o <= kSubSerialScan) {
        if ((int)lane == owner) {
            ckey = ~0ull; cpos = 0;
            for (uint32_t i = 0; i < cnt; i++) {
                const uint64_t k = w.ent(lane, i)->key;
                if (k < ckey) { ckey = k; cpos = i; }
            }
        }
        return;
    }
    __syncwarp();
    uint64_t bk = ~0ull; uint32_t bp = 0;
    for (uint32_t i = lane; i < cnt_o; i += 32) {
        const uint64_t k = w.ent((uint32_t)owner, i)->key;
        if (k < bk) { bk = k; bp = i; }
    }
    const uint64_t mk = wmin64(bk);
    const int wl = __ffs(__ballot_sync(HP_FULL_MASK, bk == mk)) - 1;
    const uint32_t mp = __shfl_sync(HP_FULL_MASK, bp, wl);
    if ((int)lane == owner) { ckey = mk; cpos = mp; }
}

// astar_subsolver (astar_phaser.rs:311-405).  Returns est in .x, solved depth in .y (both warp-uniform).
//
// "cur" is the node on top of the queue, held in registers.  After an expansion the best child becomes cur
// directly when its key beats the queue minimum (the dive): it is never written to the queue, and its score
// vector comes from the expansion cache.  Otherwise cur is pushed and a real pop (redux.sync select) happens.
template <int K, bool kCount>
__device__ uint2 sub_solve_generic(const AstarArgs& a, const BlkMeta& m, WarpCtx& w, uint32_t v, uint32_t clip, uint64_t badwin,
                           uint32_t blk) {
    const uint32_t lane = w.lane;
    const uint32_t* aoff = a.act_off + m.var_base + blk;
    const uint32_t* aidx = a.act_idx + m.cell_base;
    const uint32_t* col = a.col + m.cell_base;
    const ReadMeta* rmeta = a.rmeta + m.read_base;

    // queue state: cached stripe minimum + stripe count in registers; qmin = warp-uniform queue minimum
    uint64_t ckey = ~0ull, qmin = ~0ull;
    uint32_t cpos = 0, cnt = 0;
    // root: AstarNode::new(H[v+1]) (:325), kept in registers as the current top
    uint32_t cur_total = w.H(v + 1), cur_lo = 63u << 26, cur_frozen = 0;
    uint64_t cur_h1 = 0, cur_h2 = 0;
    int cur_src = SRC_ROOT;
    uint32_t cur_x1 = 0, cur_x2 = 0;
    ExpCache<K> cache;
    cache.first_idx = 0xffffffffu; cache.present = 0;
    uint32_t next_idx = 1, next_expected = 0, max_cost = 0, visits = 0, rr = 0;
    const uint32_t max_visits = a.min_queue_size / 10 + a.queue_increment * clip;    // :266, :333

    for (;;) {
        if (qmin < mk64(cur_total, cur_lo)) {
            // ---- the dive broke: cur goes back to the queue, then a real pop of the entry whose key is qmin ----
            {
                const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < w.capl);
                if (room == 0) { w.status = HP_BLOCK_ASSERT; break; }
                uint32_t target = rr & 31u; rr++;
                if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                if (lane == target) {
                    const uint64_t k = mk64(cur_total, cur_lo);
                    sub_store(w.ent(lane, cnt), k, cur_h1, cur_h2, cur_frozen);
                    if (k < ckey) { ckey = k; cpos = cnt; }
                    cnt++;
                }
                __syncwarp();
            }
            const int owner = __ffs(__ballot_sync(HP_FULL_MASK, ckey == qmin)) - 1;
            const uint32_t pos = __shfl_sync(HP_FULL_MASK, cpos, owner);
            const SubEntry* e = w.ent(owner, pos);
            cur_total = (uint32_t)(qmin >> 32); cur_lo = (uint32_t)qmin;
            cur_h1 = e->h1; cur_h2 = e->h2; cur_frozen = e->frozen;
            __syncwarp();
            sub_remove_rescan(w, owner, pos, cnt, ckey, cpos);
            qmin = wmin64(ckey);
            // where does this node's score vector come from?
            const uint32_t d = ((cur_lo >> 6) & 0xfffffu) - cache.first_idx;
            if ((cur_lo & 63u) == 0u) cur_src = SRC_ROOT;
            else if (K > 0 && d < (uint32_t)__popc(cache.present)) {
                const uint32_t cs = slot_of_ordinal(cache.present, d);
                cur_src = SRC_CACHE; cur_x1 = cs & 1u; cur_x2 = (0x9u >> cs) & 1u;
            } else cur_src = SRC_PLANES;
        }
        // ---- cur is the top of the queue (peek) ----
        const uint32_t L = cur_lo & 63u;
        if (L >= clip) {                                                 // :395-399 (peek, not pop)
            max_cost = max(max_cost, cur_total);
            next_expected++;
            break;
        }
        if (visits >= max_visits) break;
        visits++;
        if (L == next_expected) { max_cost = max(max_cost, cur_total); next_expected++; }   // :342-346

        // ---- expand ----
        const uint32_t p = v + L;
        const bool bad_col = (badwin >> L) & 1ull;
        const uint32_t heur = w.H(p + 1);
        const uint32_t o0 = __ldg(aoff + p), o1 = __ldg(aoff + p + 1);
        const bool ident = (cur_h1 == cur_h2);
        auto hap = [&](int which, int i0) { return shift_signed(which ? cur_h2 : cur_h1, i0); };
        uint32_t tot[4], fro[4];
        uint64_t cells = 0;
        expand<K, kCount>(a, m, lane, col + o0, aidx + o0, rmeta, o1 - o0, p, (int)v, (int)L, bad_col, ident, cur_src,
                          cur_x1, cur_x2, hap, cache, tot, fro, cells);
        const uint32_t present = present_mask(bad_col, ident);
        const uint32_t nchild = bad_col ? 1u : (ident ? 3u : 4u);
        cache.first_idx = next_idx; cache.present = present;
        if (kCount) { w.pops++; w.evals += nchild; w.sum_lp += (uint64_t)nchild * L; w.cells += cells * nchild; }

        // candidate keys.  total_c = frozen + tot_c + heur; the low words are ordered lo0 < lo1 < lo2 < lo3 (more
        // hets first, then creation order), so ties on the total are won by the lowest candidate slot.
        const uint32_t fh = cur_frozen + heur;
        const uint32_t t0 = bad_col ? 0xffffffffu : fh + tot[0];
        const uint32_t t1 = (bad_col || ident) ? 0xffffffffu : fh + tot[1];
        const uint32_t t2 = fh + tot[2];
        const uint32_t t3 = bad_col ? 0xffffffffu : fh + tot[3];
        if (bad_col && t2 != cur_total) { w.status = HP_BLOCK_ASSERT; break; }             // :360
        const uint32_t lo_base = (cur_lo & 0xfc000000u) | (next_idx << 6) | (L + 1);
        const uint32_t lo0 = lo_base - (1u << 26), lo1 = lo0 + 64u;
        const uint32_t lo2 = lo_base + (bad_col ? 0u : (ident ? 64u : 128u)), lo3 = lo2 + 64u;
        const uint32_t tmin = min(min(t0, t1), min(t2, t3));
        const uint32_t best = (t0 == tmin) ? 0u : (t1 == tmin) ? 1u : (t2 == tmin) ? 2u : 3u;

        // ---- push the siblings of the best child (one lane each, round-robin over the stripes) ----
        const uint64_t bit = bad_col ? 0ull : (1ull << L);
        {
            const uint32_t c = (lane - rr) & 31u;
            const bool mine = c < 4u && ((present >> c) & 1u) && c != best;
            const uint32_t fullmask = __ballot_sync(HP_FULL_MASK, mine && cnt >= w.capl);
            const uint32_t mt = (c == 0u) ? t0 : (c == 1u) ? t1 : (c == 2u) ? t2 : t3;
            const uint32_t ml = (c == 0u) ? lo0 : (c == 1u) ? lo1 : (c == 2u) ? lo2 : lo3;
            const uint32_t mf = cur_frozen + ((c == 0u) ? fro[0] : (c == 1u) ? fro[1] : (c == 2u) ? fro[2] : fro[3]);
            if (fullmask == 0) {
                if (mine) {
                    const uint64_t k = mk64(mt, ml);
                    sub_store(w.ent(lane, cnt), k, cur_h1 | ((c & 1u) ? bit : 0ull), cur_h2 | (((0x9u >> c) & 1u) ? bit : 0ull), mf);
                    if (k < ckey) { ckey = k; cpos = cnt; }
                    cnt++;
                }
            } else {                                                     // rare: a target stripe is full
                for (uint32_t cc = 0; cc < 4; cc++) {
                    if (((present >> cc) & 1u) && cc != best) {
                        const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < w.capl);
                        if (room == 0) { w.status = HP_BLOCK_ASSERT; break; }
                        const uint32_t src_lane = (rr + cc) & 31u;        // the lane that computed this child's fields
                        uint32_t target = src_lane;
                        if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                        const uint32_t xt = __shfl_sync(HP_FULL_MASK, mt, src_lane), xl = __shfl_sync(HP_FULL_MASK, ml, src_lane);
                        const uint32_t xf = __shfl_sync(HP_FULL_MASK, mf, src_lane);
                        if (lane == target) {
                            const uint64_t k = mk64(xt, xl);
                            sub_store(w.ent(lane, cnt), k, cur_h1 | ((cc & 1u) ? bit : 0ull), cur_h2 | (((0x9u >> cc) & 1u) ? bit : 0ull), xf);
                            if (k < ckey) { ckey = k; cpos = cnt; }
                            cnt++;
                        }
                    }
                }
            }
            rr += 4;
        }
        // queue minimum now includes the siblings
        {
            const uint64_t k0 = (best == 0u) ? ~0ull : mk64(t0, lo0), k1 = (best == 1u) ? ~0ull : mk64(t1, lo1);
            const uint64_t k2 = (best == 2u) ? ~0ull : mk64(t2, lo2), k3 = (best == 3u) ? ~0ull : mk64(t3, lo3);
            const uint64_t ka = k0 < k1 ? k0 : k1, kb = k2 < k3 ? k2 : k3;
            const uint64_t kc = ka < kb ? ka : kb;
            qmin = kc < qmin ? kc : qmin;
        }
        next_idx += nchild;
        // ---- the best child is the new cur; its score vector is in the expansion cache ----
        cur_total = tmin;
        cur_lo = (best == 0u) ? lo0 : (best == 1u) ? lo1 : (best == 2u) ? lo2 : lo3;
        cur_frozen += (best == 0u) ? fro[0] : (best == 1u) ? fro[1] : (best == 2u) ? fro[2] : fro[3];
        cur_x1 = best & 1u; cur_x2 = (0x9u >> best) & 1u;
        cur_h1 |= cur_x1 ? bit : 0ull;
        cur_h2 |= cur_x2 ? bit : 0ull;
        cur_src = (K > 0) ? SRC_CACHE : SRC_PLANES;
        __syncwarp();
        if (w.status != HP_BLOCK_OK) break;
    }
    return make_uint2(max_cost, next_expected - 1);
}

extern __shared__ __align__(16) uint8_t hp_dyn_smem[];     // the CTA's dynamic shared memory (true shared-space addressing)

// astar_subsolver, register-resident fast path for blocks with at most 32*K reads per column.
//
// Critical path per pop (the dive): select (s1,s2) of the chosen child -> A0..B1 -> four mins -> deltas against
// base = min(s1,s2) -> 2 redux.sync on 16-bit packed deltas -> child totals -> best child -> next pop.
//   * child total = cur_total - H[p] + H[p+1] + sum_r delta_c(r): the parent's fluid cost is already inside its
//     total, and each delta is <= the column quality (<= 255), so two deltas share one 32-bit redux.
//   * the frozen part (reads ending at this column) is only reduced when some read ends here.
//   * the four child values per read are shuffled into the NEXT column's slot order right away (carry index of the
//     prefetched column record), off the critical path, so the next pop starts with two selects.
template <int K, bool kCount>
__device__ uint2 sub_solve_fast(const AstarArgs& a, const BlkMeta& m, WarpCtx& w, uint32_t v, uint32_t clip, uint64_t badwin,
                                uint32_t blk) {
    const uint32_t lane = w.lane;
    const uint32_t N = m.n_var;
    const uint32_t* const hring = w.hring;                 // H[] ring and speculation floor in registers for the whole sub-solve
    const uint32_t h_floor = w.h_floor;
    auto Hq = [&](uint32_t idx) { return hring[max(idx, h_floor) & 63u]; };
    const uint32_t* aoff = a.act_off + m.var_base + blk;
    const uint32_t* aidx = a.act_idx + m.cell_base;
    const uint32_t* col = a.col + m.cell_base;
    const uint32_t* col_lane = col + lane;
    const ReadMeta* rmeta = a.rmeta + m.read_base;
    constexpr uint32_t kEmpty = 0xffff0000u;      // column record of an unused slot: no carry, quality 0
    // queue layout of this path: keys only in shared memory (KS per stripe, all stripes fit), one 16-byte payload per node
    // index in the warp's global region: {h1, h2 | tag << 56}.  tag 0xff: the node's own haplotypes; tag < 4: the PARENT's
    // haplotypes and the node's candidate slot.  Entries never move their payload; the payload load of a pop is issued
    // as soon as the popped key is known and overlaps the stripe bookkeeping.
    const uint32_t KS = w.capl_s * 4u;                                      // 32-byte AoS slots -> 8-byte keys
    uint64_t* const keys = reinterpret_cast<uint64_t*>(hp_dyn_smem) + (size_t)(threadIdx.x >> 5) * 32u * KS;
    uint64_t* const my_keys = keys + lane * KS;
    uint4* const pay = reinterpret_cast<uint4*>(w.sq_spill);
    auto pay_store = [&](uint32_t idx, uint64_t h1, uint64_t h2, uint32_t tag) {
        pay[idx] = make_uint4((uint32_t)h1, (uint32_t)(h1 >> 32), (uint32_t)h2, (uint32_t)(h2 >> 32) | (tag << 24));
    };

    // A sub-solve whose dive never breaks (the best child stays the queue minimum all the way to the clip: ~95 % of them on
    // HiFi-like data) never reads what it pushed.  So the first attempt is LAZY: siblings are not stored, only their
    // minimum key is tracked; if the dive would break, the sub-solve starts over with a real queue.  Same results.
    const uint64_t c_pops = w.pops, c_evals = w.evals, c_lp = w.sum_lp, c_cells = w.cells;
    bool lazy = true;
restart:
    // queue state
    uint64_t ckey = ~0ull, qmin = ~0ull;
    uint32_t cpos = 0, cnt = 0;
    // current top (root, :325)
    uint32_t cur_total = Hq(v + 1), cur_lo = 63u << 26;
    uint64_t cur_h1 = 0, cur_h2 = 0;
    int cur_src = SRC_ROOT;
    uint32_t cur_x1 = 0, cur_x2 = 0;
    uint32_t heur_p = cur_total;                    // heuristic term inside cur_total (root quirk: H[v+1])
    // children of the last expansion, already in the slot order of the column after it
#if HP_SUB_PACKED
    uint32_t nX[K], nY[K], nW[K];                   // X = A0 | B0 << 16, Y = A1 | B1 << 16
#else
    uint32_t nA0[K], nA1[K], nB0[K], nB1[K], nW[K];
#endif
    uint32_t cache_first = 0xffffffffu, cache_present = 0;
    // column records of the current column p (slot = lane + 32k) and the offsets of p and p+1
    // column records two columns ahead are always in flight (the small L1 next to 196 KB of shared memory misses often)
    uint32_t o_p = __ldg(aoff + v), o_p1 = __ldg(aoff + v + 1), o_p2 = __ldg(aoff + min(v + 2, N));      // aoff[N] = all cells: empty columns beyond N
    uint32_t colc[K], coln[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
#if HP_SUB_PACKED
        nX[k] = nY[k] = nW[k] = 0;
#else
        nA0[k] = nA1[k] = nB0[k] = nB1[k] = nW[k] = 0;
#endif
        colc[k] = (lane + 32u * k < o_p1 - o_p) ? __ldg(col + o_p + lane + 32u * k) : kEmpty;
        coln[k] = (v + 1 < N) ? __ldg(col_lane + o_p1 + 32u * k) : kEmpty;
    }
    uint32_t next_idx = 1, next_expected = 0, max_cost = 0, visits = 0, rr = 0;
    uint64_t lbit = 1ull;                           // 1 << length of cur: shifted along the dive, re-seated after a real pop
    const uint32_t max_visits = a.min_queue_size / 10 + a.queue_increment * clip;    // :266, :333

    long long tq = 0;
    for (;;) {
        if (kCount) tq = clock64();
        if (qmin < mk64(cur_total, cur_lo)) {
            if (lazy) {                                                  // start over, this time keeping the siblings
                lazy = false;
                if (kCount) { w.pops = c_pops; w.evals = c_evals; w.sum_lp = c_lp; w.cells = c_cells; }
                goto restart;
            }
            if (kCount) w.ns_real++;
            // ---- the dive broke: cur goes back to the queue, then a real pop of the entry whose key is qmin ----
            {
                const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < KS);
                if (room == 0) { w.status = HP_BLOCK_ASSERT; break; }
                uint32_t target = rr & 31u; rr++;
                if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                if (lane == target) {
                    const uint64_t k = mk64(cur_total, cur_lo);
                    my_keys[cnt] = k;
                    pay_store((cur_lo >> 6) & 0xfffffu, cur_h1, cur_h2, 0xffu);
                    if (k < ckey) { ckey = k; cpos = cnt; }
                    cnt++;
                }
                __syncwarp();
            }
            if (kCount) w.ts_popa += clock64() - tq;          // push-back of cur
            cur_total = (uint32_t)(qmin >> 32); cur_lo = (uint32_t)qmin;
            const uint4 pl = pay[(cur_lo >> 6) & 0xfffffu];               // in flight while the stripe is fixed up
            const int owner = __ffs(__ballot_sync(HP_FULL_MASK, ckey == qmin)) - 1;
            const uint32_t pos = __shfl_sync(HP_FULL_MASK, cpos, owner);
            if (kCount) w.ts_popb += clock64() - tq;          // + owner
            {   // remove (swap with the stripe's last key) and rescan that stripe
                const uint32_t left = __shfl_sync(HP_FULL_MASK, cnt, owner) - 1u;
                uint64_t* const ok = keys + (uint32_t)owner * KS;
                if ((int)lane == owner) { cnt--; if (pos != cnt) ok[pos] = ok[cnt]; }
                __syncwarp();
                uint64_t bk = ~0ull; uint32_t bp = 0;
                for (uint32_t q = lane; q < left; q += 32) { const uint64_t k = ok[q]; if (k < bk) { bk = k; bp = q; } }
                const uint64_t mk = wmin64(bk);
                const int wl = __ffs(__ballot_sync(HP_FULL_MASK, bk == mk)) - 1;
                const uint32_t mp = __shfl_sync(HP_FULL_MASK, bp, wl);
                if ((int)lane == owner) { ckey = mk; cpos = mp; }
            }
            qmin = wmin64(ckey);
            cur_h1 = ((uint64_t)pl.y << 32) | pl.x; cur_h2 = ((uint64_t)(pl.w & 0x00ffffffu) << 32) | pl.z;
            {
                const uint32_t tag = pl.w >> 24;
                if (tag < 4u) {                                          // parent's haplotypes + this node's candidate slot
                    const uint32_t lb = (cur_lo & 63u) - 1u;
                    cur_h1 |= (uint64_t)(tag & 1u) << lb;
                    cur_h2 |= (uint64_t)((0x9u >> tag) & 1u) << lb;
                }
            }
            if (kCount) { const long long t1 = clock64(); w.ts_pop += t1 - tq; tq = t1; }
            // re-seat the column state on this node's position
            const uint32_t Lp = cur_lo & 63u;
            const uint32_t pp = v + Lp;
            lbit = 1ull << Lp;
            const uint32_t d = ((cur_lo >> 6) & 0xfffffu) - cache_first;
            if (Lp == 0u) cur_src = SRC_ROOT;
            else if (d < (uint32_t)__popc(cache_present)) {
                // a child of the last expansion: the pre-permuted vectors are for exactly this column
                const uint32_t cs = slot_of_ordinal(cache_present, d);
                cur_src = SRC_CACHE; cur_x1 = cs & 1u; cur_x2 = (0x9u >> cs) & 1u;
            } else cur_src = SRC_PLANES;
            heur_p = Hq(Lp == 0u ? v + 1 : pp);
            if (pp < N) {
                o_p = __ldg(aoff + pp); o_p1 = __ldg(aoff + pp + 1); o_p2 = __ldg(aoff + min(pp + 2, N));
#pragma unroll
                for (int k = 0; k < K; k++) {
                    colc[k] = (lane + 32u * k < o_p1 - o_p) ? __ldg(col + o_p + lane + 32u * k) : kEmpty;
                    coln[k] = (pp + 1 < N) ? __ldg(col_lane + o_p1 + 32u * k) : kEmpty;
                }
            }
        }
        if (kCount) { const long long t1 = clock64(); w.ts_seat += t1 - tq; tq = t1; }
        // ---- cur is the top of the queue (peek) ----
        const uint32_t L = cur_lo & 63u;
        if (L >= clip) {                                                 // :395-399 (peek, not pop)
            max_cost = max(max_cost, cur_total);
            next_expected++;
            break;
        }
        if (visits >= max_visits) break;
        visits++;
        if (L == next_expected) { max_cost = max(max_cost, cur_total); next_expected++; }   // :342-346

        const uint32_t p = v + L;
        const uint32_t a_cur = o_p1 - o_p;
        // ---- prefetch the next column's records (addresses are known; validity is masked once o_p2 arrives) ----
        const uint32_t o_p3 = __ldg(aoff + min(p + 3, N));
        // (records read past the block's last column are masked by the empty column's size; the array has 130 entries of slack)
        uint32_t colnn[K];
#pragma unroll
        for (int k = 0; k < K; k++) colnn[k] = __ldg(col_lane + o_p2 + 32u * k);
        const uint32_t heur = Hq(p + 1);
        const bool bad_col = (badwin & lbit) != 0ull;
        const bool ident = (cur_h1 == cur_h2);

#if HP_SUB_PACKED
        // ---- this node's (s1 | s2 << 16) per slot ----
        uint32_t s12[K], wv[K];
        if (cur_src == SRC_CACHE) {
            // low half from X (A0) or Y (A1) by the node's h1 allele, high half from X (B0) or Y (B1) by its h2 allele
            const uint32_t sel = (cur_x1 ? 0x54u : 0x10u) | (cur_x2 ? 0x7600u : 0x3200u);
#pragma unroll
            for (int k = 0; k < K; k++) {
                // slots beyond the column's coverage hold nothing (warp-uniform test): their work is skipped below as well
                if (k == 0 || a_cur > 32u * k) { s12[k] = __byte_perm(nX[k], nY[k], sel); wv[k] = nW[k]; }
                else { s12[k] = 0; wv[k] = 0; }
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) { s12[k] = 0; wv[k] = 0; }
            if (cur_src == SRC_PLANES) {
                auto hap = [&](int which, int i0) { return shift_signed(which ? cur_h2 : cur_h1, i0); };
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const uint32_t j = lane + 32u * k;
                    if (j < a_cur) {
                        const ReadMeta rm = rmeta[__ldg(aidx + o_p + j)];
                        uint32_t s1, s2;
                        score_planes(a, m, rm, (int)v - (int)rm.start, (int)L, hap, s1, s2);
                        s12[k] = s1 | (s2 << 16);
                        wv[k] = p - (uint32_t)max((int)rm.start, (int)v);
                    }
                }
            }
        }
        // ---- children: X = (A0 | B0 << 16) = s12 + q0, Y = (A1 | B1 << 16) = s12 + q1; deltas against base, two per word ----
        uint32_t X[K], Y[K];
        uint32_t pk0 = 0, pk1 = 0;
        uint64_t cells = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (k > 0 && a_cur <= 32u * k) { X[k] = Y[k] = 0; continue; }                       // empty slot: contributes nothing
            const uint32_t c = colc[k];
            const uint32_t qq = bad_col ? 0u : (c & 0xffu) * 0x10001u;                          // quality in both halves
            const uint32_t al = c & 0x300u;
            X[k] = s12[k] + ((al != 0u) ? qq : 0u);
            Y[k] = s12[k] + ((al != 0x100u) ? qq : 0u);
            const uint32_t base2 = __vminu2(s12[k], __byte_perm(s12[k], 0, 0x1032));            // min(s1, s2) in both halves
            const uint32_t m01 = __vminu2(X[k], __byte_perm(Y[k], 0, 0x1032));                  // min(A0, B1) | min(B0, A1) << 16
            const uint32_t m00 = __vminu2(__byte_perm(X[k], Y[k], 0x5410), __byte_perm(X[k], Y[k], 0x7632));   // min(A0, B0) | min(A1, B1) << 16
            pk0 += m01 - base2;
            pk1 += m00 - base2;
            if (kCount) { wv[k] += 1; if (lane + 32u * k < a_cur) cells += wv[k]; }
        }
#else
        // ---- this node's (s1, s2) per slot ----
        uint32_t s1[K], s2[K], wv[K];
        if (cur_src == SRC_CACHE) {
#pragma unroll
            for (int k = 0; k < K; k++) {
                // slots beyond the column's coverage hold nothing (warp-uniform test): their work is skipped below as well
                if (k == 0 || a_cur > 32u * k) { s1[k] = cur_x1 ? nA1[k] : nA0[k]; s2[k] = cur_x2 ? nB1[k] : nB0[k]; wv[k] = nW[k]; }
                else { s1[k] = 0; s2[k] = 0; wv[k] = 0; }
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) { s1[k] = 0; s2[k] = 0; wv[k] = 0; }
            if (cur_src == SRC_PLANES) {
                auto hap = [&](int which, int i0) { return shift_signed(which ? cur_h2 : cur_h1, i0); };
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const uint32_t j = lane + 32u * k;
                    if (j < a_cur) {
                        const ReadMeta rm = rmeta[__ldg(aidx + o_p + j)];
                        score_planes(a, m, rm, (int)v - (int)rm.start, (int)L, hap, s1[k], s2[k]);
                        wv[k] = p - (uint32_t)max((int)rm.start, (int)v);
                    }
                }
            }
        }
        // ---- children: deltas against base, packed two per word ----
        uint32_t A0[K], A1[K], B0[K], B1[K];
        uint32_t pk0 = 0, pk1 = 0;
        uint64_t cells = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (k > 0 && a_cur <= 32u * k) { A0[k] = A1[k] = B0[k] = B1[k] = 0; continue; }      // empty slot: contributes nothing
            const uint32_t c = colc[k];
            const uint32_t q = bad_col ? 0u : (c & 0xffu);
            const uint32_t al = (c >> 8) & 3u;
            const uint32_t q0 = (al != 0u) ? q : 0u, q1 = (al != 1u) ? q : 0u;
            A0[k] = s1[k] + q0; A1[k] = s1[k] + q1; B0[k] = s2[k] + q0; B1[k] = s2[k] + q1;
            const uint32_t base = min(s1[k], s2[k]);
            const uint32_t c01 = min(A0[k], B1[k]), c10 = min(A1[k], B0[k]), c00 = min(A0[k], B0[k]), c11 = min(A1[k], B1[k]);
            pk0 += (c01 - base) | ((c10 - base) << 16);
            pk1 += (c00 - base) | ((c11 - base) << 16);
            if (kCount) { wv[k] += 1; if (lane + 32u * k < a_cur) cells += wv[k]; }
        }
#endif
        const uint32_t r0 = wsum(pk0), r1 = wsum(pk1);

        // the common expansion (a heterozygous choice was made earlier, the column is not ignored) has all four children:
        // a warp-uniform branch keeps its key arithmetic free of the selects of the special cases
        const bool plain = !bad_col && !ident;
        const uint32_t present = plain ? 0xfu : present_mask(bad_col, ident);
        const uint32_t nchild = plain ? 4u : (bad_col ? 1u : 3u);
        if (kCount) { const long long t1 = clock64(); w.ts_score += t1 - tq; tq = t1; w.ns_exp++; if (cur_src == SRC_PLANES) w.ns_planes++; }
        if (kCount) { w.pops++; w.evals += nchild; w.sum_lp += (uint64_t)nchild * L; w.cells += cells * nchild; }

        // candidate totals / keys; low words are ordered lo0 < lo1 < lo2 < lo3, so the first minimum wins ties
        const uint32_t tb = cur_total - heur_p + heur;
        const uint32_t lo_base = (cur_lo & 0xfc000000u) | (next_idx << 6) | (L + 1);
        // lo of candidate c = (c < 2 ? lo0 : lo2) + ((c & 1) << 6)
        const uint32_t lo0 = lo_base - (1u << 26);
        uint32_t t0, t1, t2, t3, lo2;
        if (plain) {
            t0 = tb + (r0 & 0xffffu); t1 = tb + (r0 >> 16); t2 = tb + (r1 & 0xffffu); t3 = tb + (r1 >> 16);
            lo2 = lo_base + 128u;
        } else {
            t0 = bad_col ? 0xffffffffu : tb + (r0 & 0xffffu);
            t1 = 0xffffffffu;                                            // (1|0) only exists next to a different (0|1)
            t2 = tb + (r1 & 0xffffu);
            t3 = bad_col ? 0xffffffffu : tb + (r1 >> 16);
            if (bad_col && t2 != cur_total) { w.status = HP_BLOCK_ASSERT; break; }             // :360
            lo2 = lo_base + (bad_col ? 0u : 64u);
        }
        const uint32_t tmin = min(min(t0, t1), min(t2, t3));
        const uint32_t best = (t0 == tmin) ? 0u : (t1 == tmin) ? 1u : (t2 == tmin) ? 2u : 3u;

        // ---- children vectors into the next column's slot order (off the critical path) ----
        {
#if HP_SUB_PACKED
            const uint32_t a_next = o_p2 - o_p1;
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (k > 0 && a_next <= 32u * k) {                        // the next column has no read in this slot
                    nX[k] = nY[k] = nW[k] = 0;
                    colc[k] = kEmpty;
                    coln[k] = colnn[k];
                    continue;
                }
                const uint32_t cn = (lane + 32u * k < a_next) ? coln[k] : kEmpty;
                const uint32_t carry = cn >> 16, sl = carry & 31u;
                uint32_t g0 = __shfl_sync(HP_FULL_MASK, X[0], sl), g1 = __shfl_sync(HP_FULL_MASK, Y[0], sl);
                uint32_t gw = kCount ? __shfl_sync(HP_FULL_MASK, wv[0], sl) : 0u;
#pragma unroll
                for (int kk = 1; kk < K; kk++) {
                    if (a_cur <= 32u * kk) continue;                     // nothing to carry out of an empty slot
                    const uint32_t u0 = __shfl_sync(HP_FULL_MASK, X[kk], sl), u1 = __shfl_sync(HP_FULL_MASK, Y[kk], sl);
                    const uint32_t uw = kCount ? __shfl_sync(HP_FULL_MASK, wv[kk], sl) : 0u;
                    if ((carry >> 5) == (uint32_t)kk) { g0 = u0; g1 = u1; gw = uw; }
                }
                const bool has = carry != 0xffffu;
                nX[k] = has ? g0 : 0u; nY[k] = has ? g1 : 0u;
                nW[k] = has ? gw : 0u;
                colc[k] = cn;
                coln[k] = colnn[k];
            }
#else
            const uint32_t a_next = o_p2 - o_p1;
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (k > 0 && a_next <= 32u * k) {                        // the next column has no read in this slot
                    nA0[k] = nA1[k] = nB0[k] = nB1[k] = nW[k] = 0;
                    colc[k] = kEmpty;
                    coln[k] = colnn[k];
                    continue;
                }
                const uint32_t cn = (lane + 32u * k < a_next) ? coln[k] : kEmpty;
                const uint32_t carry = cn >> 16, sl = carry & 31u;
                uint32_t g0 = __shfl_sync(HP_FULL_MASK, A0[0], sl), g1 = __shfl_sync(HP_FULL_MASK, A1[0], sl);
                uint32_t g2 = __shfl_sync(HP_FULL_MASK, B0[0], sl), g3 = __shfl_sync(HP_FULL_MASK, B1[0], sl);
                uint32_t gw = kCount ? __shfl_sync(HP_FULL_MASK, wv[0], sl) : 0u;
#pragma unroll
                for (int kk = 1; kk < K; kk++) {
                    if (a_cur <= 32u * kk) continue;                     // nothing to carry out of an empty slot
                    const uint32_t u0 = __shfl_sync(HP_FULL_MASK, A0[kk], sl), u1 = __shfl_sync(HP_FULL_MASK, A1[kk], sl);
                    const uint32_t u2 = __shfl_sync(HP_FULL_MASK, B0[kk], sl), u3 = __shfl_sync(HP_FULL_MASK, B1[kk], sl);
                    const uint32_t uw = kCount ? __shfl_sync(HP_FULL_MASK, wv[kk], sl) : 0u;
                    if ((carry >> 5) == (uint32_t)kk) { g0 = u0; g1 = u1; g2 = u2; g3 = u3; gw = uw; }
                }
                const bool has = carry != 0xffffu;
                nA0[k] = has ? g0 : 0u; nA1[k] = has ? g1 : 0u; nB0[k] = has ? g2 : 0u; nB1[k] = has ? g3 : 0u;
                nW[k] = has ? gw : 0u;
                colc[k] = cn;
                coln[k] = colnn[k];
            }
#endif
            o_p = o_p1; o_p1 = o_p2; o_p2 = o_p3;
        }
        cache_first = next_idx; cache_present = present;

        // ---- push the siblings of the best child: candidate c is handled by lane (rr + c) & 31, into its own stripe.
        //      The entry keeps the PARENT's haplotypes and the candidate slot as a tag. ----
        if (!lazy) {
            const uint32_t c = (lane - rr) & 31u;
            const bool mine = c < 4u && ((present >> c) & 1u) && c != best;
            const uint32_t rsel = (c & 2u) ? r1 : r0;
            const uint32_t mt = tb + ((c & 1u) ? (rsel >> 16) : (rsel & 0xffffu));
            const uint32_t ml = ((c & 2u) ? lo2 : lo0) + ((c & 1u) << 6);
            if (__ballot_sync(HP_FULL_MASK, mine && cnt >= KS) == 0) {
                if (mine) {
                    const uint64_t k = mk64(mt, ml);
                    my_keys[cnt] = k;
                    pay_store((ml >> 6) & 0xfffffu, cur_h1, cur_h2, c);
                    if (k < ckey) { ckey = k; cpos = cnt; }
                    cnt++;
                }
            } else {                                                     // rare: a target stripe is full
                for (uint32_t cc = 0; cc < 4; cc++) {
                    if (((present >> cc) & 1u) && cc != best) {
                        const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < KS);
                        if (room == 0) { w.status = HP_BLOCK_ASSERT; break; }
                        const uint32_t src_lane = (rr + cc) & 31u;
                        uint32_t target = src_lane;
                        if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                        const uint32_t xt = __shfl_sync(HP_FULL_MASK, mt, src_lane), xl = __shfl_sync(HP_FULL_MASK, ml, src_lane);
                        if (lane == target) {
                            const uint64_t k = mk64(xt, xl);
                            my_keys[cnt] = k;
                            pay_store((xl >> 6) & 0xfffffu, cur_h1, cur_h2, cc);
                            if (k < ckey) { ckey = k; cpos = cnt; }
                            cnt++;
                        }
                    }
                }
                if (w.status != HP_BLOCK_OK) break;
            }
            rr += 4;
        }
        // queue minimum now includes the siblings: fold in the smallest sibling key (second best candidate)
        {
            const uint32_t u0 = (best == 0u) ? 0xffffffffu : t0, u1 = (best == 1u) ? 0xffffffffu : t1;
            const uint32_t u2 = (best == 2u) ? 0xffffffffu : t2, u3 = (best == 3u) ? 0xffffffffu : t3;
            const uint32_t m2 = min(min(u0, u1), min(u2, u3));
            if (m2 != 0xffffffffu) {
                const uint32_t c2 = (u0 == m2) ? 0u : (u1 == m2) ? 1u : (u2 == m2) ? 2u : 3u;
                const uint64_t k2 = mk64(m2, ((c2 & 2u) ? lo2 : lo0) + ((c2 & 1u) << 6));
                qmin = k2 < qmin ? k2 : qmin;
            }
        }
        next_idx += nchild;
        // ---- the best child is the new cur ----
        cur_total = tmin;
        cur_lo = ((best & 2u) ? lo2 : lo0) + ((best & 1u) << 6);
        cur_x1 = best & 1u; cur_x2 = (0x9u >> best) & 1u;
        if (!bad_col) {
            cur_h1 |= cur_x1 ? lbit : 0ull;
            cur_h2 |= cur_x2 ? lbit : 0ull;
        }
        lbit <<= 1;
        cur_src = SRC_CACHE;
        heur_p = heur;
        __syncwarp();
        if (kCount) w.ts_rest += clock64() - tq;
    }
    return make_uint2(max_cost, next_expected - 1);
}

template <int K, bool kCount>
__device__ __forceinline__ uint2 sub_solve(const AstarArgs& a, const BlkMeta& m, WarpCtx& w, uint32_t v, uint32_t clip,
                                           uint64_t badwin, uint32_t blk) {
    if constexpr (K == 0) return sub_solve_generic<0, kCount>(a, m, w, v, clip, badwin, blk);
    else {
        // the fast path keeps every stripe's keys in shared memory (4 keys per AoS slot): larger queues (non-default
        // min_queue_size / queue_increment) take the generic path
        if (w.capl > 4u * w.capl_s) return sub_solve_generic<K, kCount>(a, m, w, v, clip, badwin, blk);
        return sub_solve_fast<K, kCount>(a, m, w, v, clip, badwin, blk);
    }
}

// ---- main-queue slab (global memory, private to one warp) -----------------------------------------------------
struct Slab {
    uint64_t* khi;       // [qcap] total << 32 | (0xffffffff - hets)
    uint32_t* kidx;      // [qcap] node index
    uint32_t* klen;      // [qcap] length | identical-haplotypes flag << 31
    uint32_t* kfrozen;   // [qcap]
    uint32_t* krec;      // [qcap] record slot
    uint32_t* freelist;  // [qcap]
    uint32_t* lencnt;    // [hap_words*64 + 2] PQueueHapTracker::length_counts
    uint64_t* pool_hi;   // [qcap] dead pool (entries shorter than min_progress, waiting to be popped and discarded)
    uint32_t* pool_idx;  // [qcap]
    uint64_t* recs;      // [qcap][2*hap_words]
};

__host__ __device__ inline uint64_t slab_bytes_for(uint32_t qcap, uint32_t hap_words) {
    uint64_t b = (uint64_t)qcap * (8 + 4 * 5 + 8 + 4) + (uint64_t)(hap_words * 64 + 2) * 4;
    b = (b + 15) & ~15ull;
    b += (uint64_t)qcap * 2 * hap_words * 8;
    return (b + 255) & ~255ull;
}

__device__ __forceinline__ Slab carve_slab(uint8_t* p, uint32_t qcap, uint32_t hap_words) {
    Slab s;
    s.khi = (uint64_t*)p; p += (uint64_t)qcap * 8;
    s.pool_hi = (uint64_t*)p; p += (uint64_t)qcap * 8;
    s.kidx = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.klen = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.kfrozen = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.krec = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.freelist = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.pool_idx = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.lencnt = (uint32_t*)p; p += (uint64_t)(hap_words * 64 + 2) * 4;
    p = (uint8_t*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    s.recs = (uint64_t*)p;
    return s;
}

// 96-bit main-queue key: hi = total << 32 | (0xffffffff - hets), then the node index.
struct MainKey {
    uint64_t hi;
    uint32_t idx;
};
__device__ __forceinline__ bool key_less(uint64_t hi_a, uint32_t idx_a, uint64_t hi_b, uint32_t idx_b) {
    return hi_a < hi_b || (hi_a == hi_b && idx_a < idx_b);
}
__device__ __forceinline__ MainKey wmin96(uint64_t hi, uint32_t idx) {
    const uint32_t t = (uint32_t)(hi >> 32), h = (uint32_t)hi;
    const uint32_t mt = wmin(t);
    const uint32_t mh = wmin(t == mt ? h : 0xffffffffu);
    const uint32_t mi = wmin((t == mt && h == mh) ? idx : 0xffffffffu);
    MainKey r; r.hi = ((uint64_t)mt << 32) | mh; r.idx = mi;
    return r;
}

// warp-cooperative scan of stripe `owner` (cnt_o entries) -> its minimum and position; valid on every lane
template <class HiAt, class IdxAt>
__device__ __forceinline__ void stripe_min(HiAt khi_at, IdxAt kidx_at, int owner, uint32_t cnt_o, uint32_t lane,
                                           uint64_t& out_hi, uint32_t& out_idx, uint32_t& out_pos) {
    uint64_t bhi = ~0ull; uint32_t bidx = 0xffffffffu, bpos = 0;
    for (uint32_t i = lane; i < cnt_o; i += 32) {
        const uint64_t hi = *khi_at(owner, i);
        const uint32_t idx = *kidx_at(owner, i);
        if (key_less(hi, idx, bhi, bidx)) { bhi = hi; bidx = idx; bpos = i; }
    }
    const MainKey mk = wmin96(bhi, bidx);
    const int wl = __ffs(__ballot_sync(HP_FULL_MASK, bhi == mk.hi && bidx == mk.idx)) - 1;
    out_hi = mk.hi; out_idx = mk.idx;
    out_pos = __shfl_sync(HP_FULL_MASK, bpos, wl);
}

// astar_solver main loop (astar_phaser.rs:480-633) for one block, after the heuristic pre-pass.
// Same structure as sub_solve: cur in registers, dive when the best child beats the queue minimum.  Nodes carry
// full-length haplotype records in the slab (the reference clones h1/h2 per node, :79-82); the best child
// inherits its parent's record in place.
template <int K, bool kCount>
__device__ void main_solve(const AstarArgs& a, const BlkMeta& m, WarpCtx& w, const Slab& s, uint32_t blk,
                           const uint32_t* Hg) {
    const uint32_t lane = w.lane;
    const uint32_t N = m.n_var;
    const uint32_t HW = a.hap_words;
    const uint32_t scap = a.qcap / 32;
    const uint32_t* aoff = a.act_off + m.var_base + blk;
    const uint32_t* aidx = a.act_idx + m.cell_base;
    const uint32_t* col = a.col + m.cell_base;
    const ReadMeta* rmeta = a.rmeta + m.read_base;
    const uint8_t* ign = a.ignored + m.var_base;
    // key (hi, idx) of entry i of a stripe: shared memory for the first mq_cap_s entries, the slab beyond
    uint64_t* const mq_hi = w.mq_hi; uint32_t* const mq_idx = w.mq_idx; const uint32_t mqs = w.mq_cap_s;
    auto khi_at = [&](uint32_t stripe, uint32_t i) -> uint64_t* { return i < mqs ? mq_hi + stripe * mqs + i : s.khi + stripe * scap + i; };
    auto kidx_at = [&](uint32_t stripe, uint32_t i) -> uint32_t* { return i < mqs ? mq_idx + stripe * mqs + i : s.kidx + stripe * scap + i; };
    uint32_t* const mq_len = w.mq_len; uint32_t* const mq_rec = w.mq_rec;
    auto klen_at = [&](uint32_t stripe, uint32_t i) -> uint32_t* { return i < mqs ? mq_len + stripe * mqs + i : s.klen + stripe * scap + i; };
    auto krec_at = [&](uint32_t stripe, uint32_t i) -> uint32_t* { return i < mqs ? mq_rec + stripe * mqs + i : s.krec + stripe * scap + i; };
    // the frozen cost of a queued node lives at s.kfrozen[record slot] (stable while the entry moves inside its stripe)

    // tracker (PQueueHapTracker, :171-231): counts in the slab, totals in registers
    for (uint32_t i = lane; i <= N; i += 32) s.lencnt[i] = 0u;
    __syncwarp();
    uint32_t trk_total = 1, trk_thresh = 0;                              // root counted (:488)
    uint32_t curr_thresh = a.min_queue_size;
    const uint32_t max_queue = 10u * a.min_queue_size;                   // :457
    uint32_t min_progress = 0, next_expected = 0;
    uint64_t num_pruned = 0;
    uint32_t next_idx = 1, rr = 0, qsize = 1;                            // qsize = pqueue.len() (cur included)
    uint32_t free_top = 0, rec_next = 1;                                 // record 0 = root
    uint32_t fs_top = 0;                                                 // entries in the shared-memory free stack
    if (lane == 0) atomicAdd(s.lencnt + 0, 1u);

    // per-lane cached stripe minimum + warp-uniform queue minimum
    uint64_t c_hi = ~0ull; uint32_t c_idx = 0xffffffffu, c_pos = 0, cnt = 0;
    MainKey qmin; qmin.hi = ~0ull; qmin.idx = 0xffffffffu;
    // current top (root, :485-487)
    uint64_t cur_hi = ((uint64_t)Hg[0] << 32) | 0xffffffffull;
    uint32_t cur_idx = 0, cur_len = 0, cur_frozen = 0, cur_rec = 0;
    bool cur_ident = true, have_cur = true;
    uint64_t cur_w1 = 0, cur_w2 = 0;                                     // lane wi: words wi of cur's h1 / h2 (when HW <= 32)
    const bool regs_hap = HW <= 32;
    ExpCache<K> cache;
    cache.first_idx = 0xffffffffu; cache.present = 0;
    long long tm_pop = 0, tm_exp = 0, tm_rest = 0, n_real = 0, t0 = 0;   // counting variant only

    for (;;) {
        if (kCount) t0 = clock64();
        if (!have_cur || key_less(qmin.hi, qmin.idx, cur_hi, cur_idx)) {
            n_real++;
            if (have_cur) {                                              // cur goes back to the queue
                const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < scap);
                if (room == 0) { w.status = HP_BLOCK_QUEUE_OVERFLOW; break; }
                uint32_t target = rr & 31u; rr++;
                if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                if (lane == target) {
                    *khi_at(lane, cnt) = cur_hi; *kidx_at(lane, cnt) = cur_idx; *klen_at(lane, cnt) = cur_len | (cur_ident ? 0x80000000u : 0u);
                    *krec_at(lane, cnt) = cur_rec; s.kfrozen[cur_rec] = cur_frozen;
                    if (key_less(cur_hi, cur_idx, c_hi, c_idx)) { c_hi = cur_hi; c_idx = cur_idx; c_pos = cnt; }
                    cnt++;
                }
                __syncwarp();
            }
            // ---- real pop ----
            if (qmin.hi == ~0ull) { w.status = HP_BLOCK_ASSERT; break; }   // empty queue: the reference panics (:631)
            const int owner = __ffs(__ballot_sync(HP_FULL_MASK, c_hi == qmin.hi && c_idx == qmin.idx)) - 1;
            const uint32_t pos = __shfl_sync(HP_FULL_MASK, c_pos, owner);
            cur_hi = qmin.hi; cur_idx = qmin.idx;
            const uint32_t lenf = *klen_at(owner, pos);
            cur_len = lenf & 0x7fffffffu; cur_ident = (lenf >> 31) != 0;
            cur_rec = *krec_at(owner, pos);
            if (cur_len >= min_progress) {                                // a pruned node never needs its payload
                cur_frozen = s.kfrozen[cur_rec];
                if (regs_hap) {
                    const uint64_t* r = s.recs + (uint64_t)cur_rec * 2 * HW;
                    const bool own = lane < ((cur_len + 63) >> 6);
                    cur_w1 = own ? r[lane] : 0ull; cur_w2 = own ? r[HW + lane] : 0ull;
                }
            }
            const uint32_t cnt_o = __shfl_sync(HP_FULL_MASK, cnt, owner) - 1;
            __syncwarp();
            if ((int)lane == owner) {
                cnt--;
                if (pos != cnt) {
                    *khi_at(owner, pos) = *khi_at(owner, cnt); *kidx_at(owner, pos) = *kidx_at(owner, cnt);
                    *klen_at(owner, pos) = *klen_at(owner, cnt); *krec_at(owner, pos) = *krec_at(owner, cnt);
                }
            }
            __syncwarp();
            {
                uint64_t nhi; uint32_t nidx, npos;
                stripe_min(khi_at, kidx_at, owner, cnt_o, lane, nhi, nidx, npos);
                if ((int)lane == owner) { c_hi = nhi; c_idx = nidx; c_pos = npos; }
            }
            qmin = wmin96(c_hi, c_idx);
            have_cur = true;
        }
        if (kCount) { const long long t1 = clock64(); tm_pop += t1 - t0; t0 = t1; }
        // ---- cur is the top ----
        const uint32_t L = cur_len;
        const uint32_t total = (uint32_t)(cur_hi >> 32);
        if (L >= N) break;                                                // :492
        // pop bookkeeping: hap_tracker.remove_hap (:495)
        qsize--;
        if (lane == 0) atomicAdd(s.lencnt + L, 0xffffffffu);
        if (L >= trk_thresh) trk_total--;
        w.pops++;
        if (L == next_expected) {                                         // :497-504
            next_expected++;
            if (num_pruned == 0) curr_thresh += a.queue_increment;
        }
        if (L < min_progress) {                                           // :507-515
            if (num_pruned == 0) curr_thresh = a.min_queue_size;
            num_pruned++;
            if (fs_top < kFreeStack) { if (lane == 0) w.free_stack[fs_top] = cur_rec; fs_top++; }
            else { if (lane == 0) s.freelist[free_top] = cur_rec; free_top++; }
            have_cur = false;
            __syncwarp();
            continue;
        }

        // ---- expand ----
        const uint32_t p = L;
        const bool bad_col = __ldg(ign + p) != 0;
        const uint32_t heur = Hg[p + 1];
        const uint32_t hets = 0xffffffffu - (uint32_t)cur_hi;
        const uint64_t* prow = s.recs + (uint64_t)cur_rec * 2 * HW;
        const uint32_t o0 = __ldg(aoff + p), o1 = __ldg(aoff + p + 1);
        const int nwords = (int)((L + 63) >> 6);                          // words of the parent that hold set bits
        int src = SRC_PLANES;
        uint32_t x1 = 0, x2 = 0;
        if (L == 0) src = SRC_ROOT;
        else if (K > 0 && cur_idx - cache.first_idx < (uint32_t)__popc(cache.present)) {
            src = SRC_CACHE;
            const uint32_t cslot = slot_of_ordinal(cache.present, cur_idx - cache.first_idx);
            x1 = cslot & 1u; x2 = (0x9u >> cslot) & 1u;
        }
        auto hap = [&](int which, int i0) -> uint64_t {                   // 64 bits from haplotype position i0 >= 0
            const uint64_t* hw = prow + (which ? HW : 0);
            const int wi = i0 >> 6, sh = i0 & 63;
            uint64_t x = (wi < nwords) ? (hw[wi] >> sh) : 0ull;
            if (sh && wi + 1 < nwords) x |= hw[wi + 1] << (64 - sh);
            return x;
        };
        uint32_t tot[4], fro[4];
        uint64_t cells = 0;
        expand<K, kCount>(a, m, lane, col + o0, aidx + o0, rmeta, o1 - o0, p, 0, (int)L, bad_col, cur_ident, src, x1, x2,
                          hap, cache, tot, fro, cells);
        const uint32_t present = present_mask(bad_col, cur_ident);
        const uint32_t nchild = __popc(present);
        cache.first_idx = next_idx; cache.present = present;
        if (kCount) { const long long t1 = clock64(); tm_exp += t1 - t0; t0 = t1; }
        if (kCount) { w.evals += nchild; w.sum_lp += (uint64_t)nchild * L; w.cells += cells * nchild; }
        if (bad_col && cur_frozen + tot[2] + heur != total) { w.status = HP_BLOCK_ASSERT; break; }   // :529
        if (next_idx > 0xfffffff0u) { w.status = HP_BLOCK_INDEX_EXHAUSTED; break; }
        if (qsize + nchild + 64 > a.qcap) { w.status = HP_BLOCK_QUEUE_OVERFLOW; break; }

        // keys, creation order
        uint64_t khi[4]; uint32_t kix[4];
        uint32_t best = 2;
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
            const uint32_t ord = __popc(present & ((1u << c) - 1u));
            const uint32_t ch = hets + ((c < 2 && !bad_col) ? 1u : 0u);
            const bool pr = (present >> c) & 1u;
            khi[c] = pr ? (((uint64_t)(cur_frozen + tot[c] + heur) << 32) | (uint64_t)(0xffffffffu - ch)) : ~0ull;
            kix[c] = pr ? next_idx + ord : 0xffffffffu;
        }
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) if (key_less(khi[c], kix[c], khi[best], kix[best])) best = c;

        // ---- records: siblings get copies of the parent's words + their allele bit; the best child then takes the
        //      parent's record in place ----
        uint32_t crec[4];
        {
            const uint32_t nsib = nchild - 1;
            uint32_t mine = 0;
            const uint32_t from_fs = min(nsib, fs_top);                   // shared-memory stack first, then the slab list, then fresh
            const uint32_t from_gl = min(nsib - from_fs, free_top);
            if (lane < nsib) {
                if (lane < from_fs) mine = w.free_stack[fs_top - 1 - lane];
                else if (lane - from_fs < from_gl) mine = s.freelist[free_top - 1 - (lane - from_fs)];
                else mine = rec_next + (lane - from_fs - from_gl);
            }
            uint32_t ordinal = 0;
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) {
                const bool sib = ((present >> c) & 1u) && c != best;
                crec[c] = sib ? __shfl_sync(HP_FULL_MASK, mine, ordinal & 31u) : cur_rec;
                if (sib) ordinal++;
            }
            fs_top -= from_fs; free_top -= from_gl; rec_next += nsib - from_fs - from_gl;
        }
        {
            const int wl = (int)(L >> 6);
            const uint64_t bit = bad_col ? 0ull : (1ull << (L & 63));
            for (int wi = lane; wi <= wl; wi += 32) {
                uint64_t w1 = 0, w2 = 0;
                if (regs_hap) { w1 = cur_w1; w2 = cur_w2; }
                else if (wi < nwords) { w1 = prow[wi]; w2 = prow[HW + wi]; }
#pragma unroll
                for (uint32_t c = 0; c < 4; c++) {
                    if (((present >> c) & 1u) && (c != best || wi == wl)) {   // best: only the word that changes
                        uint64_t* crow = s.recs + (uint64_t)crec[c] * 2 * HW;
                        crow[wi] = (wi == wl && (c == 1u || c == 3u)) ? (w1 | bit) : w1;
                        crow[HW + wi] = (wi == wl && (c == 0u || c == 3u)) ? (w2 | bit) : w2;
                    }
                }
            }
        }
        // ---- push the siblings ----
        {
            const uint32_t c = (lane - rr) & 31u;
            const bool mine = c < 4u && ((present >> c) & 1u) && c != best;
            const uint32_t fullmask = __ballot_sync(HP_FULL_MASK, mine && cnt >= scap);
            const uint32_t cident = (cur_ident && c >= 2u) ? 0x80000000u : 0u;
            if (fullmask == 0) {
                if (mine) {
                    const uint64_t hi = (c == 0) ? khi[0] : (c == 1) ? khi[1] : (c == 2) ? khi[2] : khi[3];
                    const uint32_t ix = (c == 0) ? kix[0] : (c == 1) ? kix[1] : (c == 2) ? kix[2] : kix[3];
                    const uint32_t fr = cur_frozen + ((c == 0) ? fro[0] : (c == 1) ? fro[1] : (c == 2) ? fro[2] : fro[3]);
                    const uint32_t rc = (c == 0) ? crec[0] : (c == 1) ? crec[1] : (c == 2) ? crec[2] : crec[3];
                    *khi_at(lane, cnt) = hi; *kidx_at(lane, cnt) = ix; *klen_at(lane, cnt) = (L + 1) | cident;
                    *krec_at(lane, cnt) = rc; s.kfrozen[rc] = fr;
                    if (key_less(hi, ix, c_hi, c_idx)) { c_hi = hi; c_idx = ix; c_pos = cnt; }
                    cnt++;
                }
            } else {
#pragma unroll
                for (uint32_t cc = 0; cc < 4; cc++) {
                    if (((present >> cc) & 1u) && cc != best) {
                        const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < scap);
                        if (room == 0) { w.status = HP_BLOCK_QUEUE_OVERFLOW; break; }
                        uint32_t target = (rr + cc) & 31u;
                        if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                        if (lane == target) {
                            *khi_at(lane, cnt) = khi[cc]; *kidx_at(lane, cnt) = kix[cc];
                            *klen_at(lane, cnt) = (L + 1) | ((cur_ident && cc >= 2u) ? 0x80000000u : 0u);
                            *krec_at(lane, cnt) = crec[cc]; s.kfrozen[crec[cc]] = cur_frozen + fro[cc];
                            if (key_less(khi[cc], kix[cc], c_hi, c_idx)) { c_hi = khi[cc]; c_idx = kix[cc]; c_pos = cnt; }
                            cnt++;
                        }
                    }
                }
            }
            rr += 4;
#pragma unroll
            for (uint32_t cc = 0; cc < 4; cc++)
                if (cc != best && key_less(khi[cc], kix[cc], qmin.hi, qmin.idx)) { qmin.hi = khi[cc]; qmin.idx = kix[cc]; }
        }
        if (w.status != HP_BLOCK_OK) break;
        next_idx += nchild; qsize += nchild;
        // tracker.add_hap(L+1) x nchild (:531, :558)
        if (lane == 0) atomicAdd(s.lencnt + L + 1, nchild);
        if (L + 1 >= trk_thresh) trk_total += nchild;
        // ---- the best child is the new cur (record inherited in place) ----
        {
            const uint32_t bf = (best == 0) ? fro[0] : (best == 1) ? fro[1] : (best == 2) ? fro[2] : fro[3];
            cur_hi = khi[best]; cur_idx = kix[best]; cur_len = L + 1; cur_frozen += bf;
            cur_ident = cur_ident && best >= 2u;
            if (regs_hap && !bad_col && lane == (L >> 6)) {
                const uint64_t b = 1ull << (L & 63);
                if (best == 1u || best == 3u) cur_w1 |= b;
                if (best == 0u || best == 3u) cur_w2 |= b;
            }
        }
        __syncwarp();

        // ---- pruning bookkeeping (:564-585) ----
        while (trk_total > curr_thresh && min_progress < next_expected) {
            min_progress++;
            uint32_t dropped = 0;
            if (lane == 0) dropped = atomicAdd(s.lencnt + min_progress - 1, 0u);
            dropped = __shfl_sync(HP_FULL_MASK, dropped, 0);
            trk_total -= dropped; trk_thresh = min_progress;
            if (qsize > max_queue) {
                // "full prune": every queued entry shorter than min_progress gets the cleared priority (cost 0);
                // cur has length >= next_expected >= min_progress and is never affected
                c_hi = ~0ull; c_idx = 0xffffffffu; c_pos = 0;
                for (uint32_t i = 0; i < cnt; i++) {
                    uint64_t hi = *khi_at(lane, i);
                    const uint32_t ix = *kidx_at(lane, i);
                    if ((*klen_at(lane, i) & 0x7fffffffu) < min_progress) { hi &= 0xffffffffull; *khi_at(lane, i) = hi; }
                    if (key_less(hi, ix, c_hi, c_idx)) { c_hi = hi; c_idx = ix; c_pos = i; }
                }
                __syncwarp();
                qmin = wmin96(c_hi, c_idx);
            }
        }
        if (kCount) { const long long t1 = clock64(); tm_rest += t1 - t0; t0 = t1; }
    }
    if (kCount && a.dbg_cycles && lane == 0) {
        uint64_t* d = a.dbg_cycles + 16ull * blk;
        d[8] = tm_pop; d[9] = tm_exp; d[10] = tm_rest; d[13] = n_real; d[14] = num_pruned; d[15] = qsize;
    }
    if (w.status != HP_BLOCK_OK) return;

    // ---- final node (:588-628) ----
    const uint32_t top_total = (uint32_t)(cur_hi >> 32);
    const uint64_t* frow = s.recs + (uint64_t)cur_rec * 2 * HW;
    uint32_t phased = 0, phased_snv = 0, skipped = 0;
    const uint8_t* snv = a.is_snv + m.var_base;
    for (uint32_t i = lane; i < N; i += 32) {
        uint32_t b1 = (uint32_t)(frow[i >> 6] >> (i & 63)) & 1u;
        uint32_t b2 = (uint32_t)(frow[HW + (i >> 6)] >> (i & 63)) & 1u;
        if (__ldg(ign + i)) { b1 = 2; b2 = 2; skipped++; }
        else if (b1 != b2) { phased++; if (__ldg(snv + i)) phased_snv++; }
        a.out_h1[m.var_base + i] = (uint8_t)b1;
        a.out_h2[m.var_base + i] = (uint8_t)b2;
    }
    phased = wsum(phased); phased_snv = wsum(phased_snv); skipped = wsum(skipped);
    if (lane == 0) {
        uint64_t* st = a.out_stats + (uint64_t)blk * 7;
        st[0] = num_pruned;
        st[1] = Hg[0];
        st[2] = top_total;
        st[3] = phased; st[4] = phased_snv; st[5] = N - phased - skipped; st[6] = skipped;
    }
    if (top_total < Hg[0]) w.status = HP_BLOCK_ASSERT;                     // phase_stats.rs:163
}

// astar_solver main loop, register-resident fast path (K >= 1): the structure of sub_solve_fast (cur in registers,
// packed-delta totals, child vectors pre-permuted into the next column's slot order, prefetched column records)
// with the main loop's 96-bit keys, haplotype records, tracker and pruning rules (astar_phaser.rs:480-633).
template <int K, bool kCount>
__device__ void main_solve_fast(const AstarArgs& a, const BlkMeta& m, WarpCtx& w, const Slab& s, uint32_t blk,
                                const uint32_t* Hg) {
    const uint32_t lane = w.lane;
    const uint32_t N = m.n_var;
    const uint32_t HW = a.hap_words;
    const uint32_t scap = a.qcap / 32;
    const uint32_t* aoff = a.act_off + m.var_base + blk;
    const uint32_t* aidx = a.act_idx + m.cell_base;
    const uint32_t* col = a.col + m.cell_base;
    const ReadMeta* rmeta = a.rmeta + m.read_base;
    const uint8_t* ign = a.ignored + m.var_base;
    constexpr uint32_t kEmpty = 0xffff0000u;
    uint64_t* const mq_hi = w.mq_hi; uint32_t* const mq_idx = w.mq_idx; const uint32_t mqs = w.mq_cap_s;
    uint32_t* const mq_len = w.mq_len; uint32_t* const mq_rec = w.mq_rec;
    auto khi_at = [&](uint32_t stripe, uint32_t i) -> uint64_t* { return i < mqs ? mq_hi + stripe * mqs + i : s.khi + stripe * scap + i; };
    auto kidx_at = [&](uint32_t stripe, uint32_t i) -> uint32_t* { return i < mqs ? mq_idx + stripe * mqs + i : s.kidx + stripe * scap + i; };
    auto klen_at = [&](uint32_t stripe, uint32_t i) -> uint32_t* { return i < mqs ? mq_len + stripe * mqs + i : s.klen + stripe * scap + i; };
    auto krec_at = [&](uint32_t stripe, uint32_t i) -> uint32_t* { return i < mqs ? mq_rec + stripe * mqs + i : s.krec + stripe * scap + i; };

    // tracker length counts: only lane 0 touches them (plain read-modify-write, program order)
    uint32_t* const lc = (N + 1 <= w.lencnt_cap_s) ? w.lencnt_s : s.lencnt;
    for (uint32_t i = lane; i <= N; i += 32) lc[i] = 0u;
    __syncwarp();
    // dead pool: entries that fell below min_progress leave their stripe at once (keys only, records freed); they are
    // "popped" (counted) in key order as the live pops pass them
    uint64_t* const pool_hi = s.pool_hi; uint32_t* const pool_idx = s.pool_idx;
    uint32_t pool_n = 0;
    MainKey pool_min; pool_min.hi = ~0ull; pool_min.idx = 0xffffffffu;
    // A pool entry with key k is popped (and discarded) by the reference right before the first live node with a larger
    // key that is processed after the entry entered the pool.  All entries of the pool share one epoch (it restarts at
    // every evaluation), so "popped by now" == k < pool_max, the largest live key processed since the epoch began: the
    // count is only needed where the reference looks at it (first prune, full-prune test, final statistics).
    MainKey pool_max; pool_max.hi = 0ull; pool_max.idx = 0u;
    uint32_t trk_total = 1, trk_thresh = 0;                              // root counted (:488)
    uint32_t curr_thresh = a.min_queue_size;
    const uint32_t max_queue = 10u * a.min_queue_size;                   // :457
    uint32_t min_progress = 0, next_expected = 0;
    uint64_t num_pruned = 0;
    uint32_t next_idx = 1, rr = 0, qsize = 1;                            // qsize = pqueue.len() (cur included)
    uint32_t free_top = 0, rec_next = 1, fs_top = 0;                     // record 0 = root
    if (lane == 0) lc[0] = 1u;

    uint64_t c_hi = ~0ull; uint32_t c_idx = 0xffffffffu, c_pos = 0, cnt = 0;
    MainKey qmin; qmin.hi = ~0ull; qmin.idx = 0xffffffffu;
    // current top (root, :485-487)
    uint32_t cur_total = Hg[0], cur_nh = 0xffffffffu, cur_idx = 0, cur_len = 0, cur_rec = 0;
    bool cur_ident = true, have_cur = true;
    int cur_src = SRC_ROOT;
    uint32_t cur_x1 = 0, cur_x2 = 0;
    uint32_t heur_p = cur_total;                                         // heuristic term inside cur_total
    // H[p+1] and ignored[p] of cur's column p are loaded one expansion ahead (L2 latency off the dive's critical path)
    uint32_t heur_c = Hg[1 <= N ? 1 : 0];
    bool bad_c = N > 0 && __ldg(ign) != 0;
    uint64_t cur_w1 = 0, cur_w2 = 0;                                     // lane wi: words wi of cur's h1 / h2 (HW <= 32)
    const bool regs_hap = HW <= 32;
    uint32_t nA0[K], nA1[K], nB0[K], nB1[K], nW[K];
    uint32_t cache_first = 0xffffffffu, cache_present = 0;
    const uint32_t* col_lane = col + lane;
    uint32_t o_p = __ldg(aoff + 0), o_p1 = __ldg(aoff + 1), o_p2 = (2 <= N) ? __ldg(aoff + 2) : o_p1;
    uint32_t colc[K], coln[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        nA0[k] = nA1[k] = nB0[k] = nB1[k] = nW[k] = 0;
        colc[k] = (lane + 32u * k < o_p1 - o_p) ? __ldg(col + o_p + lane + 32u * k) : kEmpty;
        coln[k] = (1 < N) ? __ldg(col_lane + o_p1 + 32u * k) : kEmpty;
    }

    // removes entry `pos` of stripe `owner` (swap with its last entry), then the warp rescans that stripe; `left` = entries
    // left in the stripe.  Stripes that fit the shared-memory part take the direct path.
    auto remove_rescan = [&](int owner, uint32_t pos, uint32_t left) {
        const bool in_smem = left < mqs;                                   // positions 0..left all in shared memory
        if ((int)lane == owner) {
            cnt--;
            if (pos != cnt) {
                if (in_smem) {
                    const uint32_t o = lane * mqs;
                    mq_hi[o + pos] = mq_hi[o + cnt]; mq_idx[o + pos] = mq_idx[o + cnt]; mq_len[o + pos] = mq_len[o + cnt]; mq_rec[o + pos] = mq_rec[o + cnt];
                } else {
                    *khi_at(owner, pos) = *khi_at(owner, cnt); *kidx_at(owner, pos) = *kidx_at(owner, cnt);
                    *klen_at(owner, pos) = *klen_at(owner, cnt); *krec_at(owner, pos) = *krec_at(owner, cnt);
                }
            }
        }
        __syncwarp();
        uint64_t nhi; uint32_t nidx, npos;
        if (in_smem) {
            uint64_t bhi = ~0ull; uint32_t bidx = 0xffffffffu, bpos = 0;
            const uint32_t o = (uint32_t)owner * mqs;
            for (uint32_t q = lane; q < left; q += 32) {
                const uint64_t hi = mq_hi[o + q]; const uint32_t ix = mq_idx[o + q];
                if (key_less(hi, ix, bhi, bidx)) { bhi = hi; bidx = ix; bpos = q; }
            }
            const MainKey mk = wmin96(bhi, bidx);
            const int wl2 = __ffs(__ballot_sync(HP_FULL_MASK, bhi == mk.hi && bidx == mk.idx)) - 1;
            nhi = mk.hi; nidx = mk.idx; npos = __shfl_sync(HP_FULL_MASK, bpos, wl2);
        } else stripe_min(khi_at, kidx_at, owner, left, lane, nhi, nidx, npos);
        if ((int)lane == owner) { c_hi = nhi; c_idx = nidx; c_pos = npos; }
        qmin = wmin96(c_hi, c_idx);
    };
    auto eval_pool = [&]() {
        // counts and removes the pool entries popped so far (key < pool_max), restarts the epoch
        uint32_t base = 0;
        uint64_t m_hi = ~0ull; uint32_t m_idx = 0xffffffffu;
        for (uint32_t c0 = 0; c0 < pool_n; c0 += 32) {
            const uint32_t i = c0 + lane;
            const bool in = i < pool_n;
            uint64_t hi = ~0ull; uint32_t ix = 0xffffffffu;
            if (in) { hi = pool_hi[i]; ix = pool_idx[i]; }
            const bool keep = in && !key_less(hi, ix, pool_max.hi, pool_max.idx);
            const uint32_t km = __ballot_sync(HP_FULL_MASK, keep);
            if (keep) {
                const uint32_t d = base + __popc(km & ((1u << lane) - 1u));
                pool_hi[d] = hi; pool_idx[d] = ix;
                if (key_less(hi, ix, m_hi, m_idx)) { m_hi = hi; m_idx = ix; }
            }
            base += __popc(km);
        }
        __syncwarp();
        const uint32_t k = pool_n - base;
        pool_n = base;
        pool_min = wmin96(m_hi, m_idx);
        pool_max.hi = 0ull; pool_max.idx = 0u;
        if (k != 0) {
            if (num_pruned == 0) curr_thresh = a.min_queue_size;          // :508-510
            num_pruned += k; qsize -= k; w.pops += k;
        }
    };
#ifdef HP_DBG_RING_AGE
    uint32_t hist_first = 0xffffffffu, hist_n = 0;     // lane l: first child index of the expansion made (l+1) expansions ago ... ring of 32
    unsigned long long age_le[4] = {0, 0, 0, 0};        // plane-rescored pops whose parent expansion is <= 4 / 8 / 16 / 32 expansions old
#endif
    long long tm_pop = 0, tm_exp = 0, tm_rest = 0, tm_planes = 0, n_real = 0, n_planes = 0, n_swept = 0, tq0 = 0, tm_vec = 0, tm_rec = 0, tm_push = 0, tm_dead = 0;   // counting variant only
    for (;;) {
        if (kCount) tq0 = clock64();
        if (!have_cur || key_less(qmin.hi, qmin.idx, ((uint64_t)cur_total << 32) | cur_nh, cur_idx)) {
            if (kCount) n_real++;
            if (qmin.hi == ~0ull) { w.status = HP_BLOCK_ASSERT; break; }   // empty queue (only without cur): the reference panics (:631)
            const int owner = __ffs(__ballot_sync(HP_FULL_MASK, c_hi == qmin.hi && c_idx == qmin.idx)) - 1;
            const uint32_t pos = __shfl_sync(HP_FULL_MASK, c_pos, owner);
            const uint32_t lenf = *klen_at(owner, pos);
            if ((lenf & 0x7fffffffu) < min_progress) {
                // ---- the top entry is dead: the reference pops and discards it (:507-515).  cur stays where it is (in
                //      registers): the discard happens before cur is looked at again, exactly as in the reference's order ----
                const uint32_t drec = *krec_at(owner, pos);
                const uint32_t cnt_d = __shfl_sync(HP_FULL_MASK, cnt, owner) - 1;
                __syncwarp();
                remove_rescan(owner, pos, cnt_d);
                qsize--;
                if (lane == 0) atomicAdd(lc + (lenf & 0x7fffffffu), 0xffffffffu);   // hap_tracker.remove_hap (:495); below the threshold
                w.pops++;
                if (num_pruned == 0) curr_thresh = a.min_queue_size;        // :508-510
                num_pruned++;
                if (fs_top < kFreeStack) { if (lane == 0) w.free_stack[fs_top] = drec; fs_top++; }
                else { if (lane == 0) s.freelist[free_top] = drec; free_top++; }
                __syncwarp();
                if (kCount) { const long long t1 = clock64(); tm_pop += t1 - tq0; tm_dead += t1 - tq0; }
                continue;
            }
            if (have_cur) {                                              // cur goes back to the queue (appended: pos stays valid)
                const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < scap);
                if (room == 0) { w.status = HP_BLOCK_QUEUE_OVERFLOW; break; }
                uint32_t target = rr & 31u; rr++;
                if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                if (lane == target) {
                    const uint64_t hi = ((uint64_t)cur_total << 32) | cur_nh;
                    const uint32_t lenf_c = cur_len | (cur_ident ? 0x80000000u : 0u);
                    if (cnt < mqs) { const uint32_t o = lane * mqs + cnt; mq_hi[o] = hi; mq_idx[o] = cur_idx; mq_len[o] = lenf_c; mq_rec[o] = cur_rec; }
                    else { *khi_at(lane, cnt) = hi; *kidx_at(lane, cnt) = cur_idx; *klen_at(lane, cnt) = lenf_c; *krec_at(lane, cnt) = cur_rec; }
                    if (key_less(hi, cur_idx, c_hi, c_idx)) { c_hi = hi; c_idx = cur_idx; c_pos = cnt; }
                    cnt++;
                }
                __syncwarp();
            }
            // ---- real pop of a live entry ----
            cur_total = (uint32_t)(qmin.hi >> 32); cur_nh = (uint32_t)qmin.hi; cur_idx = qmin.idx;
            cur_len = lenf & 0x7fffffffu; cur_ident = (lenf >> 31) != 0;
            cur_rec = *krec_at(owner, pos);
            if (cur_len >= min_progress && cur_len < N) {                 // pruned / final nodes never need the payload
                if (regs_hap) {
                    const uint64_t* r = s.recs + (uint64_t)cur_rec * 2 * HW;
                    const bool own = lane < ((cur_len + 63) >> 6);
                    cur_w1 = own ? r[lane] : 0ull; cur_w2 = own ? r[HW + lane] : 0ull;
                }
                // re-seat the column state
                const uint32_t d = cur_idx - cache_first;
                if (cur_len == 0u) cur_src = SRC_ROOT;
                else if (d < (uint32_t)__popc(cache_present)) {
                    const uint32_t cs = slot_of_ordinal(cache_present, d);
                    cur_src = SRC_CACHE; cur_x1 = cs & 1u; cur_x2 = (0x9u >> cs) & 1u;
                } else cur_src = SRC_PLANES;
                heur_p = Hg[cur_len];
                heur_c = Hg[cur_len + 1];
                bad_c = __ldg(ign + cur_len) != 0;
                o_p = __ldg(aoff + cur_len); o_p1 = __ldg(aoff + cur_len + 1);
                o_p2 = (cur_len + 2 <= N) ? __ldg(aoff + cur_len + 2) : o_p1;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    colc[k] = (lane + 32u * k < o_p1 - o_p) ? __ldg(col + o_p + lane + 32u * k) : kEmpty;
                    coln[k] = (cur_len + 1 < N) ? __ldg(col_lane + o_p1 + 32u * k) : kEmpty;
                }
            }
            const uint32_t cnt_o = __shfl_sync(HP_FULL_MASK, cnt, owner) - 1;
            __syncwarp();
            remove_rescan(owner, pos, cnt_o);
            have_cur = true;
        }
        // ---- cur is the top of the live entries: every pool entry with a smaller key has been popped before it (:507-515) ----
        {
            const uint64_t k_hi = ((uint64_t)cur_total << 32) | cur_nh;
            if (key_less(pool_max.hi, pool_max.idx, k_hi, cur_idx)) { pool_max.hi = k_hi; pool_max.idx = cur_idx; }
            // the first prune resets the queue threshold (:508-510): until it happened the count must be exact in time
            if (num_pruned == 0 && pool_n != 0 && key_less(pool_min.hi, pool_min.idx, k_hi, cur_idx)) eval_pool();
        }
        const uint32_t L = cur_len;
        if (L >= N) break;                                                // :492
        qsize--;
        if (lane == 0) atomicAdd(lc + L, 0xffffffffu);                    // hap_tracker.remove_hap (:495); no result needed
        if (L >= trk_thresh) trk_total--;
        w.pops++;
        if (L == next_expected) {                                         // :497-504
            next_expected++;
            if (num_pruned == 0) curr_thresh += a.queue_increment;
        }
        if (L < min_progress) {                                           // :507-515
            if (num_pruned == 0) curr_thresh = a.min_queue_size;
            num_pruned++;
            if (fs_top < kFreeStack) { if (lane == 0) w.free_stack[fs_top] = cur_rec; fs_top++; }
            else { if (lane == 0) s.freelist[free_top] = cur_rec; free_top++; }
            have_cur = false;
            __syncwarp();
            if (kCount) tm_pop += clock64() - tq0;
            continue;
        }
        if (kCount) { const long long t1 = clock64(); tm_pop += t1 - tq0; tq0 = t1; }

        // ---- expand column p = L ----
        const uint32_t p = L;
        const uint32_t a_cur = o_p1 - o_p;
        const uint32_t o_p3 = (p + 3 <= N) ? __ldg(aoff + p + 3) : o_p2;
        uint32_t colnn[K];
#pragma unroll
        for (int k = 0; k < K; k++) colnn[k] = (p + 2 < N) ? __ldg(col_lane + o_p2 + 32u * k) : kEmpty;
        const uint32_t heur = heur_c;
        const bool bad_col = bad_c;
        const bool ident = cur_ident;
        if (p + 1 < N) { heur_c = Hg[p + 2]; bad_c = __ldg(ign + p + 1) != 0; }

        uint32_t s1[K], s2[K], wv[K];
        if (cur_src == SRC_CACHE) {
#pragma unroll
            for (int k = 0; k < K; k++) { s1[k] = cur_x1 ? nA1[k] : nA0[k]; s2[k] = cur_x2 ? nB1[k] : nB0[k]; wv[k] = nW[k]; }
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) { s1[k] = 0; s2[k] = 0; wv[k] = 0; }
            if (cur_src == SRC_PLANES) {
                const uint64_t* prow = s.recs + (uint64_t)cur_rec * 2 * HW;
                const int nwords = (int)((L + 63) >> 6);
                auto hap = [&](int which, int i0) -> uint64_t {           // 64 bits from haplotype position i0 >= 0
                    const uint64_t* hw = prow + (which ? HW : 0);
                    const int wi = i0 >> 6, sh = i0 & 63;
                    uint64_t x = (wi < nwords) ? (hw[wi] >> sh) : 0ull;
                    if (sh && wi + 1 < nwords) x |= hw[wi + 1] << (64 - sh);
                    return x;
                };
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const uint32_t j = lane + 32u * k;
                    if (j < a_cur) {
                        const ReadMeta rm = rmeta[__ldg(aidx + o_p + j)];
                        score_planes(a, m, rm, -(int)rm.start, (int)L, hap, s1[k], s2[k]);
                        wv[k] = p - rm.start;
                    }
                }
            }
        }
        uint32_t A0[K], A1[K], B0[K], B1[K];
        uint32_t pk0 = 0, pk1 = 0;
        uint64_t cells = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const uint32_t c = colc[k];
            const uint32_t q = bad_col ? 0u : (c & 0xffu);
            const uint32_t al = (c >> 8) & 3u;
            const uint32_t q0 = (al != 0u) ? q : 0u, q1 = (al != 1u) ? q : 0u;
            A0[k] = s1[k] + q0; A1[k] = s1[k] + q1; B0[k] = s2[k] + q0; B1[k] = s2[k] + q1;
            const uint32_t base = min(s1[k], s2[k]);
            const uint32_t c01 = min(A0[k], B1[k]), c10 = min(A1[k], B0[k]), c00 = min(A0[k], B0[k]), c11 = min(A1[k], B1[k]);
            pk0 += (c01 - base) | ((c10 - base) << 16);
            pk1 += (c00 - base) | ((c11 - base) << 16);
            if (kCount) { wv[k] += 1; if (lane + 32u * k < a_cur) cells += wv[k]; }
        }
        const uint32_t r0 = wsum(pk0), r1 = wsum(pk1);
        const uint32_t present = present_mask(bad_col, ident);
        const uint32_t nchild = bad_col ? 1u : (ident ? 3u : 4u);
        if (kCount) {
            const long long t1 = clock64(); tm_exp += t1 - tq0;
            if (cur_src == SRC_PLANES) { tm_planes += t1 - tq0; n_planes++; }
#ifdef HP_DBG_RING_AGE
            if (cur_src == SRC_PLANES) {
                // which of the last 32 expansions created cur?  (children of an expansion have consecutive indices)
                const uint32_t dd = cur_idx - hist_first;
                const uint32_t hit = __ballot_sync(HP_FULL_MASK, hist_first != 0xffffffffu && dd < 4u);
                if (hit) {
                    const uint32_t slot = __ffs(hit) - 1;                    // lane = ring slot; age = (hist_n - 1 - slot) mod 32 + 1
                    const uint32_t age = ((hist_n - 1u - slot) & 31u) + 1u;
                    if (age <= 4) age_le[0]++; if (age <= 8) age_le[1]++; if (age <= 16) age_le[2]++; age_le[3]++;
                }
            }
            if (lane == (hist_n & 31u)) hist_first = next_idx;               // this expansion's first child index
            hist_n++;
#endif
            tq0 = t1;
        }
        if (kCount) { w.evals += nchild; w.sum_lp += (uint64_t)nchild * L; w.cells += cells * nchild; }
        if (next_idx > 0xfffffff0u) { w.status = HP_BLOCK_INDEX_EXHAUSTED; break; }
        if (qsize + nchild + 64 > a.qcap) { w.status = HP_BLOCK_QUEUE_OVERFLOW; break; }

        // candidate keys: total_c = cur_total - H[p] + H[p+1] + delta_c; hi_c = total_c << 32 | (~hets_c); idx in creation
        // order.  (~hets) is smaller for the heterozygous candidates 0,1, so ties on the total go to the lowest slot.
        const uint32_t tb = cur_total - heur_p + heur;
        const uint32_t t0 = bad_col ? 0xffffffffu : tb + (r0 & 0xffffu);
        const uint32_t t1 = (bad_col || ident) ? 0xffffffffu : tb + (r0 >> 16);
        const uint32_t t2 = tb + (r1 & 0xffffu);
        const uint32_t t3 = bad_col ? 0xffffffffu : tb + (r1 >> 16);
        if (bad_col && t2 != cur_total) { w.status = HP_BLOCK_ASSERT; break; }              // :529
        const uint32_t nh_het = cur_nh - 1u;                               // hets + 1
        const uint32_t i0 = next_idx, i1 = next_idx + 1u;
        const uint32_t i2 = next_idx + (bad_col ? 0u : (ident ? 1u : 2u)), i3 = i2 + 1u;
        const uint32_t tmin = min(min(t0, t1), min(t2, t3));
        const uint32_t best = (t0 == tmin) ? 0u : (t1 == tmin) ? 1u : (t2 == tmin) ? 2u : 3u;

        // ---- children vectors into the next column's slot order ----
        {
            const uint32_t a_next = o_p2 - o_p1;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const uint32_t cn = (lane + 32u * k < a_next) ? coln[k] : kEmpty;
                const uint32_t carry = cn >> 16, sl = carry & 31u;
                uint32_t g0 = __shfl_sync(HP_FULL_MASK, A0[0], sl), g1 = __shfl_sync(HP_FULL_MASK, A1[0], sl);
                uint32_t g2 = __shfl_sync(HP_FULL_MASK, B0[0], sl), g3 = __shfl_sync(HP_FULL_MASK, B1[0], sl);
                uint32_t gw = kCount ? __shfl_sync(HP_FULL_MASK, wv[0], sl) : 0u;
#pragma unroll
                for (int kk = 1; kk < K; kk++) {
                    const uint32_t u0 = __shfl_sync(HP_FULL_MASK, A0[kk], sl), u1 = __shfl_sync(HP_FULL_MASK, A1[kk], sl);
                    const uint32_t u2 = __shfl_sync(HP_FULL_MASK, B0[kk], sl), u3 = __shfl_sync(HP_FULL_MASK, B1[kk], sl);
                    const uint32_t uw = kCount ? __shfl_sync(HP_FULL_MASK, wv[kk], sl) : 0u;
                    if ((carry >> 5) == (uint32_t)kk) { g0 = u0; g1 = u1; g2 = u2; g3 = u3; gw = uw; }
                }
                const bool has = carry != 0xffffu;
                nA0[k] = has ? g0 : 0u; nA1[k] = has ? g1 : 0u; nB0[k] = has ? g2 : 0u; nB1[k] = has ? g3 : 0u;
                nW[k] = has ? gw : 0u;
                colc[k] = cn;
                coln[k] = colnn[k];
            }
            o_p = o_p1; o_p1 = o_p2; o_p2 = o_p3;
        }
        cache_first = next_idx; cache_present = present;

        if (kCount) { const long long t1 = clock64(); tm_vec += t1 - tq0; tq0 = t1; }
        // ---- records: siblings get copies of the parent's words + their allele bit; the best child takes the parent's
        //      record in place ----
        const uint32_t c_mine = (lane - rr) & 31u;                         // this lane's candidate (pushes), if < 4
        const bool mine = c_mine < 4u && ((present >> c_mine) & 1u) && c_mine != best;
        uint32_t my_rec = 0;
        {
            const uint32_t nsib = nchild - 1;
            // sibling ordinal of candidate c = number of present non-best candidates below it
            const uint32_t sibmask = present & ~(1u << best);
            const uint32_t my_ord = __popc(sibmask & ((1u << (c_mine & 31u)) - 1u));
            if (fs_top < nsib && free_top != 0) {
                // refill the shared-memory stack from the slab's free list in one coalesced pass (bulk frees land there)
                const uint32_t n = min(min(free_top, kFreeStack - fs_top), 128u);
                for (uint32_t q = lane; q < n; q += 32) w.free_stack[fs_top + q] = s.freelist[free_top - n + q];
                fs_top += n; free_top -= n;
                __syncwarp();
            }
            const uint32_t from_fs = min(nsib, fs_top);                   // shared-memory stack first, then the slab list, then fresh
            const uint32_t from_gl = min(nsib - from_fs, free_top);
            if (mine) {
                if (my_ord < from_fs) my_rec = w.free_stack[fs_top - 1 - my_ord];
                else if (my_ord - from_fs < from_gl) my_rec = s.freelist[free_top - 1 - (my_ord - from_fs)];
                else my_rec = rec_next + (my_ord - from_fs - from_gl);
            }
            fs_top -= from_fs; free_top -= from_gl; rec_next += nsib - from_fs - from_gl;
            // haplotype words
            const uint32_t wl = L >> 6;
            const uint64_t bit = bad_col ? 0ull : (1ull << (L & 63));
            if (regs_hap) {
                // lane = (sibling ordinal, word within a group of 8): up to 4 siblings x 8 words per step
                const uint32_t sib = lane >> 3, wq = lane & 7u;
                uint32_t sm = sibmask;
                if (sib >= 1u) sm &= sm - 1u;
                if (sib >= 2u) sm &= sm - 1u;
                if (sib >= 3u) sm &= sm - 1u;
                const bool has_sib = sib < nsib;
                const uint32_t csl = has_sib ? (uint32_t)(__ffs(sm) - 1) : 0u;           // candidate slot of that sibling
                const uint32_t rc = __shfl_sync(HP_FULL_MASK, my_rec, (rr + csl) & 31u);
                uint64_t* crow = s.recs + (uint64_t)rc * 2 * HW;
                const bool s1bit = (csl & 1u) != 0, s2bit = ((0x9u >> csl) & 1u) != 0;
                for (uint32_t w0 = 0; w0 <= wl; w0 += 8) {
                    const uint32_t wi = w0 + wq;
                    const uint64_t x1 = __shfl_sync(HP_FULL_MASK, cur_w1, wi & 31u), x2 = __shfl_sync(HP_FULL_MASK, cur_w2, wi & 31u);
                    if (has_sib && wi <= wl) {
                        crow[wi] = (wi == wl && s1bit) ? (x1 | bit) : x1;
                        crow[HW + wi] = (wi == wl && s2bit) ? (x2 | bit) : x2;
                    }
                }
                if (lane == wl) {                                                     // the best child: in place, one word
                    uint64_t* brow = s.recs + (uint64_t)cur_rec * 2 * HW;
                    brow[wl] = (best & 1u) ? (cur_w1 | bit) : cur_w1;
                    brow[HW + wl] = ((0x9u >> best) & 1u) ? (cur_w2 | bit) : cur_w2;
                }
            } else {
                const uint64_t* prow = s.recs + (uint64_t)cur_rec * 2 * HW;
                const uint32_t nwords = (L + 63) >> 6;
#pragma unroll
                for (uint32_t cc = 0; cc < 4; cc++) {
                    if ((present >> cc) & 1u) {
                        const uint32_t rc = (cc == best) ? cur_rec : __shfl_sync(HP_FULL_MASK, my_rec, (rr + cc) & 31u);
                        uint64_t* crow = s.recs + (uint64_t)rc * 2 * HW;
                        for (uint32_t wi = lane; wi <= wl; wi += 32) {
                            if (cc != best || wi == wl) {
                                const uint64_t w1 = wi < nwords ? prow[wi] : 0ull, w2 = wi < nwords ? prow[HW + wi] : 0ull;
                                crow[wi] = (wi == wl && (cc & 1u)) ? (w1 | bit) : w1;
                                crow[HW + wi] = (wi == wl && ((0x9u >> cc) & 1u)) ? (w2 | bit) : w2;
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        if (kCount) { const long long t1 = clock64(); tm_rec += t1 - tq0; tq0 = t1; }
        // ---- push the siblings (one lane each) ----
        {
            const uint32_t mt = (c_mine == 0u) ? t0 : (c_mine == 1u) ? t1 : (c_mine == 2u) ? t2 : t3;
            const uint32_t mi = (c_mine == 0u) ? i0 : (c_mine == 1u) ? i1 : (c_mine == 2u) ? i2 : i3;
            const uint64_t mhi = ((uint64_t)mt << 32) | ((c_mine < 2u && !bad_col) ? nh_het : cur_nh);
            const uint32_t mlen = (L + 1) | ((cur_ident && c_mine >= 2u) ? 0x80000000u : 0u);
            const uint32_t fullmask = __ballot_sync(HP_FULL_MASK, mine && cnt >= scap);
            if (fullmask == 0) {
                if (mine) {
                    if (cnt < mqs) {                                       // shared-memory part of the stripe (the common case)
                        const uint32_t o = lane * mqs + cnt;
                        mq_hi[o] = mhi; mq_idx[o] = mi; mq_len[o] = mlen; mq_rec[o] = my_rec;
                    } else {
                        const size_t o = (size_t)lane * scap + cnt;
                        s.khi[o] = mhi; s.kidx[o] = mi; s.klen[o] = mlen; s.krec[o] = my_rec;
                    }
                    if (key_less(mhi, mi, c_hi, c_idx)) { c_hi = mhi; c_idx = mi; c_pos = cnt; }
                    cnt++;
                }
            } else {                                                      // rare: a target stripe is full
                for (uint32_t cc = 0; cc < 4; cc++) {
                    if (((present >> cc) & 1u) && cc != best) {
                        const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < scap);
                        if (room == 0) { w.status = HP_BLOCK_QUEUE_OVERFLOW; break; }
                        const uint32_t src_lane = (rr + cc) & 31u;
                        uint32_t target = src_lane;
                        if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                        const uint64_t xhi = __shfl_sync(HP_FULL_MASK, mhi, src_lane);
                        const uint32_t xi = __shfl_sync(HP_FULL_MASK, mi, src_lane), xl = __shfl_sync(HP_FULL_MASK, mlen, src_lane);
                        const uint32_t xr = __shfl_sync(HP_FULL_MASK, my_rec, src_lane);
                        if (lane == target) {
                            *khi_at(lane, cnt) = xhi; *kidx_at(lane, cnt) = xi; *klen_at(lane, cnt) = xl; *krec_at(lane, cnt) = xr;
                            if (key_less(xhi, xi, c_hi, c_idx)) { c_hi = xhi; c_idx = xi; c_pos = cnt; }
                            cnt++;
                        }
                    }
                }
            }
            rr += 4;
        }
        if (w.status != HP_BLOCK_OK) break;
        // queue minimum now includes the siblings
        {
            // among the siblings the smallest key is the smallest total, ties to the lowest slot (slots 0,1 carry one more
            // het, and node indices grow with the slot)
            const uint32_t u0 = (best == 0u) ? 0xffffffffu : t0, u1 = (best == 1u) ? 0xffffffffu : t1;
            const uint32_t u2 = (best == 2u) ? 0xffffffffu : t2, u3 = (best == 3u) ? 0xffffffffu : t3;
            const uint32_t m2 = min(min(u0, u1), min(u2, u3));
            if (m2 != 0xffffffffu) {
                const uint32_t c2 = (u0 == m2) ? 0u : (u1 == m2) ? 1u : (u2 == m2) ? 2u : 3u;
                const uint64_t h = ((uint64_t)m2 << 32) | ((c2 < 2u && !bad_col) ? nh_het : cur_nh);
                const uint32_t ix = (c2 == 0u) ? i0 : (c2 == 1u) ? i1 : (c2 == 2u) ? i2 : i3;
                if (key_less(h, ix, qmin.hi, qmin.idx)) { qmin.hi = h; qmin.idx = ix; }
            }
        }
        next_idx += nchild; qsize += nchild;
        if (lane == 0) atomicAdd(lc + L + 1, nchild);                     // tracker.add_hap(L+1) x nchild (:531, :558)
        if (L + 1 >= trk_thresh) trk_total += nchild;
        // ---- the best child is the new cur (record inherited in place) ----
        cur_total = tmin;
        cur_nh = (best < 2u && !bad_col) ? nh_het : cur_nh;
        cur_idx = (best == 0u) ? i0 : (best == 1u) ? i1 : (best == 2u) ? i2 : i3;
        cur_len = L + 1;
        cur_ident = cur_ident && best >= 2u;
        cur_x1 = best & 1u; cur_x2 = (0x9u >> best) & 1u;
        cur_src = SRC_CACHE;
        heur_p = heur;
        if (regs_hap && !bad_col && lane == (L >> 6)) {
            const uint64_t b = 1ull << (L & 63);
            if (cur_x1) cur_w1 |= b;
            if (cur_x2) cur_w2 |= b;
        }
        __syncwarp();

        if (kCount) { const long long t1 = clock64(); tm_push += t1 - tq0; tq0 = t1; }
        // ---- pruning bookkeeping (:564-585) ----
        while (trk_total > curr_thresh && min_progress < next_expected) {
            min_progress++;
            const uint32_t dl = min_progress - 1;                          // entries of exactly this length die now
            uint32_t dropped = 0;
            if (lane == 0) dropped = lc[dl];
            dropped = __shfl_sync(HP_FULL_MASK, dropped, 0);
            trk_total -= dropped; trk_thresh = min_progress;
            // sweep policy: any subset of the dead entries may move to the pool at any time (dead pops commute); entries
            // left in their stripe are discarded by the ordinary pop path when they surface
            // (swept when enough dead entries have piled up in the stripes to pay for the pass: a sweep plus the pool
            //  evaluation cost a few thousand cycles, discarding one dead entry at the top ~2000)
            if (kDeadPool && dropped != 0 && qsize - pool_n - trk_total >= kSweepMinDead) {
                eval_pool();                                               // one epoch for old and new pool entries
                if (lane == 0) { *w.free_ctr = free_top; *w.pool_ctr = pool_n; }
                __syncwarp();
                uint32_t j = 0;
                uint64_t m_hi = ~0ull; uint32_t m_idx = 0xffffffffu;
                for (uint32_t i = 0; i < cnt; i++) {
                    const uint32_t ln = *klen_at(lane, i);
                    if ((ln & 0x7fffffffu) <= dl) {
                        const uint64_t hi = *khi_at(lane, i); const uint32_t ix = *kidx_at(lane, i);
                        const uint32_t ps = atomicAdd(w.pool_ctr, 1u);
                        pool_hi[ps] = hi; pool_idx[ps] = ix;
                        s.freelist[atomicAdd(w.free_ctr, 1u)] = *krec_at(lane, i);
                        if (key_less(hi, ix, m_hi, m_idx)) { m_hi = hi; m_idx = ix; }
                    } else {
                        if (j != i) {
                            *khi_at(lane, j) = *khi_at(lane, i); *kidx_at(lane, j) = *kidx_at(lane, i);
                            *klen_at(lane, j) = ln; *krec_at(lane, j) = *krec_at(lane, i);
                        }
                        j++;
                    }
                }
                if (j != cnt) {                                            // this stripe lost entries: new cached minimum
                    cnt = j;
                    c_hi = ~0ull; c_idx = 0xffffffffu; c_pos = 0;
                    for (uint32_t i = 0; i < cnt; i++) {
                        const uint64_t hi = *khi_at(lane, i); const uint32_t ix = *kidx_at(lane, i);
                        if (key_less(hi, ix, c_hi, c_idx)) { c_hi = hi; c_idx = ix; c_pos = i; }
                    }
                }
                __syncwarp();
                free_top = *w.free_ctr; pool_n = *w.pool_ctr;
                __syncwarp();
                if (have_cur && cur_len <= dl) {                           // the node in registers dies as well
                    const uint64_t hi = ((uint64_t)cur_total << 32) | cur_nh;
                    if (lane == 0) { pool_hi[pool_n] = hi; pool_idx[pool_n] = cur_idx; s.freelist[free_top] = cur_rec; }
                    pool_n++; free_top++;
                    if (key_less(hi, cur_idx, m_hi, m_idx)) { m_hi = hi; m_idx = cur_idx; }
                    have_cur = false;
                    __syncwarp();
                }
                const MainKey nm = wmin96(m_hi, m_idx);
                if (key_less(nm.hi, nm.idx, pool_min.hi, pool_min.idx)) pool_min = nm;
                qmin = wmin96(c_hi, c_idx);
            }
            if (kDeadPool && qsize > max_queue && pool_n != 0) eval_pool();   // qsize was an upper bound: make it exact
            if (kDeadPool && qsize > max_queue) {
                // "full prune": every dead entry gets the cleared priority (cost 0)
                uint64_t m_hi = ~0ull; uint32_t m_idx = 0xffffffffu;
                for (uint32_t i = lane; i < pool_n; i += 32) {
                    const uint64_t hi = pool_hi[i] & 0xffffffffull; const uint32_t ix = pool_idx[i];
                    pool_hi[i] = hi;
                    if (key_less(hi, ix, m_hi, m_idx)) { m_hi = hi; m_idx = ix; }
                }
                __syncwarp();
                pool_min = wmin96(m_hi, m_idx);
            }
            if (qsize > max_queue) {                                       // ... including the ones still in their stripe
                if (have_cur && cur_len < min_progress) cur_total = 0;     // and the node held in registers
                c_hi = ~0ull; c_idx = 0xffffffffu; c_pos = 0;
                for (uint32_t i = 0; i < cnt; i++) {
                    uint64_t hi = *khi_at(lane, i);
                    const uint32_t ix = *kidx_at(lane, i);
                    if ((*klen_at(lane, i) & 0x7fffffffu) < min_progress) { hi &= 0xffffffffull; *khi_at(lane, i) = hi; }
                    if (key_less(hi, ix, c_hi, c_idx)) { c_hi = hi; c_idx = ix; c_pos = i; }
                }
                __syncwarp();
                qmin = wmin96(c_hi, c_idx);
            }
        }
        if (kCount) tm_rest += clock64() - tq0;
    }
    if (w.status == HP_BLOCK_OK && pool_n != 0) eval_pool();                   // dead entries popped before the final node
    if (kCount && a.dbg_cycles && lane == 0) {
        uint64_t* d = a.dbg_cycles + 16ull * blk;
        d[8] = tm_pop; d[9] = tm_exp; d[10] = tm_rest; d[11] = n_planes; d[12] = tm_planes;
        d[13] = n_real; d[14] = num_pruned; d[15] = qsize; d[7] = n_swept;
#ifdef HP_DBG_MAIN_SPLIT
        d[15] = tm_dead;     // cycles spent discarding dead top entries (part of the real-pop time)
#endif
        d[5] = tm_vec; d[6] = tm_rec; d[4] = tm_push;
#ifdef HP_DBG_RING_AGE
        d[8] = age_le[0]; d[9] = age_le[1]; d[10] = age_le[2]; d[12] = age_le[3];
#endif   // (overwritten below by the team's round statistics unless HP_DBG_MAIN_SPLIT)
    }
    if (w.status != HP_BLOCK_OK) return;

    // ---- final node (:588-628) ----
    const uint32_t top_total = cur_total;
    const uint64_t* frow = s.recs + (uint64_t)cur_rec * 2 * HW;
    uint32_t phased = 0, phased_snv = 0, skipped = 0;
    const uint8_t* snv = a.is_snv + m.var_base;
    __syncwarp();
    for (uint32_t i = lane; i < N; i += 32) {
        uint32_t b1 = (uint32_t)(frow[i >> 6] >> (i & 63)) & 1u;
        uint32_t b2 = (uint32_t)(frow[HW + (i >> 6)] >> (i & 63)) & 1u;
        if (__ldg(ign + i)) { b1 = 2; b2 = 2; skipped++; }
        else if (b1 != b2) { phased++; if (__ldg(snv + i)) phased_snv++; }
        a.out_h1[m.var_base + i] = (uint8_t)b1;
        a.out_h2[m.var_base + i] = (uint8_t)b2;
    }
    phased = wsum(phased); phased_snv = wsum(phased_snv); skipped = wsum(skipped);
    if (lane == 0) {
        uint64_t* st = a.out_stats + (uint64_t)blk * 7;
        st[0] = num_pruned;
        st[1] = Hg[0];
        st[2] = top_total;
        st[3] = phased; st[4] = phased_snv; st[5] = N - phased - skipped; st[6] = skipped;
    }
    if (top_total < Hg[0]) w.status = HP_BLOCK_ASSERT;                     // phase_stats.rs:163
}

// ---- team: the warps of one CTA work on one phase block ---------------------------------------------------------
// The heuristic pre-pass is a chain: sub-solve v needs H[v+1..v+40] and the clip size left by sub-solve v+1.  But
// H[v] == H[v+1] for most variants, so warp i of the team solves variant v_hi - i SPECULATIVELY, reading every not yet
// known H entry as H[v_hi+1] and assuming every earlier sub-solve of the round solves its full clip.  After a CTA
// barrier the results are verified in chain order and the longest valid prefix is committed (warp 0 never
// speculates, so every round commits at least one variant).  Results are exactly those of the serial chain.
#ifndef HP_MAX_TEAM
#define HP_MAX_TEAM 4
#endif
constexpr int kMaxTeam = HP_MAX_TEAM;
#ifndef HP_CTAS_PER_SM
#define HP_CTAS_PER_SM (16 / HP_MAX_TEAM)
#endif
// The "dense" build of the production kernels: 20 warps per SM in 96 registers.  The sub-solver's dive stays in registers (the
// packed score vectors made room), the main loop and the real-pop path spill, so a block's chain is ~25 % slower -- but with
// batches in flight (throughput regime: more blocks than warps) five warps per scheduler cover the dive's dependent chain
// better than four: +10 % blocks/s on the C3 stream (profiles/r2l_warps_per_sm.txt).  A launch alone on the device, and any
// launch with fewer blocks than warps, keeps the 128-register build (its length is its slowest block's chain).
#ifndef HP_CTAS_PER_SM_DENSE
#define HP_CTAS_PER_SM_DENSE (20 / HP_MAX_TEAM)
#endif
#ifndef HP_SPEC_FAIL_ROUNDS
#define HP_SPEC_FAIL_ROUNDS 6
#endif
#ifndef HP_SPEC_PAUSE_ROUNDS
#define HP_SPEC_PAUSE_ROUNDS 24
#endif
constexpr uint32_t kSpecFailRounds = HP_SPEC_FAIL_ROUNDS, kSpecPauseRounds = HP_SPEC_PAUSE_ROUNDS;

struct TeamShared {
    uint32_t blk;
    uint32_t est[kMaxTeam], solved[kMaxTeam];
    int32_t status[kMaxTeam];
    int32_t final_status;
    unsigned long long ctr[4];       // accepted work counters (evals, cells, sum_lp, pops)
    uint32_t hring[64];
    uint32_t free_stack[kFreeStack];
    uint32_t free_ctr, pool_ctr;
};

__device__ __forceinline__ uint64_t bad_window(const uint8_t* ign, uint32_t v, uint32_t N, uint32_t lane) {
    // bit i = ignored[v + i], i < 40
    const uint32_t lo = __ballot_sync(HP_FULL_MASK, v + lane < N && __ldg(ign + v + lane) != 0);
    const uint32_t hi = __ballot_sync(HP_FULL_MASK, lane < 8 && v + 32 + lane < N && __ldg(ign + v + 32 + lane) != 0);
    return ((uint64_t)hi << 32) | lo;
}

template <int K, bool kCount>
__device__ void solve_block(const AstarArgs& a, const BlkMeta& m, WarpCtx& w, const Slab& slab, uint32_t blk,
                            TeamShared& ts, uint32_t warp, uint32_t team) {
    const uint32_t N = m.n_var;
    uint32_t* Hg = a.heur + m.var_base + blk;
    const long long t_start = kCount ? clock64() : 0;
    const uint8_t* ign = a.ignored + m.var_base;
    if (threadIdx.x == 0) { ts.hring[N & 63] = 0; Hg[N] = 0; ts.final_status = HP_BLOCK_OK; }
    if (threadIdx.x < 4) ts.ctr[threadIdx.x] = 0;
    __syncthreads();

    // ---- calculate_astar_heuristic (:246-292), `team` variants per round ----
    int v_hi = (int)N - 1;
    uint32_t clip0 = 1;
    int status = HP_BLOCK_OK;
    uint32_t n_rounds = 0;
    long long t_wait = 0, t_subs = 0;
    // Speculation pays only where H[v] == H[v+1] is common.  In noisy stretches every guess fails and the guessing warps
    // only slow warp 0 down (shared issue slots, the round lasts as long as its slowest sub-solve): after kSpecFailRounds
    // rounds in a row that committed a single variant the team runs warp 0 alone for kSpecPauseRounds rounds, then tries
    // again.  Which warps compute never changes what is committed.
    uint32_t spec_streak = 0, spec_pause = 0;
    while (v_hi >= 0 && status == HP_BLOCK_OK) {
        const uint32_t eff_team = spec_pause ? 1u : team;
        const int v = v_hi - (int)warp;
        const uint32_t clip_guess = min(clip0 + warp, HP_MAX_SEGMENT);
        const uint64_t e0 = w.evals, c0 = w.cells, l0 = w.sum_lp, p0 = w.pops;
        if (v >= 0 && warp < eff_team) {
            w.h_floor = (uint32_t)v_hi + 1;
            w.status = HP_BLOCK_OK;
            const long long ts0 = kCount ? clock64() : 0;
            const uint2 r = sub_solve<K, kCount>(a, m, w, (uint32_t)v, clip_guess, bad_window(ign, (uint32_t)v, N, w.lane), blk);
            if (kCount) t_subs += clock64() - ts0;
            if (w.lane == 0) { ts.est[warp] = r.x; ts.solved[warp] = r.y; ts.status[warp] = w.status; }
        }
        const long long tw0 = kCount ? clock64() : 0;
        __syncthreads();
        if (kCount) t_wait += clock64() - tw0;
        n_rounds++;
        // ---- verify in chain order (every thread computes the same) ----
        const uint32_t h_base = ts.hring[(v_hi + 1) & 63];
        uint32_t h_next = h_base, clip_chk = clip0, accepted = 0;
        uint32_t hv_out[kMaxTeam];
        bool chain_ok = true;
#pragma unroll
        for (uint32_t i = 0; i < (uint32_t)kMaxTeam; i++) {
            const int vi = v_hi - (int)i;
            if (i < eff_team && vi >= 0 && chain_ok && status == HP_BLOCK_OK && clip_chk == min(clip0 + i, HP_MAX_SEGMENT)) {
                const uint32_t est = ts.est[i], solved = ts.solved[i];
                if (ts.status[i] != HP_BLOCK_OK) status = ts.status[i];
                else if (solved < min(clip_chk, 2u)) status = HP_BLOCK_ASSERT;                 // :268
                else {
                    uint32_t hv = h_next;
                    if (!__ldg(ign + vi)) {
                        if (est < h_next) status = HP_BLOCK_ASSERT;                              // :284
                        hv = est;
                    }
                    hv_out[i] = hv;
                    accepted = i + 1;
                    h_next = hv;
                    clip_chk = min(solved + 1, HP_MAX_SEGMENT);                                  // :288
                    chain_ok = (hv == h_base);        // later warps read H[vi] as h_base
                }
            } else chain_ok = false;
        }
        if (kCount && warp < accepted && w.lane == 0) {
            atomicAdd(&ts.ctr[0], (unsigned long long)(w.evals - e0)); atomicAdd(&ts.ctr[2], (unsigned long long)(w.sum_lp - l0));
            atomicAdd(&ts.ctr[3], (unsigned long long)(w.pops - p0));
        }
        if (kCount && warp < accepted) {
            const uint64_t dc = w.cells - c0;
            const uint64_t tot = __reduce_add_sync(HP_FULL_MASK, (uint32_t)(dc & 0xffffffffu)) +
                                 ((uint64_t)__reduce_add_sync(HP_FULL_MASK, (uint32_t)(dc >> 32)) << 32);
            if (w.lane == 0) atomicAdd(&ts.ctr[1], (unsigned long long)tot);
        }
        __syncthreads();                      // everyone has read hring / results of this round
        if (threadIdx.x == 0) {
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kMaxTeam; i++)
                if (i < accepted) { ts.hring[(v_hi - (int)i) & 63] = hv_out[i]; Hg[v_hi - (int)i] = hv_out[i]; }
        }
        if (spec_pause) spec_pause--;
        else if (team > 1) {
            spec_streak = (accepted <= 1) ? spec_streak + 1 : 0;
            if (spec_streak >= kSpecFailRounds) { spec_pause = kSpecPauseRounds; spec_streak = 0; }
        }
        v_hi -= (int)accepted;
        clip0 = clip_chk;
        if (accepted == 0 && status == HP_BLOCK_OK) status = HP_BLOCK_ASSERT;      // cannot happen: warp 0 is never speculative
        __syncthreads();
    }
    w.status = status;
    w.h_floor = 0;
    const long long t_mid = kCount ? clock64() : 0;
    // ---- main loop: one warp ----
    if (warp == 0) {
        w.evals = w.sum_lp = w.pops = w.cells = 0;
        if (w.status == HP_BLOCK_OK) {
            if constexpr (K == 0) main_solve<K, kCount>(a, m, w, slab, blk, Hg);
            else main_solve_fast<K, kCount>(a, m, w, slab, blk, Hg);
        }
        w.status = __shfl_sync(HP_FULL_MASK, w.status, 0);
        if (w.lane == 0) ts.final_status = w.status;
        if (kCount) {
            const uint64_t tot = __reduce_add_sync(HP_FULL_MASK, (uint32_t)(w.cells & 0xffffffffu)) +
                                 ((uint64_t)__reduce_add_sync(HP_FULL_MASK, (uint32_t)(w.cells >> 32)) << 32);
            if (w.lane == 0) {
                if (a.dbg_cycles) {
                    uint64_t* d = a.dbg_cycles + 16ull * blk;
                    d[0] = (uint64_t)(t_mid - t_start); d[1] = (uint64_t)(clock64() - t_mid); d[2] = ts.ctr[3]; d[3] = w.pops;
#ifndef HP_DBG_MAIN_SPLIT
                    d[4] = n_rounds; d[5] = team; d[6] = (uint64_t)t_wait;
#endif
#ifdef HP_DBG_SUB_SPLIT
                    // sub-solver phase split of warp 0 (its sub-solves are never speculative) instead of the main-loop split
                    d[8] = (uint64_t)w.ts_pop; d[9] = (uint64_t)w.ts_seat; d[10] = (uint64_t)w.ts_score; d[11] = (uint64_t)w.ts_rest;
                    d[12] = w.ns_real; d[13] = w.ns_planes; d[14] = w.ns_exp; d[15] = (uint64_t)w.ts_popa; d[7] = (uint64_t)w.ts_popb; d[5] = (uint64_t)t_subs; d[6] = (uint64_t)t_wait;
#endif
                }
                ts.ctr[0] += w.evals; ts.ctr[1] += tot; ts.ctr[2] += w.sum_lp; ts.ctr[3] += w.pops;
            }
        }
    }
    __syncthreads();
}

#ifndef HP_SUB_CAPL_S
#define HP_SUB_CAPL_S 8
#endif
constexpr int kSubCaplShared = HP_SUB_CAPL_S;    // sub-solver queue entries per stripe kept in shared memory (rest: global spill)

// One kernel per score-vector class K (1: <= 32 reads per column, 2: <= 64, 0: any) keeps the register footprint of the
// common class small.  Class c owns order[class_start[c] .. +class_count[c]) and ticket[c].
template <int K, bool kCount, bool kDense>
__global__ void __launch_bounds__(kMaxTeam * 32, kDense ? HP_CTAS_PER_SM_DENSE : HP_CTAS_PER_SM) astar_solve_kernel(AstarArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ TeamShared ts;
    constexpr int cls = (K == 1) ? 0 : (K == 2 ? 1 : 2);
    const uint32_t n_mine = a.class_info[cls], first = a.class_info[4 + cls];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t team = blockDim.x >> 5;
    WarpCtx w;
    w.lane = lane_id();
    w.capl = a.sub_capl;
    w.capl_s = min(a.sub_capl, (uint32_t)kSubCaplShared);
    w.sq = (SubEntry*)(smem_raw + (size_t)warp * w.capl_s * 32 * sizeof(SubEntry));
    w.hring = ts.hring;
    w.h_floor = 0;
    w.evals = w.sum_lp = w.pops = w.cells = 0;
    w.ts_pop = w.ts_seat = w.ts_score = w.ts_rest = w.ts_popa = w.ts_popb = 0; w.ns_real = w.ns_planes = w.ns_exp = 0;

    // one slab per CTA (team): the main queue (used by warp 0) followed by one sub-queue spill region per warp.  Slabs
    // come from a pool shared by every launch of the context (batches in flight on different lanes run concurrently): the
    // CTA claims a free one when it gets its first block and gives it back when it exits.  The pool holds at least as many
    // slabs as CTAs can be resident, and a waiting CTA holds nothing, so the probe loop always terminates.
    __shared__ uint32_t s_slab;
    bool have_slab = false;
    Slab slab;
    const uint64_t spill_bytes = sub_spill_bytes(w.capl, w.capl_s);
    // main-queue keys reuse the whole team's sub-queue shared memory (12 B per key)
    // ... minus a 4 KB tail for the tracker's length counts
    const size_t team_smem = (size_t)team * w.capl_s * 32 * sizeof(SubEntry);
    const size_t lencnt_bytes = team_smem >= 8192 ? 4096 : 0;
    w.lencnt_s = (uint32_t*)(smem_raw + team_smem - lencnt_bytes);
    w.lencnt_cap_s = (uint32_t)(lencnt_bytes / 4);
    w.mq_cap_s = (uint32_t)((team_smem - lencnt_bytes) / (32 * 20));
    w.mq_hi = (uint64_t*)smem_raw;
    w.mq_idx = (uint32_t*)(smem_raw + (size_t)32 * w.mq_cap_s * 8);
    w.mq_len = w.mq_idx + (size_t)32 * w.mq_cap_s;
    w.mq_rec = w.mq_len + (size_t)32 * w.mq_cap_s;
    w.free_stack = ts.free_stack;
    w.free_ctr = &ts.free_ctr;
    w.pool_ctr = &ts.pool_ctr;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) ts.blk = atomicAdd(a.ticket + cls, 1u);
        __syncthreads();
        const uint32_t t = ts.blk;
        if (t >= n_mine) break;
        if (!have_slab) {
            if (threadIdx.x == 0) {
                uint32_t i = (blockIdx.x + a.slab_seed) % a.n_slabs;
                while (atomicCAS(a.slab_busy + i, 0u, 1u) != 0u) { i = (i + 1u == a.n_slabs) ? 0u : i + 1u; }
                __threadfence();
                s_slab = i;
            }
            __syncthreads();
            uint8_t* my_slab = a.slabs + (uint64_t)s_slab * a.slab_bytes;
            slab = carve_slab(my_slab, a.qcap, a.hap_words);
            w.sq_spill = (SubEntry*)(my_slab + a.slab_bytes - (uint64_t)(kMaxTeam - warp) * spill_bytes);
            have_slab = true;
        }
        const uint32_t blk = a.order[first + t];
        const BlkMeta m = a.meta[blk];
        w.evals = w.sum_lp = w.pops = w.cells = 0;
        w.ts_pop = w.ts_seat = w.ts_score = w.ts_rest = w.ts_popa = w.ts_popb = 0; w.ns_real = w.ns_planes = w.ns_exp = 0;
        w.status = m.status;
        if (threadIdx.x == 0) ts.final_status = m.status;

        if (m.status == HP_BLOCK_OK) solve_block<K, kCount>(a, m, w, slab, blk, ts, warp, team);
        __syncthreads();
        const int status = ts.final_status;

        // ---- per-block outputs (warp 0) ----
        if (warp == 0) {
            if (status != HP_BLOCK_OK) {
                if (w.lane < 7) a.out_stats[(uint64_t)blk * 7 + w.lane] = 0;
            }
            if (a.out_heur && status == HP_BLOCK_OK) {
                const uint32_t* Hg = a.heur + m.var_base + blk;
                for (uint32_t i = w.lane; i <= m.n_var; i += 32) a.out_heur[m.var_base + blk + i] = Hg[i];
            }
            if (kCount && a.out_counters && w.lane == 0) {
                uint64_t* c = a.out_counters + (uint64_t)blk * 4;
                if (status == HP_BLOCK_OK) { c[0] = ts.ctr[0]; c[1] = ts.ctr[1]; c[2] = ts.ctr[2]; c[3] = ts.ctr[3]; }
                else { c[0] = c[1] = c[2] = c[3] = 0; }
            }
            if (w.lane == 0) a.out_status[blk] = status;
        }
    }
    if (have_slab) {
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); atomicExch(a.slab_busy + s_slab, 0u); }
    }
}

}  // namespace hp

// Test aid: score_planes (= ReadSegment::score_partial_haplotype for h1 and h2, read_segments.rs:177-206) of every read of
// block 0 against two haplotypes given as bit masks (bit j = allele at haplotype position j), problem offset and length.
namespace hp {
__global__ void score_planes_debug_kernel(AstarArgs a, uint64_t h1, uint64_t h2, int offset, int L, uint32_t* s1, uint32_t* s2) {
    const BlkMeta m = a.meta[0];
    auto hap = [&](int which, int i0) { return shift_signed(which ? h2 : h1, i0); };
    for (uint32_t r = threadIdx.x; r < m.n_reads; r += blockDim.x) {
        const ReadMeta rm = a.rmeta[m.read_base + r];
        uint32_t x1, x2;
        score_planes(a, m, rm, offset - (int)rm.start, L, hap, x1, x2);
        s1[r] = x1; s2[r] = x2;
    }
}
cudaError_t launch_score_planes_debug(const AstarArgs& a, uint64_t h1, uint64_t h2, int offset, int L, uint32_t* s1, uint32_t* s2,
                                      cudaStream_t stream) {
    score_planes_debug_kernel<<<1, 128, 0, stream>>>(a, h1, h2, offset, L, s1, s2);
    return cudaGetLastError();
}
}  // namespace hp

// ---- host-side launchers (called from hp_api.cu) ---------------------------------------------------------------
namespace hp {

size_t astar_smem_bytes(uint32_t sub_capl, int team) {
    const uint32_t capl_s = std::min<uint32_t>(sub_capl, kSubCaplShared);
    return (size_t)team * capl_s * 32 * sizeof(SubEntry);
}
int astar_max_team() { return kMaxTeam; }
int astar_warps_per_sm(bool dense) { return (dense ? HP_CTAS_PER_SM_DENSE : HP_CTAS_PER_SM) * kMaxTeam; }
uint64_t astar_slab_bytes(uint32_t qcap, uint32_t hap_words, uint32_t sub_capl) {
    const uint32_t capl_s = std::min<uint32_t>(sub_capl, kSubCaplShared);
    const uint64_t spill = (uint64_t)kMaxTeam * sub_spill_bytes(sub_capl, capl_s);
    return ((slab_bytes_for(qcap, hap_words) + spill + 255) & ~255ull);
}

// smem_cap (variants per block whose coverage counters fit shared memory) is chosen here from the largest block of the batch.
cudaError_t launch_astar_prep(PrepArgs pa, uint32_t max_block_vars, cudaStream_t stream) {
    if (pa.n_blocks == 0) return cudaSuccess;
    pa.smem_cap = std::min<uint32_t>(std::max<uint32_t>(max_block_vars, 64u), 8190u) | 1u;     // odd: (cap + 1) u32 + (cap + 1) u16 stay 4-aligned
    const size_t smem = (size_t)kPrepStages * 2 * kStageBytes + 4ull * (pa.smem_cap + 1) + 2ull * (pa.smem_cap + 1);
    cudaError_t e = cudaFuncSetAttribute(astar_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    astar_prep_kernel<<<pa.n_blocks, kPrepThreads, smem, stream>>>(pa);
    return cudaGetLastError();
}

template <int K, bool kCount, bool kDense = false>
static cudaError_t launch_one(const AstarArgs& a, int n_ctas, int team, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(astar_solve_kernel<K, kCount, kDense>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    astar_solve_kernel<K, kCount, kDense><<<n_ctas, team * 32, smem, stream>>>(a);
    return cudaGetLastError();
}

// Three launches (one per score-vector class, each on its own stream so they run side by side); a class without blocks
// exits at once.
cudaError_t launch_astar_solve(const AstarArgs& a, int n_ctas, int team, cudaStream_t* streams, bool dense) {
    const size_t smem = astar_smem_bytes(a.sub_capl, team);
    cudaError_t e;
    if (dense && !a.out_counters) {        // the two register-vector classes have a dense build; the generic class is rare
        if ((e = launch_one<1, false, true>(a, n_ctas, team, smem, streams[0])) != cudaSuccess) return e;
        if ((e = launch_one<2, false, true>(a, n_ctas, team, smem, streams[1])) != cudaSuccess) return e;
        return launch_one<0, false>(a, n_ctas, team, smem, streams[2]);
    }
    if (a.out_counters) {
        if ((e = launch_one<1, true>(a, n_ctas, team, smem, streams[0])) != cudaSuccess) return e;
        if ((e = launch_one<2, true>(a, n_ctas, team, smem, streams[1])) != cudaSuccess) return e;
        return launch_one<0, true>(a, n_ctas, team, smem, streams[2]);
    }
    if ((e = launch_one<1, false>(a, n_ctas, team, smem, streams[0])) != cudaSuccess) return e;
    if ((e = launch_one<2, false>(a, n_ctas, team, smem, streams[1])) != cudaSuccess) return e;
    return launch_one<0, false>(a, n_ctas, team, smem, streams[2]);
}

// Upper bound of solver CTAs resident on one SM over every kernel variant at the smallest team (sizes the slab pool).
int astar_max_ctas_per_sm(uint32_t sub_capl) {
    int best = 0;
    const size_t smem = astar_smem_bytes(sub_capl, 1);
    auto probe = [&](auto kern) {
        int nb = 0;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32, smem) == cudaSuccess) best = std::max(best, nb);
        else cudaGetLastError();
    };
    probe(astar_solve_kernel<1, false, false>); probe(astar_solve_kernel<2, false, false>); probe(astar_solve_kernel<0, false, false>);
    probe(astar_solve_kernel<1, true, false>); probe(astar_solve_kernel<2, true, false>); probe(astar_solve_kernel<0, true, false>);
    probe(astar_solve_kernel<1, false, true>); probe(astar_solve_kernel<2, false, true>);
    return best > 0 ? best : 32;
}

}  // namespace hp
