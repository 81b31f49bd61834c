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

constexpr int kPrepThreads = 128;

__global__ void __launch_bounds__(kPrepThreads) astar_prep_kernel(PrepArgs a) {
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
    __shared__ uint32_t s_g, s_planes;

    if (tid < 8) s_presence[tid] = 0;
    if (tid == 0) { s_qsum = 0; s_status = HP_BLOCK_OK; s_maxspan = 0; }
    __syncthreads();

    uint32_t* cnt = a.act_off + v0 + b;   // N + 1 entries
    uint32_t* cur = a.act_cur + v0 + b;

    // ---- pass 1: validate, count coverage, collect the set of quality values ----
    {
        uint32_t pres[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned long long qsum = 0;
        int status = HP_BLOCK_OK;
        uint32_t maxspan = 0;
        if (N == 0) status = HP_BLOCK_ASSERT;
        for (uint64_t r = r0 + tid; r < r1; r += kPrepThreads) {
            const uint32_t s = a.read_start[r], e = a.read_end[r];
            const uint64_t c = a.cell_off[r];
            if (e < s || e > N || a.cell_off[r + 1] - c != (uint64_t)(e - s)) { status = HP_BLOCK_ASSERT; continue; }
            maxspan = max(maxspan, e - s);
            ReadMeta rm;
            rm.start = s; rm.end = e; rm.word_idx = (uint32_t)(c / 64 + r); rm.cell_rel = (uint32_t)(c - c0);
            a.rmeta[r] = rm;
            for (uint32_t i = 0; i < e - s; i++) {
                const uint8_t al = a.alleles[c + i];
                const uint8_t q = a.quals[c + i];
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
        uint32_t sum = 0;
        for (uint32_t i = lo; i < hi; i++) sum += __ldcg(&cnt[i]);
        s_partial[tid] = sum;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int t = 0; t < kPrepThreads; t++) { uint32_t x = s_partial[t]; s_partial[t] = run; run += x; }
    }
    __syncthreads();
    {
        uint32_t run = s_partial[tid];
        for (uint32_t i = lo; i < hi; i++) { uint32_t x = __ldcg(&cnt[i]); __stcg(&cnt[i], run); run += x; }
    }
    __syncthreads();

    // ---- pass 2: bit planes + active lists ----
    if (s_status == HP_BLOCK_OK) {
        for (uint64_t r = r0 + tid; r < r1; r += kPrepThreads) {
            const uint32_t s = a.read_start[r], e = a.read_end[r];
            const uint64_t c = a.cell_off[r];
            uint64_t* rec = a.planes + (uint64_t)(c / 64 + r) * HP_PLANE_STRIDE;
            uint64_t w[HP_PLANE_STRIDE];
#pragma unroll
            for (int k = 0; k < (int)HP_PLANE_STRIDE; k++) w[k] = 0;
            for (uint32_t i = 0; i < e - s; i++) {
                const uint8_t al = a.alleles[c + i];
                const uint32_t p = s + i;
                // quality of an ignored column can never be charged (the haplotype is Ambiguous there): drop it
                const uint32_t q = a.ignored[v0 + p] ? 0u : (uint32_t)s_div[a.quals[c + i]];
                const uint64_t bit = 1ull << (i & 63);
                if (al & 1) w[0] |= bit;          // allele bit (meaningful for 0/1; 3 sets it too but nb masks it)
                if (al >= 2) w[1] |= bit;         // non-binary: mismatches both 0 and 1
#pragma unroll
                for (int k = 0; k < 8; k++) if (q >> k & 1u) w[2 + k] |= bit;
                const uint32_t slot = atomicAdd(&cur[p], 1u);
                a.act_idx[c0 + __ldcg(&cnt[p]) + slot] = (uint32_t)(r - r0);
                if ((i & 63) == 63 || i + 1 == e - s) {
#pragma unroll
                    for (int k = 0; k < (int)HP_PLANE_STRIDE; k++) { rec[k] = w[k]; w[k] = 0; }
                    rec += HP_PLANE_STRIDE;
                }
            }
        }
    }
    if (tid == 0) {
        BlkMeta m;
        m.var_base = v0; m.read_base = r0; m.cell_base = c0;
        m.n_var = N; m.n_reads = R; m.n_cells = (uint32_t)(c1 - c0);
        m.qgcd = g; m.n_planes = s_planes; m.status = s_status; m.max_span = s_maxspan; m.pad = 0;
        a.meta[b] = m;
    }
}

// =============================================================================================================
// solver kernel
// =============================================================================================================

// ---- warp reductions -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t wsum(uint32_t x) { return __reduce_add_sync(HP_FULL_MASK, x); }
__device__ __forceinline__ uint32_t wmin(uint32_t x) { return __reduce_min_sync(HP_FULL_MASK, x); }

// Per-lane accumulators of one expansion: deltas of the four children (0|1),(1|0),(0/0),(1/1) and of the
// single (2,2) child of an ignored variant; "t" = all active reads, "f" = reads that end at this column.
struct ChildAcc {
    uint32_t t01, t10, t00, t11, f01, f10, f00, f11, tb, fb;
};

// Scores one active read against both parent haplotypes and accumulates the child deltas.
//   o      = read coordinate of haplotype position 0   (problem_offset - read.start, may be negative)
//   L      = parent haplotype length (positions [0, L) are set)
//   HapFn  = functor(which, bitpos) -> 64 haplotype bits starting at haplotype position bitpos (may be < 0)
template <class HapFn>
__device__ __forceinline__ void score_read(const AstarArgs& a, const BlkMeta& m, const ReadMeta rm, int o, int L,
                                           uint32_t p, bool bad_col, HapFn hap, ChildAcc& acc, uint64_t& cells) {
    // parent scores over read coordinates [max(0,o), o+L)
    uint32_t s1 = 0, s2 = 0;
    const int c_lo = max(0, o), c_hi = o + L;           // c_hi > c_lo whenever L > 0 and the read is active at p
    if (c_hi > c_lo) {
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
    const bool ends = (rm.end <= p + 1);
    if (bad_col) {
        const uint32_t c = min(s1, s2);
        acc.tb += c; if (ends) acc.fb += c;
    } else {
        const uint32_t ci = rm.cell_rel + (p - rm.start);
        const uint32_t al = __ldg(a.alleles + m.cell_base + ci);
        const uint32_t q = __ldg(a.quals + m.cell_base + ci);
        const uint32_t q0 = (al != 0u) ? q : 0u;         // cost of haplotype allele 0 at this column
        const uint32_t q1 = (al != 1u) ? q : 0u;         // cost of haplotype allele 1
        const uint32_t c01 = min(s1 + q0, s2 + q1);
        const uint32_t c10 = min(s1 + q1, s2 + q0);
        const uint32_t c00 = min(s1, s2) + q0;
        const uint32_t c11 = min(s1, s2) + q1;
        acc.t01 += c01; acc.t10 += c10; acc.t00 += c00; acc.t11 += c11;
        if (ends) { acc.f01 += c01; acc.f10 += c10; acc.f00 += c00; acc.f11 += c11; }
    }
    cells += (uint64_t)(p + 1 - (uint32_t)max((int)rm.start, (int)rm.start + o));   // w_r = p+1 - max(start, offset)
}

// ---- sub-solver key: [total:32][63-hets:6][node_index:20][len:6] -------------------------------------------
__device__ __forceinline__ uint64_t sub_key(uint32_t total, uint32_t hets, uint32_t idx, uint32_t len) {
    return ((uint64_t)total << 32) | ((uint64_t)(63u - hets) << 26) | ((uint64_t)idx << 6) | len;
}

struct WarpCtx {
    // shared-memory views of this warp's sub-solver queue (stripe-major: lane l owns [l*capl, (l+1)*capl))
    uint64_t* sq_key;
    uint64_t* sq_h1;
    uint64_t* sq_h2;
    uint32_t* sq_frozen;
    uint32_t* hring;        // H[] ring buffer, 64 entries
    uint32_t capl;
    uint32_t lane;
    // counters
    uint64_t evals, sum_lp, pops, cells;
    int status;
};

// astar_subsolver (astar_phaser.rs:311-405).  Returns est in .x, solved depth in .y (both warp-uniform).
__device__ uint2 sub_solve(const AstarArgs& a, const BlkMeta& m, WarpCtx& w, uint32_t v, uint32_t clip, uint64_t badwin,
                           uint32_t blk) {
    const uint32_t lane = w.lane;
    const uint32_t* aoff = a.act_off + m.var_base + blk;
    const uint32_t* aidx = a.act_idx + m.cell_base;
    const ReadMeta* rmeta = a.rmeta + m.read_base;
    const uint32_t base = lane * w.capl;

    // queue state: cached stripe minimum + stripe count, in registers
    uint64_t ckey = ~0ull;
    uint32_t cpos = 0, cnt = 0;
    if (lane == 0) {                                                     // root: AstarNode::new(H[v+1]), :325
        w.sq_key[base] = sub_key(w.hring[(v + 1) & 63], 0, 0, 0);
        w.sq_h1[base] = 0; w.sq_h2[base] = 0; w.sq_frozen[base] = 0;
        ckey = w.sq_key[base]; cnt = 1;
    }
    __syncwarp();
    uint32_t next_idx = 1, next_expected = 0, max_cost = 0, visits = 0, rr = 1;
    const uint32_t max_visits = a.min_queue_size / 10 + a.queue_increment * clip;    // :266, :333

    for (;;) {
        // ---- peek: warp-wide minimum of the cached stripe minima ----
        const uint32_t khi = (uint32_t)(ckey >> 32), klo = (uint32_t)ckey;
        const uint32_t mhi = wmin(khi);
        const uint32_t mlo = wmin(khi == mhi ? klo : 0xffffffffu);
        const uint32_t L = mlo & 63u;
        if (L >= clip) {                                                 // :395-399 (peek, not pop)
            max_cost = max(max_cost, mhi);
            next_expected++;
            break;
        }
        if (visits >= max_visits) break;
        const int owner = __ffs(__ballot_sync(HP_FULL_MASK, khi == mhi && klo == mlo)) - 1;
        const uint32_t pos = __shfl_sync(HP_FULL_MASK, cpos, owner);
        const uint32_t slot = owner * w.capl + pos;
        const uint64_t ph1 = w.sq_h1[slot], ph2 = w.sq_h2[slot];
        const uint32_t pfrozen = w.sq_frozen[slot];
        const uint32_t phets = 63u - ((mlo >> 26) & 63u);
        __syncwarp();
        if ((int)lane == owner) {                                        // remove + rescan own stripe
            cnt--;
            if (pos != cnt) {
                w.sq_key[slot] = w.sq_key[base + cnt]; w.sq_h1[slot] = w.sq_h1[base + cnt];
                w.sq_h2[slot] = w.sq_h2[base + cnt]; w.sq_frozen[slot] = w.sq_frozen[base + cnt];
            }
            ckey = ~0ull; cpos = 0;
            for (uint32_t i = 0; i < cnt; i++) {
                const uint64_t k = w.sq_key[base + i];
                if (k < ckey) { ckey = k; cpos = i; }
            }
        }
        visits++;
        w.pops++;
        if (L == next_expected) { max_cost = max(max_cost, mhi); next_expected++; }   // :342-346

        // ---- expand: score the active reads of column p against both parent haplotypes ----
        const uint32_t p = v + L;
        const bool bad_col = (badwin >> L) & 1ull;
        const uint32_t heur = w.hring[(p + 1) & 63];
        const uint32_t a0 = __ldg(aoff + p), a1 = __ldg(aoff + p + 1);
        ChildAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        uint64_t cells = 0;
        auto hap = [&](int which, int i0) { return shift_signed(which ? ph2 : ph1, i0); };
        for (uint32_t j = a0 + lane; j < a1; j += 32) {
            const ReadMeta rm = rmeta[__ldg(aidx + j)];
            score_read(a, m, rm, (int)v - (int)rm.start, (int)L, p, bad_col, hap, acc, cells);
        }
        const bool ident = (ph1 == ph2);
        uint32_t nchild;
        uint32_t ctot[4], cfro[4], chet[4];
        uint64_t ch1[4], ch2[4];
        const uint64_t bit = 1ull << L;
        if (bad_col) {
            nchild = 1;
            ctot[0] = pfrozen + wsum(acc.tb) + heur; cfro[0] = pfrozen + wsum(acc.fb);
            chet[0] = phets; ch1[0] = ph1; ch2[0] = ph2;
            if (ctot[0] != mhi) w.status = HP_BLOCK_ASSERT;               // :360
        } else {
            const uint32_t t01 = wsum(acc.t01), t00 = wsum(acc.t00), t11 = wsum(acc.t11);
            const uint32_t f01 = wsum(acc.f01), f00 = wsum(acc.f00), f11 = wsum(acc.f11);
            nchild = 0;
            ctot[nchild] = pfrozen + t01 + heur; cfro[nchild] = pfrozen + f01; chet[nchild] = phets + 1;
            ch1[nchild] = ph1; ch2[nchild] = ph2 | bit; nchild++;
            if (!ident) {                                                // :376 symmetry break
                const uint32_t t10 = wsum(acc.t10), f10 = wsum(acc.f10);
                ctot[nchild] = pfrozen + t10 + heur; cfro[nchild] = pfrozen + f10; chet[nchild] = phets + 1;
                ch1[nchild] = ph1 | bit; ch2[nchild] = ph2; nchild++;
            }
            ctot[nchild] = pfrozen + t00 + heur; cfro[nchild] = pfrozen + f00; chet[nchild] = phets;
            ch1[nchild] = ph1; ch2[nchild] = ph2; nchild++;
            ctot[nchild] = pfrozen + t11 + heur; cfro[nchild] = pfrozen + f11; chet[nchild] = phets;
            ch1[nchild] = ph1 | bit; ch2[nchild] = ph2 | bit; nchild++;
        }
        w.evals += nchild; w.sum_lp += (uint64_t)nchild * L; w.cells += cells * nchild;

        // ---- push children, round-robin over the stripes ----
        const uint32_t fullmask = __ballot_sync(HP_FULL_MASK, cnt >= w.capl);
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
            if (c < nchild) {
                uint32_t target = (rr + c) & 31u;
                if (fullmask) {                                          // rare: pick any stripe with room
                    const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < w.capl);
                    if (room == 0) { w.status = HP_BLOCK_ASSERT; break; }
                    if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                }
                if (lane == target) {
                    const uint64_t key = sub_key(ctot[c], chet[c], next_idx + c, L + 1);
                    w.sq_key[base + cnt] = key; w.sq_h1[base + cnt] = ch1[c]; w.sq_h2[base + cnt] = ch2[c];
                    w.sq_frozen[base + cnt] = cfro[c];
                    if (key < ckey) { ckey = key; cpos = cnt; }
                    cnt++;
                }
            }
        }
        rr += nchild; next_idx += nchild;
        __syncwarp();
        if (w.status != HP_BLOCK_OK) break;
    }
    return make_uint2(max_cost, next_expected - 1);
}

// ---- main-queue slab (global memory, private to one warp) -----------------------------------------------------
struct Slab {
    uint64_t* khi;       // [qcap] total << 32 | (0xffffffff - hets)
    uint32_t* kidx;      // [qcap] node index
    uint32_t* klen;      // [qcap]
    uint32_t* kfrozen;   // [qcap]
    uint32_t* krec;      // [qcap] record slot
    uint32_t* freelist;  // [qcap]
    uint32_t* lencnt;    // [hap_words*64 + 2] PQueueHapTracker::length_counts
    uint64_t* recs;      // [qcap][2*hap_words]
};

__host__ __device__ inline uint64_t slab_bytes_for(uint32_t qcap, uint32_t hap_words) {
    uint64_t b = (uint64_t)qcap * (8 + 4 * 5) + (uint64_t)(hap_words * 64 + 2) * 4;
    b = (b + 15) & ~15ull;
    b += (uint64_t)qcap * 2 * hap_words * 8;
    return (b + 255) & ~255ull;
}

__device__ __forceinline__ Slab carve_slab(uint8_t* p, uint32_t qcap, uint32_t hap_words) {
    Slab s;
    s.khi = (uint64_t*)p; p += (uint64_t)qcap * 8;
    s.kidx = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.klen = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.kfrozen = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.krec = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.freelist = (uint32_t*)p; p += (uint64_t)qcap * 4;
    s.lencnt = (uint32_t*)p; p += (uint64_t)(hap_words * 64 + 2) * 4;
    p = (uint8_t*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
    s.recs = (uint64_t*)p;
    return s;
}

// Main-queue cached minimum of one stripe (registers of the owning lane).
struct MainMin {
    uint64_t hi;     // ~0 when the stripe is empty
    uint32_t idx;
    uint32_t pos;
};

__device__ __forceinline__ bool key_less(uint64_t hi_a, uint32_t idx_a, uint64_t hi_b, uint32_t idx_b) {
    return hi_a < hi_b || (hi_a == hi_b && idx_a < idx_b);
}

// warp-cooperative scan of stripe `owner` (cnt_o entries) -> its minimum; result valid on every lane
__device__ __forceinline__ MainMin stripe_min(const Slab& s, uint32_t scap, int owner, uint32_t cnt_o, uint32_t lane) {
    MainMin best = {~0ull, 0xffffffffu, 0};
    const uint32_t b0 = owner * scap;
    for (uint32_t i = lane; i < cnt_o; i += 32) {
        const uint64_t hi = __ldcg(s.khi + b0 + i);
        const uint32_t idx = __ldcg(s.kidx + b0 + i);
        if (key_less(hi, idx, best.hi, best.idx)) { best.hi = hi; best.idx = idx; best.pos = i; }
    }
    const uint32_t t = (uint32_t)(best.hi >> 32), h = (uint32_t)best.hi;
    const uint32_t mt = wmin(t);
    const uint32_t mh = wmin(t == mt ? h : 0xffffffffu);
    const uint32_t mi = wmin((t == mt && h == mh) ? best.idx : 0xffffffffu);
    const uint32_t win = __ballot_sync(HP_FULL_MASK, t == mt && h == mh && best.idx == mi);
    const int wl = __ffs(win) - 1;
    MainMin r;
    r.hi = ((uint64_t)mt << 32) | mh;
    r.idx = mi;
    r.pos = __shfl_sync(HP_FULL_MASK, best.pos, wl);
    return r;
}

// astar_solver main loop (astar_phaser.rs:480-633) for one block, after the heuristic pre-pass.
__device__ void main_solve(const AstarArgs& a, const BlkMeta& m, WarpCtx& w, const Slab& s, uint32_t blk,
                           const uint32_t* Hg) {
    const uint32_t lane = w.lane;
    const uint32_t N = m.n_var;
    const uint32_t HW = a.hap_words;
    const uint32_t scap = a.qcap / 32;
    const uint32_t* aoff = a.act_off + m.var_base + blk;
    const uint32_t* aidx = a.act_idx + m.cell_base;
    const ReadMeta* rmeta = a.rmeta + m.read_base;
    const uint8_t* ign = a.ignored + m.var_base;

    // tracker (PQueueHapTracker, :171-231)
    for (uint32_t i = lane; i <= N; i += 32) __stcg(s.lencnt + i, 0u);
    uint32_t trk_total = 0, trk_thresh = 0;
    uint32_t curr_thresh = a.min_queue_size;
    const uint32_t max_queue = 10u * a.min_queue_size;                   // :457
    uint32_t min_progress = 0, next_expected = 0;
    uint64_t num_pruned = 0;
    uint32_t next_idx = 1, rr = 1, qsize = 0;
    uint32_t free_top = 0, rec_next = 0;

    MainMin cm = {~0ull, 0xffffffffu, 0};
    uint32_t cnt = 0;
    // root node (:485-488): record 0, empty haplotypes
    if (lane == 0) {
        const uint32_t h0 = __ldcg(Hg + 0);
        __stcg(s.khi + 0, ((uint64_t)h0 << 32) | 0xffffffffull);
        __stcg(s.kidx + 0, 0u); __stcg(s.klen + 0, 0u); __stcg(s.kfrozen + 0, 0u); __stcg(s.krec + 0, 0u);
        __stcg(s.lencnt + 0, 1u);
        cm.hi = ((uint64_t)h0 << 32) | 0xffffffffull; cm.idx = 0; cm.pos = 0; cnt = 1;
    }
    rec_next = 1; qsize = 1; trk_total = 1;
    __syncwarp();

    uint32_t top_slot = 0, top_total = 0;
    for (;;) {
        // ---- peek ----
        const uint32_t t = (uint32_t)(cm.hi >> 32), h = (uint32_t)cm.hi;
        const uint32_t mt = wmin(t);
        const uint32_t mh = wmin(t == mt ? h : 0xffffffffu);
        const uint32_t mi = wmin((t == mt && h == mh) ? cm.idx : 0xffffffffu);
        const int owner = __ffs(__ballot_sync(HP_FULL_MASK, t == mt && h == mh && cm.idx == mi)) - 1;
        const uint32_t pos = __shfl_sync(HP_FULL_MASK, cm.pos, owner);
        const uint32_t slot = owner * scap + pos;
        const uint32_t L = __ldcg(s.klen + slot);
        top_slot = slot; top_total = mt;
        if (L >= N) break;                                                // :492
        const uint32_t pfrozen = __ldcg(s.kfrozen + slot);
        const uint32_t prec = __ldcg(s.krec + slot);
        const uint32_t phets = 0xffffffffu - mh;
        __syncwarp();
        // ---- pop: owner moves its last entry into the hole, then the warp rescans that stripe ----
        const uint32_t cnt_o = __shfl_sync(HP_FULL_MASK, cnt, owner) - 1;
        if ((int)lane == owner) {
            cnt--;
            if (pos != cnt) {
                const uint32_t last = owner * scap + cnt;
                __stcg(s.khi + slot, __ldcg(s.khi + last)); __stcg(s.kidx + slot, __ldcg(s.kidx + last));
                __stcg(s.klen + slot, __ldcg(s.klen + last)); __stcg(s.kfrozen + slot, __ldcg(s.kfrozen + last));
                __stcg(s.krec + slot, __ldcg(s.krec + last));
            }
        }
        __syncwarp();
        {
            const MainMin nm = stripe_min(s, scap, owner, cnt_o, lane);
            if ((int)lane == owner) cm = nm;
        }
        qsize--;
        // hap_tracker.remove_hap (:495)
        if (lane == 0) __stcg(s.lencnt + L, __ldcg(s.lencnt + L) - 1u);
        if (L >= trk_thresh) trk_total--;
        w.pops++;
        if (L == next_expected) {                                         // :497-504
            next_expected++;
            if (num_pruned == 0) curr_thresh += a.queue_increment;
        }
        if (L < min_progress) {                                           // :507-515
            if (num_pruned == 0) curr_thresh = a.min_queue_size;
            num_pruned++;
            if (lane == 0) __stcg(s.freelist + free_top, prec);
            free_top++;
            __syncwarp();
            continue;
        }

        // ---- expand ----
        const uint32_t p = L;
        const bool bad_col = __ldg(ign + p) != 0;
        const uint32_t heur = __ldcg(Hg + p + 1);
        const uint64_t* prow = s.recs + (uint64_t)prec * 2 * HW;
        const uint32_t a0 = __ldg(aoff + p), a1 = __ldg(aoff + p + 1);
        ChildAcc acc = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        uint64_t cells = 0;
        const int nwords = (int)((L + 63) >> 6);                          // words of the parent that hold set bits
        auto hap = [&](int which, int i0) -> uint64_t {                   // 64 bits from haplotype position i0 >= 0
            const uint64_t* hw = prow + (which ? HW : 0);
            const int wi = i0 >> 6, sh = i0 & 63;
            uint64_t x = (wi < nwords) ? (__ldcg(hw + wi) >> sh) : 0ull;
            if (sh && wi + 1 < nwords) x |= __ldcg(hw + wi + 1) << (64 - sh);
            return x;
        };
        for (uint32_t j = a0 + lane; j < a1; j += 32) {
            const ReadMeta rm = rmeta[__ldg(aidx + j)];
            score_read(a, m, rm, -(int)rm.start, (int)L, p, bad_col, hap, acc, cells);
        }
        // parent words: identical test (:163) and the source of the child copies
        bool differ = false;
        for (int wi = lane; wi < nwords; wi += 32) differ |= (__ldcg(prow + wi) != __ldcg(prow + HW + wi));
        const bool ident = !__any_sync(HP_FULL_MASK, differ);

        uint32_t nchild;
        uint32_t ctot[4], cfro[4], chet[4];
        uint32_t ca1[4], ca2[4];
        if (bad_col) {
            nchild = 1;
            ctot[0] = pfrozen + wsum(acc.tb) + heur; cfro[0] = pfrozen + wsum(acc.fb); chet[0] = phets;
            ca1[0] = 0; ca2[0] = 0;
            if (ctot[0] != mt) w.status = HP_BLOCK_ASSERT;                // :529
        } else {
            const uint32_t t01 = wsum(acc.t01), t00 = wsum(acc.t00), t11 = wsum(acc.t11);
            const uint32_t f01 = wsum(acc.f01), f00 = wsum(acc.f00), f11 = wsum(acc.f11);
            nchild = 0;
            ctot[nchild] = pfrozen + t01 + heur; cfro[nchild] = pfrozen + f01; chet[nchild] = phets + 1;
            ca1[nchild] = 0; ca2[nchild] = 1; nchild++;
            if (!ident) {
                const uint32_t t10 = wsum(acc.t10), f10 = wsum(acc.f10);
                ctot[nchild] = pfrozen + t10 + heur; cfro[nchild] = pfrozen + f10; chet[nchild] = phets + 1;
                ca1[nchild] = 1; ca2[nchild] = 0; nchild++;
            }
            ctot[nchild] = pfrozen + t00 + heur; cfro[nchild] = pfrozen + f00; chet[nchild] = phets;
            ca1[nchild] = 0; ca2[nchild] = 0; nchild++;
            ctot[nchild] = pfrozen + t11 + heur; cfro[nchild] = pfrozen + f11; chet[nchild] = phets;
            ca1[nchild] = 1; ca2[nchild] = 1; nchild++;
        }
        w.evals += nchild; w.sum_lp += (uint64_t)nchild * L; w.cells += cells * nchild;
        if (qsize + nchild > a.qcap - 32 || next_idx > 0xfffffff0u) { w.status = HP_BLOCK_QUEUE_OVERFLOW; break; }

        // ---- allocate child records: recycled slots first, then fresh ones ----
        uint32_t crec[4];
        {
            uint32_t mine = 0;
            if (lane < nchild) mine = (lane < free_top) ? __ldcg(s.freelist + free_top - 1 - lane) : rec_next + (lane - free_top);
#pragma unroll
            for (uint32_t c = 0; c < 4; c++) crec[c] = __shfl_sync(HP_FULL_MASK, mine, c);
            const uint32_t from_free = min(nchild, free_top);
            free_top -= from_free; rec_next += nchild - from_free;
        }
        // ---- write child haplotype records: parent words + the new allele bit ----
        {
            const int wl = (int)(L >> 6);
            const uint64_t bit = 1ull << (L & 63);
            for (int wi = lane; wi <= wl; wi += 32) {
                uint64_t w1 = 0, w2 = 0;
                if (wi < nwords) { w1 = __ldcg(prow + wi); w2 = __ldcg(prow + HW + wi); }
#pragma unroll
                for (uint32_t c = 0; c < 4; c++) {
                    if (c < nchild) {
                        uint64_t* crow = s.recs + (uint64_t)crec[c] * 2 * HW;
                        __stcg(crow + wi, (wi == wl && ca1[c]) ? (w1 | bit) : w1);
                        __stcg(crow + HW + wi, (wi == wl && ca2[c]) ? (w2 | bit) : w2);
                    }
                }
            }
        }
        // ---- push the queue entries ----
        const uint32_t fullmask = __ballot_sync(HP_FULL_MASK, cnt >= scap);
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) {
            if (c < nchild) {
                uint32_t target = (rr + c) & 31u;
                if (fullmask) {
                    const uint32_t room = __ballot_sync(HP_FULL_MASK, cnt < scap);
                    if (room == 0) { w.status = HP_BLOCK_QUEUE_OVERFLOW; break; }
                    if (!((room >> target) & 1u)) target = __ffs(room) - 1;
                }
                if (lane == target) {
                    const uint32_t e = lane * scap + cnt;
                    const uint64_t hi = ((uint64_t)ctot[c] << 32) | (uint64_t)(0xffffffffu - chet[c]);
                    __stcg(s.khi + e, hi); __stcg(s.kidx + e, next_idx + c); __stcg(s.klen + e, L + 1);
                    __stcg(s.kfrozen + e, cfro[c]); __stcg(s.krec + e, crec[c]);
                    if (key_less(hi, next_idx + c, cm.hi, cm.idx)) { cm.hi = hi; cm.idx = next_idx + c; cm.pos = cnt; }
                    cnt++;
                }
            }
        }
        if (w.status != HP_BLOCK_OK) break;
        rr += nchild; next_idx += nchild; qsize += nchild;
        // parent record back to the free list; tracker.add_hap(L+1) x nchild (:531, :558)
        if (lane == 0) {
            __stcg(s.freelist + free_top, prec);
            __stcg(s.lencnt + L + 1, __ldcg(s.lencnt + L + 1) + nchild);
        }
        free_top++;
        if (L + 1 >= trk_thresh) trk_total += nchild;
        __syncwarp();

        // ---- pruning bookkeeping (:564-585) ----
        while (trk_total > curr_thresh && min_progress < next_expected) {
            min_progress++;
            uint32_t dropped = 0;
            if (lane == 0) dropped = __ldcg(s.lencnt + min_progress - 1);
            dropped = __shfl_sync(HP_FULL_MASK, dropped, 0);
            trk_total -= dropped; trk_thresh = min_progress;
            if (qsize > max_queue) {
                // "full prune": every entry shorter than min_progress gets the cleared priority (cost 0)
                const uint32_t b0 = lane * scap;
                cm.hi = ~0ull; cm.idx = 0xffffffffu; cm.pos = 0;
                for (uint32_t i = 0; i < cnt; i++) {
                    uint64_t hi = __ldcg(s.khi + b0 + i);
                    const uint32_t idx = __ldcg(s.kidx + b0 + i);
                    if (__ldcg(s.klen + b0 + i) < min_progress) { hi &= 0xffffffffull; __stcg(s.khi + b0 + i, hi); }
                    if (key_less(hi, idx, cm.hi, cm.idx)) { cm.hi = hi; cm.idx = idx; cm.pos = i; }
                }
                __syncwarp();
            }
        }
    }
    if (w.status != HP_BLOCK_OK) return;

    // ---- final node (:588-628) ----
    const uint32_t frec = __ldcg(s.krec + top_slot);
    const uint64_t* frow = s.recs + (uint64_t)frec * 2 * HW;
    uint32_t phased = 0, phased_snv = 0, skipped = 0;
    const uint8_t* snv = a.is_snv + m.var_base;
    for (uint32_t i = lane; i < N; i += 32) {
        uint32_t b1 = (uint32_t)(__ldcg(frow + (i >> 6)) >> (i & 63)) & 1u;
        uint32_t b2 = (uint32_t)(__ldcg(frow + HW + (i >> 6)) >> (i & 63)) & 1u;
        if (__ldg(ign + i)) { b1 = 2; b2 = 2; skipped++; }
        else if (b1 != b2) { phased++; if (__ldg(snv + i)) phased_snv++; }
        a.out_h1[m.var_base + i] = (uint8_t)b1;
        a.out_h2[m.var_base + i] = (uint8_t)b2;
    }
    phased = wsum(phased); phased_snv = wsum(phased_snv); skipped = wsum(skipped);
    if (lane == 0) {
        uint64_t* st = a.out_stats + (uint64_t)blk * 7;
        st[0] = num_pruned;
        st[1] = __ldcg(Hg + 0);
        st[2] = top_total;
        st[3] = phased; st[4] = phased_snv; st[5] = N - phased - skipped; st[6] = skipped;
        if (top_total < __ldcg(Hg + 0)) w.status = HP_BLOCK_ASSERT;        // phase_stats.rs:163
    }
    w.status = __shfl_sync(HP_FULL_MASK, w.status, 0);
}

constexpr int kSolveWarps = 8;

__global__ void __launch_bounds__(kSolveWarps * 32, 1) astar_solve_kernel(AstarArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5;
    WarpCtx w;
    w.lane = lane_id();
    w.capl = a.sub_capl;
    const uint32_t cap = w.capl * 32;
    const size_t per_warp = (size_t)cap * 28 + 64 * 4;
    uint8_t* base = smem_raw + warp * ((per_warp + 15) & ~(size_t)15);
    w.sq_key = (uint64_t*)base;
    w.sq_h1 = w.sq_key + cap;
    w.sq_h2 = w.sq_h1 + cap;
    w.sq_frozen = (uint32_t*)(w.sq_h2 + cap);
    w.hring = w.sq_frozen + cap;

    const uint32_t gwarp = blockIdx.x * kSolveWarps + warp;
    const Slab slab = carve_slab(a.slabs + (uint64_t)gwarp * a.slab_bytes, a.qcap, a.hap_words);

    for (;;) {
        uint32_t t = 0;
        if (w.lane == 0) t = atomicAdd(a.ticket, 1u);
        t = __shfl_sync(HP_FULL_MASK, t, 0);
        if (t >= a.n_blocks) break;
        const uint32_t blk = a.order[t];
        const BlkMeta m = a.meta[blk];
        w.evals = w.sum_lp = w.pops = w.cells = 0;
        w.status = m.status;
        const uint32_t N = m.n_var;
        uint32_t* Hg = a.heur + m.var_base + blk;

        if (w.status == HP_BLOCK_OK) {
            // ---- calculate_astar_heuristic (:246-292) ----
            const uint8_t* ign = a.ignored + m.var_base;
            if (w.lane == 0) { w.hring[N & 63] = 0; Hg[N] = 0; }
            __syncwarp();
            uint32_t clip = 1;
            uint64_t badwin = 0;
            for (uint32_t v = N; v-- > 0;) {
                const uint32_t bad_v = __ldg(ign + v);
                badwin = (badwin << 1) | (bad_v ? 1ull : 0ull);
                const uint2 r = sub_solve(a, m, w, v, clip, badwin, blk);
                if (w.status != HP_BLOCK_OK) break;
                const uint32_t est = r.x, solved = r.y;
                if (solved < min(clip, 2u)) { w.status = HP_BLOCK_ASSERT; break; }          // :268
                const uint32_t hnext = w.hring[(v + 1) & 63];
                uint32_t hv;
                if (bad_v) hv = hnext;
                else {
                    if (est < hnext) { w.status = HP_BLOCK_ASSERT; break; }                  // :284
                    hv = est;
                }
                __syncwarp();
                if (w.lane == 0) { w.hring[v & 63] = hv; __stcg(Hg + v, hv); }
                __syncwarp();
                clip = min(solved + 1, HP_MAX_SEGMENT);                                      // :288
            }
        }
        if (w.status == HP_BLOCK_OK) main_solve(a, m, w, slab, blk, Hg);

        // ---- per-block outputs ----
        __syncwarp();
        if (w.status != HP_BLOCK_OK) {
            if (w.lane < 7) a.out_stats[(uint64_t)blk * 7 + w.lane] = 0;
        }
        if (a.out_heur && w.status == HP_BLOCK_OK)
            for (uint32_t i = w.lane; i <= N; i += 32) a.out_heur[m.var_base + blk + i] = __ldcg(Hg + i);
        if (a.out_counters) {
            const uint64_t cells = __reduce_add_sync(HP_FULL_MASK, (uint32_t)(w.cells & 0xffffffffu)) +
                                   ((uint64_t)__reduce_add_sync(HP_FULL_MASK, (uint32_t)(w.cells >> 32)) << 32);
            if (w.lane == 0) {
                uint64_t* c = a.out_counters + (uint64_t)blk * 4;
                c[0] = w.evals; c[1] = cells; c[2] = w.sum_lp; c[3] = w.pops;
            }
        }
        if (w.lane == 0) a.out_status[blk] = w.status;
    }
}

}  // namespace hp

// ---- host-side launchers (called from hp_api.cu) ---------------------------------------------------------------
namespace hp {

size_t astar_smem_bytes(uint32_t sub_capl) {
    const size_t per_warp = ((size_t)sub_capl * 32 * 28 + 64 * 4 + 15) & ~(size_t)15;
    return per_warp * kSolveWarps;
}
int astar_solve_warps() { return kSolveWarps; }
uint64_t astar_slab_bytes(uint32_t qcap, uint32_t hap_words) { return slab_bytes_for(qcap, hap_words); }

cudaError_t launch_astar_prep(const PrepArgs& pa, cudaStream_t stream) {
    if (pa.n_blocks == 0) return cudaSuccess;
    astar_prep_kernel<<<pa.n_blocks, kPrepThreads, 0, stream>>>(pa);
    return cudaGetLastError();
}

cudaError_t launch_astar_solve(const AstarArgs& a, int n_ctas, cudaStream_t stream) {
    const size_t smem = astar_smem_bytes(a.sub_capl);
    cudaError_t e = cudaFuncSetAttribute(astar_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    astar_solve_kernel<<<n_ctas, kSolveWarps * 32, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace hp
