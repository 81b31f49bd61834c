// wfa_kernels.cu -- graph-WFA realignment on sm_100a: kernel, host-side graph flattening and the C ABI entry points.
//
// Replaces (results identical): WFAGraph::from_reference_variants_with_hom (src/wfa_graph.rs:119-284, host side
// here), WFAGraph::edit_distance_with_pruning (src/wfa_graph.rs:350-650, the kernel) and the traversed-nodes ->
// allele/qual row glue of global_realignment (src/read_parsing.rs:790-835, kernel epilogue).
//
// Kernel design: persistent, ONE WARP PER ALIGNMENT JOB.  The reference keeps, per edit distance,
// a hash map node -> diagonal -> [(offset, set id)] and reduces every (node, diagonal) to "furthest offset after
// extension + union of the sets that reach it".  Extension only depends on (node, diagonal, offset), so here a wave
// is extended when it is PUSHED and immediately reduced into its (node, diagonal) slot of a per-warp open-addressing
// hash table in HBM/L2: slot = {max-front record, and for each of the two live ED generations: end offset +
// traversed-node bitset}.  Per ED the warp walks the active nodes in topological order (a 1024- or 4096-bit node mask lives
// in registers, one or four words per lane); lanes = the node's live diagonals.  Pushes are issued in phases
// (diag-1, diag, diag+1, then one phase per child) so that no two lanes of a phase touch the same slot.
// Results do not depend on the visiting order of diagonals (everything cross-diagonal is a max or a set union).
#include <algorithm>
#include <cstring>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "hp_host.h"

namespace hp {

// ---------------------------------------------------------------------------------------------------------------
// device-side graph / job layout
// ---------------------------------------------------------------------------------------------------------------
struct WfaNode {
    uint64_t seq_off;      // offset into the byte pool selected by src
    uint32_t len;
    uint32_t src;          // 0 = reference bytes, 1 = variant allele bytes, 2 = explicit sequence pool
};

struct WfaJob {
    uint64_t node_base;    // first node in the node array; child_off / amap_off rows start at node_base + job
    uint64_t read_off;     // into read bytes
    uint64_t row_off;      // output row start
    uint32_t n_nodes;
    uint32_t read_len;
    uint32_t row_len;
    uint32_t het_lo;       // first het variant (index into the variant table) of the row
    int32_t  status;       // pre-set by the host for skipped / invalid jobs, else HP_WFA_OK
    uint32_t pad;
};

struct WfaArgs {
    uint32_t n_jobs;
    const WfaJob* jobs;
    const uint32_t* order;         // processing order: longest reads first (shorter tail of the persistent kernel)
    const WfaNode* nodes;
    const uint32_t* child_off;     // CSR rows per job: n_nodes + 1 entries starting at node_base + job index
    const uint32_t* child_idx;     // job-relative node ids
    const uint32_t* amap_off;      // same row layout as child_off
    const uint32_t* amap;          // (het index relative to het_lo) << 1 | allele
    const uint8_t* reference;
    const uint8_t* allele_bytes;
    const uint8_t* seq_pool;
    const uint8_t* read_bytes;
    const uint8_t* vtype;          // variant table types (quality table)
    uint64_t prune_distance;       // UINT64_MAX = disabled
    uint32_t max_edit_distance;
    // workspace
    uint8_t* slabs;
    uint64_t slab_bytes;
    uint32_t table_cap;            // hash slots per warp (power of two)
    uint32_t set_words_max;        // set words the slab layout was sized for
    uint32_t* ticket;
    uint32_t* epochs;              // [warps of the launch] job counter of each warp's table (persists across launches)
    const uint32_t* noise;         // optional [n_jobs]: probe estimate of each job (jobs at or above kNoisyProbe try the piece filter)
    // outputs
    int32_t*  out_status;
    uint32_t* out_score;
    uint8_t*  out_alleles;
    uint8_t*  out_quals;
    uint32_t* out_n_nodes;         // optional
    uint64_t* out_traversed;       // optional
    uint32_t  trav_words;
    uint64_t* out_counters;        // optional [n_jobs * 4]
    int dbg_times;                 // profiling aid: counters 2 and 3 carry the job's start / end time (ns) and the filter stays on
};

constexpr uint32_t kWfaMaxNodes = 4096;          // node activity mask: MW 32-bit words per lane (kernel variants MW = 1, 4)
constexpr uint32_t kNil = 0xffffffffu;
constexpr uint32_t kNoisyProbe = 12;   // probe estimate (edits per 512 bases, over the start of the read) from which a job is scheduled first
constexpr uint32_t kPiece = 10;                              // piece filter: bases per read piece (2 bits each -> 2^20 codes)
constexpr uint32_t kPieceSlots = 4096;                       // ... hash set of the read's pieces (u32 entries, at most half full)
constexpr uint64_t kPieceTableBytes = 4ull * kPieceSlots;
constexpr uint64_t kEmptyKey = ~0ull;
// One warp per CTA: a warp that is stuck with a long job (a read that runs to MaxEditDistance takes ~25 ms) then holds on to its own
// registers only, and the CTAs of the next launch (another context's chunk, the A* of this one) move in beside it.  With 8 warps
// per CTA one straggler kept the other 7 warps' share of the SM idle until it was done.
#ifndef HP_WFA_WARPS
#define HP_WFA_WARPS 1
#endif
constexpr int kWfaWarps = HP_WFA_WARPS;
#ifndef HP_WFA_PRIVATE_STEPS
#define HP_WFA_PRIVATE_STEPS 2
#endif
constexpr int kPrivateSteps = HP_WFA_PRIVATE_STEPS;   // 8-base extension steps a lane takes alone before the warp finishes its run
#ifndef HP_WFA_CTAS_PER_SM
#define HP_WFA_CTAS_PER_SM (16 / HP_WFA_WARPS)
#endif

__host__ __device__ inline uint32_t wfa_slot_stride(uint32_t set_words) { return 32u + 16u * set_words; }

// slab layout: keys[cap] u64 | slots[cap * stride] | items[2][cap] u32 | seg_start[2][1024] | seg_len[2][1024] |
//              late_head[1024] | chunks[cap/8 * 34] u32 | rowtmp[4096] u32 | piece_tbl[4096] u32
__host__ __device__ inline uint64_t wfa_slab_bytes(uint32_t cap, uint32_t set_words) {
    uint64_t b = (uint64_t)cap * 8 + (uint64_t)cap * wfa_slot_stride(set_words) + 2ull * cap * 4 + 5ull * kWfaMaxNodes * 4 +
                 (uint64_t)(cap / 8) * 34 * 4 + 4096ull * 4 + kPieceTableBytes;
    return (b + 255) & ~255ull;
}

struct WfaSlab {
    unsigned long long* keys;
    uint8_t* slots;
    uint32_t* items[2];
    uint32_t* seg_start[2];
    uint32_t* seg_len[2];
    uint32_t* late_head;
    uint32_t* chunks;      // chunk c: [next, count, 32 items]
    uint32_t* rowtmp;
    uint32_t* piece_tbl;   // piece filter: hash set of the read's pieces
};

__device__ __forceinline__ WfaSlab wfa_carve(uint8_t* p, uint32_t cap, uint32_t set_words) {
    WfaSlab s;
    s.keys = (unsigned long long*)p; p += (uint64_t)cap * 8;
    s.slots = p; p += (uint64_t)cap * wfa_slot_stride(set_words);
    s.items[0] = (uint32_t*)p; p += (uint64_t)cap * 4;
    s.items[1] = (uint32_t*)p; p += (uint64_t)cap * 4;
    s.seg_start[0] = (uint32_t*)p; p += kWfaMaxNodes * 4;
    s.seg_start[1] = (uint32_t*)p; p += kWfaMaxNodes * 4;
    s.seg_len[0] = (uint32_t*)p; p += kWfaMaxNodes * 4;
    s.seg_len[1] = (uint32_t*)p; p += kWfaMaxNodes * 4;
    s.late_head = (uint32_t*)p; p += kWfaMaxNodes * 4;
    s.chunks = (uint32_t*)p; p += (uint64_t)(cap / 8) * 34 * 4;
    s.rowtmp = (uint32_t*)p; p += 4096ull * 4;
    s.piece_tbl = (uint32_t*)p;
    return s;
}

// slot fields (byte offsets inside a slot)
//   0 record u32 | 4 tag[0] u32 | 8 tag[1] u32 | 12 end[0] u32 | 16 end[1] u32 | 20 diag i32 | 24 node u32 | 28 pad
//   32 set[0][SW] u64 | 32 + 8*SW set[1][SW] u64
struct SlotRef {
    uint8_t* p;
    uint32_t sw;
    __device__ __forceinline__ uint32_t& record() const { return *(uint32_t*)(p + 0); }
    __device__ __forceinline__ uint32_t& tag(int g) const { return *(uint32_t*)(p + 4 + 4 * g); }
    __device__ __forceinline__ uint32_t& end(int g) const { return *(uint32_t*)(p + 12 + 4 * g); }
    __device__ __forceinline__ int32_t& diag() const { return *(int32_t*)(p + 20); }
    __device__ __forceinline__ uint32_t& node() const { return *(uint32_t*)(p + 24); }
    __device__ __forceinline__ uint64_t* set(int g) const { return (uint64_t*)(p + 32 + 8 * sw * g); }
};

struct WfaCtx {
    const WfaArgs* a;
    WfaSlab s;
    uint32_t lane, cap, sw, stride;
    uint32_t epoch;                // this warp's job counter (tags the hash keys)
    const WfaNode* nodes;          // this job's nodes
    const uint8_t* read;
    uint32_t read_len;
    uint32_t used;                 // hash slots in use (warp-uniform)
    uint32_t n_chunks;             // chunk arena bump pointer (warp-uniform)
    bool overflow;
    uint64_t n_cmp, n_waves, n_setops;
};

__device__ __forceinline__ const uint8_t* node_seq(const WfaArgs& a, const WfaNode& n) {
    const uint8_t* base = n.src == 0 ? a.reference : (n.src == 1 ? a.allele_bytes : a.seq_pool);
    return base + n.seq_off;
}

// 64 bits starting at an arbitrary byte address (little endian), from one or two aligned 64-bit loads
__device__ __forceinline__ uint64_t ld64_unaligned(const uint8_t* p) {
    const uintptr_t u = (uintptr_t)p;
    const uint64_t* q = (const uint64_t*)(u & ~(uintptr_t)7);
    const uint32_t sh = (uint32_t)(u & 7u) * 8u;
    const uint64_t lo = __ldg(q);
    if (sh == 0) return lo;
    return (lo >> sh) | (__ldg(q + 1) << (64u - sh));
}

// key = epoch (20 bits) | node (12 bits) | diagonal (32 bits).  The epoch is the warp's job counter: keys left behind by earlier
// jobs read as EMPTY, so the table is never cleared between jobs (it is zeroed once when the workspace is laid out; epochs
// start at 1).  The slot index keeps blocks of neighbouring diagonals of a node together (the live diagonals of a node are a contiguous
// range and the lanes of a batch walk them roughly in order), so a batch touches a few runs of slots instead of 32 random lines.
constexpr uint32_t kEpochShift = 44, kEpochMax = 0xffffeu;
__device__ __forceinline__ uint64_t wfa_key(uint32_t epoch, uint32_t node, int32_t diag) {
    return ((uint64_t)epoch << kEpochShift) | ((uint64_t)node << 32) | (uint32_t)diag;
}
#ifndef HP_WFA_HASH_BLOCK
#define HP_WFA_HASH_BLOCK 4
#endif
__device__ __forceinline__ uint32_t wfa_hash(uint32_t node, int32_t diag) {
    // blocks of 2^HP_WFA_HASH_BLOCK neighbouring diagonals of a node stay together, the blocks are scattered uniformly
    uint64_t k = ((uint64_t)node << 32) | (uint32_t)(diag >> HP_WFA_HASH_BLOCK);
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 29;
    return ((uint32_t)k << HP_WFA_HASH_BLOCK) | ((uint32_t)diag & ((1u << HP_WFA_HASH_BLOCK) - 1u));
}

// Pushes one wave per participating lane (act) of this warp into generation tag `gen_tag` (parity gp) of slot
// (node, diag): extend from `offset`, then reduce into the slot (furthest end wins, union of sets on ties).
// Source set = src_set (sw words) plus optionally bit `add_bit`.  Returns in `first` whether this lane created the
// slot's data for that generation (it then has to be listed), and the slot index in `slot_out`.
// All lanes of a call must target distinct (node, diag) pairs.
__device__ __forceinline__ void wfa_push(WfaCtx& c, bool act, uint32_t node, int32_t diag, uint32_t offset,
                                         const uint64_t* src_set, uint32_t add_bit, uint32_t gen_tag, int gp, bool& first,
                                         uint32_t& slot_out) {
    first = false; slot_out = kNil;
    uint32_t n_created = 0;
    // ---- extend (wfa_graph.rs:454-459) ----
    // Eight bases per step (unaligned 64-bit windows assembled from aligned loads; every byte pool is followed by at least
    // 16 readable bytes of staging slack).  Two private steps per lane settle the waves that stop at once (nearly all of them
    // on a wrong diagonal); a wave that is still running after 16 bases is on a matching diagonal and is finished by the whole
    // warp, 256 bases per step.  n_cmp counts the byte compares of the reference: the matches plus the mismatch that stops
    // the run.
    uint32_t off = offset;
    uint32_t pos = (uint32_t)(diag + (int32_t)offset);
    uint32_t node_len = 0;
    const uint8_t* seq = nullptr;
    bool running = false;
    if (act) {
        const WfaNode nd = c.nodes[node];
        seq = node_seq(*c.a, nd);
        node_len = nd.len;
        running = off < node_len && pos < c.read_len;
#pragma unroll 1
        for (int step = 0; step < kPrivateSteps && running; step++) {
            const uint32_t rem = min(node_len - off, c.read_len - pos);
            uint64_t x = ld64_unaligned(seq + off) ^ ld64_unaligned(c.read + pos);
            if (rem < 8u) x &= (1ull << (8u * rem)) - 1ull;
            if (x) {
                const uint32_t adv = (uint32_t)(__ffsll((long long)x) - 1) >> 3;
                off += adv; pos += adv; c.n_cmp += adv + 1;
                running = false;
            } else {
                const uint32_t adv = min(rem, 8u);
                off += adv; pos += adv; c.n_cmp += adv;
                running = off < node_len && pos < c.read_len;
            }
        }
    }
    for (uint32_t rm = __ballot_sync(HP_FULL_MASK, running); rm; rm &= rm - 1) {
        const uint32_t src = __ffs(rm) - 1;
        const uint8_t* sq = (const uint8_t*)__shfl_sync(HP_FULL_MASK, (unsigned long long)(uintptr_t)(seq + off), src);
        const uint8_t* rd = c.read + __shfl_sync(HP_FULL_MASK, pos, src);
        uint32_t rem = __shfl_sync(HP_FULL_MASK, min(node_len - off, c.read_len - pos), src);
        uint32_t total = 0, stopped = 0;
        while (rem) {
            const uint32_t my0 = c.lane * 8u;
            uint64_t x = 0;
            if (my0 < rem) {
                x = ld64_unaligned(sq + my0) ^ ld64_unaligned(rd + my0);
                if (rem - my0 < 8u) x &= (1ull << (8u * (rem - my0))) - 1ull;
            }
            const uint32_t bm = __ballot_sync(HP_FULL_MASK, x != 0);
            if (bm) {
                const uint32_t adv = my0 + ((uint32_t)(__ffsll((long long)x) - 1) >> 3);
                total += __shfl_sync(HP_FULL_MASK, adv, __ffs(bm) - 1);
                stopped = 1;
                break;
            }
            const uint32_t step = min(rem, 256u);
            total += step; rem -= step; sq += step; rd += step;
        }
        if (c.lane == src) { off += total; pos += total; c.n_cmp += total + stopped; }
    }
    // ---- find or insert the slot ----
    // The table belongs to this warp alone, so no atomic is needed (an atomicCAS is a round trip to the L2 for every new slot):
    // lanes that meet the same empty entry are told apart by __match_any_sync, the lowest of them takes the entry with a plain
    // store and the others look at it again.
    const uint64_t key = wfa_key(c.epoch, node, diag);
    uint32_t h = wfa_hash(node, diag) & (c.cap - 1);
    uint32_t slot = kNil;
    bool created = false;
    bool need = act;
    for (uint32_t probe = 0; probe < 2u * c.cap && __any_sync(HP_FULL_MASK, need); probe++) {
        bool empty = false;
        if (need) {
            const unsigned long long k = c.s.keys[h];
            empty = (uint32_t)(k >> kEpochShift) != c.epoch;                 // never used, or left by an earlier job
            if (!empty) { if (k == key) { slot = h; need = false; } else h = (h + 1) & (c.cap - 1); }
        }
        const uint32_t em = __ballot_sync(HP_FULL_MASK, empty);
        if (empty) {
            const uint32_t peers = __match_any_sync(em, h);
            if (c.lane == (uint32_t)(__ffs(peers) - 1)) { c.s.keys[h] = key; slot = h; created = true; need = false; }
        }
        __syncwarp();
    }
    if (act) {
        if (slot != kNil) {
            SlotRef sr{c.s.slots + (uint64_t)slot * c.stride, c.sw};
            if (created) { sr.record() = 0; sr.tag(0) = kNil; sr.tag(1) = kNil; sr.diag() = diag; sr.node() = node; }
            uint64_t* dst = sr.set(gp);
            if (sr.tag(gp) != gen_tag) {
                sr.tag(gp) = gen_tag; sr.end(gp) = off;
                for (uint32_t w = 0; w < c.sw; w++) dst[w] = src_set[w];
                if (add_bit != kNil) dst[add_bit >> 6] |= 1ull << (add_bit & 63);
                first = true;
            } else if (off > sr.end(gp)) {
                sr.end(gp) = off;
                for (uint32_t w = 0; w < c.sw; w++) dst[w] = src_set[w];
                if (add_bit != kNil) dst[add_bit >> 6] |= 1ull << (add_bit & 63);
            } else if (off == sr.end(gp)) {
                for (uint32_t w = 0; w < c.sw; w++) dst[w] |= src_set[w];
                if (add_bit != kNil) dst[add_bit >> 6] |= 1ull << (add_bit & 63);
                c.n_setops++;
            }
            slot_out = slot;
        } else {
            c.overflow = true;
        }
        if (created) n_created = 1;
    }
    c.used += __popc(__ballot_sync(HP_FULL_MASK, n_created != 0));      // table occupancy (warp-uniform)
    if (c.used > c.cap / 2) c.overflow = true;
}

// ---- piece filter: an exact proof that a read cannot align within max_edit_distance -------------------------------------
// Cut the read into disjoint pieces of kPiece bases.  An alignment of the read to ANY path of the graph with k unit-cost edits
// leaves all but at most k pieces untouched, and an untouched piece occurs verbatim in that path.  So the number of pieces that
// occur in NO path of the graph is a lower bound of the edit distance the reference computes (wfa_graph.rs:350-650, global in
// the read and in the graph; pruning can only make its result larger).  If the bound exceeds max_edit_distance the reference
// returns MaxEditDistance (:645-648) and so do we, without walking 500 edit-distance levels (~25 ms of one warp per read).
// The read's pieces (the first 2048 of them: a subset still gives a lower bound) go into a small hash set with their
// multiplicity (16 KB per warp, L2-resident); then every kPiece-mer of every path of the graph is looked up and its entry marked:
// those inside a node by all lanes side by side, those that start in the last kPiece - 1 bases of a node by a bounded
// depth-first walk through its descendants, one (node, start) pair per lane.  What stays unmarked is missing.  A piece is
// identified by a code that is a function of its bytes alone (two bits of each byte), so a piece that does occur always finds
// its entry; bytes outside ACGT only make codes collide, which can hide a missing piece but never invent one.  A graph whose
// walk exceeds its budget gives no verdict (the alignment runs as usual).  Only reads whose probe estimate, scaled to the read,
// reaches max_edit_distance try it.
// entry: bits 0-20 code + 1 (0 = empty) | bits 21-30 multiplicity | bit 31 seen in the graph.  A multiplicity that overflows
// runs into the "seen" bit and beyond: the entry then counts for less than it should, never for more.
__device__ __forceinline__ uint32_t base2(uint32_t b) { return (b >> 1) & 3u; }
__device__ __forceinline__ uint32_t pack8(uint64_t w) {          // base i of the 8 bytes -> bits [2i, 2i + 2)
    uint64_t x = (w >> 1) & 0x0303030303030303ull;
    x = (x | (x >> 6)) & 0x000F000F000F000Full;
    x = (x | (x >> 12)) & 0x000000FF000000FFull;
    x = (x | (x >> 24)) & 0xFFFFull;
    return (uint32_t)x;
}
static_assert(kPiece == 10, "piece_code packs 8 + 2 bases");
__device__ __forceinline__ uint32_t piece_code(const uint8_t* p) {   // code of p[0 .. 10): reads up to p + 16 (staging slack)
    const uint64_t hi = ld64_unaligned(p + 8);
    return pack8(ld64_unaligned(p)) | (base2((uint32_t)hi) << 16) | (base2((uint32_t)(hi >> 8)) << 18);
}
__device__ __forceinline__ uint32_t piece_hash(uint32_t code) { return (code * 0x9E3779B1u) >> 20; }   // 12 bits
static_assert(kPieceSlots == 4096, "piece_hash returns 12 bits");
constexpr uint32_t kPieceKeyMask = 0x1FFFFFu, kPieceOne = 1u << 21, kPieceSeen = 1u << 31;

// marks the entry of `code`, if the read has such a piece (the table is read at the L2: it was filled with atomics)
__device__ __forceinline__ void piece_mark(uint32_t* tbl, uint32_t code) {
    const uint32_t key = code + 1u;
    for (uint32_t h = piece_hash(code);; h = (h + 1u) & (kPieceSlots - 1u)) {
        const uint32_t v = __ldcg(&tbl[h]);
        if (v == 0u) return;
        if ((v & kPieceKeyMask) == key) { if (!(v & kPieceSeen)) atomicOr(&tbl[h], kPieceSeen); return; }
    }
}

__device__ bool wfa_piece_filter(WfaCtx& c, const WfaArgs& a, const uint32_t* child_off, uint32_t n_nodes) {
    const uint32_t lane = c.lane;
    const uint32_t pieces = min(c.read_len / kPiece, kPieceSlots / 2u);
    if (pieces <= a.max_edit_distance) return false;
    uint32_t* tbl = c.s.piece_tbl;
    for (uint32_t i = lane; i < kPieceSlots / 4u; i += 32) __stcg(reinterpret_cast<uint4*>(tbl) + i, make_uint4(0, 0, 0, 0));
    __threadfence_block();
    __syncwarp();
    // the read's pieces, with multiplicity
    for (uint32_t t = lane; t < pieces; t += 32) {
        const uint32_t code = piece_code(c.read + (uint64_t)t * kPiece), key = code + 1u;
        for (uint32_t h = piece_hash(code);; h = (h + 1u) & (kPieceSlots - 1u)) {
            const uint32_t old = atomicCAS(&tbl[h], 0u, key | kPieceOne);
            if (old == 0u) break;
            if ((old & kPieceKeyMask) == key) { atomicAdd(&tbl[h], kPieceOne); break; }
        }
    }
    __threadfence_block();
    __syncwarp();
    // pieces inside a node
    for (uint32_t u0 = 0; u0 < n_nodes; u0 += 32) {
        WfaNode mine{0, 0, 0};
        if (u0 + lane < n_nodes) mine = c.nodes[u0 + lane];
        const uint8_t* my_seq = node_seq(a, mine);
        const uint32_t cnt = min(32u, n_nodes - u0);
        for (uint32_t k = 0; k < cnt; k++) {
            const uint32_t len = __shfl_sync(HP_FULL_MASK, mine.len, k);
            const uint8_t* seq = (const uint8_t*)__shfl_sync(HP_FULL_MASK, (unsigned long long)(uintptr_t)my_seq, k);
            if (len < kPiece) continue;
            for (uint32_t i = lane; i + kPiece <= len; i += 32) piece_mark(tbl, piece_code(seq + i));
        }
    }
    // pieces that start in the last kPiece - 1 bases of a node and continue into its descendants
    constexpr int kMaxDepth = 12;
    constexpr uint32_t kWalkBudget = 256;
    bool giveup = false;
    const uint32_t n_starts = n_nodes * (kPiece - 1);
    for (uint32_t s0 = 0; s0 < n_starts; s0 += 32) {
        const uint32_t s = s0 + lane;
        if (s < n_starts) {
            const uint32_t u = s / (kPiece - 1), back = s % (kPiece - 1) + 1;      // the piece starts `back` bases before the node's end
            const WfaNode nd = c.nodes[u];
            if (back <= nd.len) {
                const uint8_t* seq = node_seq(a, nd) + (nd.len - back);
                uint32_t code = 0;
                for (uint32_t k = 0; k < back; k++) code |= base2(seq[k]) << (2 * k);
                uint32_t st_node[kMaxDepth], st_edge[kMaxDepth], st_code[kMaxDepth], st_have[kMaxDepth];
                int depth = 0;
                uint32_t walked = 0;
                st_node[0] = u; st_edge[0] = child_off[u]; st_code[0] = code; st_have[0] = back;
                while (depth >= 0) {
                    const uint32_t n = st_node[depth];
                    if (st_edge[depth] == child_off[n + 1]) { depth--; continue; }
                    const uint32_t ch = a.child_idx[st_edge[depth]++];
                    if (++walked > kWalkBudget) { giveup = true; break; }
                    const WfaNode cn = c.nodes[ch];
                    const uint8_t* cs = node_seq(a, cn);
                    uint32_t cc = st_code[depth], have = st_have[depth];
                    const uint32_t t = min(kPiece - have, cn.len);
                    for (uint32_t k = 0; k < t; k++) cc |= base2(cs[k]) << (2 * (have + k));
                    have += t;
                    if (have == kPiece) { piece_mark(tbl, cc); continue; }
                    if (depth + 1 >= kMaxDepth) { giveup = true; break; }
                    depth++;
                    st_node[depth] = ch; st_edge[depth] = child_off[ch]; st_code[depth] = cc; st_have[depth] = have;
                }
            }
        }
        giveup = __any_sync(HP_FULL_MASK, giveup);
        if (giveup) return false;
    }
    __threadfence_block();
    __syncwarp();
    uint32_t missing = 0;
    for (uint32_t i = lane; i < kPieceSlots; i += 32) {
        const uint32_t v = __ldcg(&tbl[i]);
        if (v != 0u && !(v & kPieceSeen)) missing += (v >> 21) & 0x3FFu;
    }
    missing = __reduce_add_sync(HP_FULL_MASK, missing);
    return missing > a.max_edit_distance;
}

template <int MW>
__global__ void __launch_bounds__(kWfaWarps * 32, HP_WFA_CTAS_PER_SM) wfa_align_kernel(WfaArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gwarp = blockIdx.x * kWfaWarps + (threadIdx.x >> 5);
    WfaCtx c;
    c.a = &a; c.lane = lane; c.cap = a.table_cap;
    uint8_t* slab = a.slabs + (uint64_t)gwarp * a.slab_bytes;

    for (;;) {
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(a.ticket, 1u);
        j = __shfl_sync(HP_FULL_MASK, j, 0);
        if (j >= a.n_jobs) break;
        j = a.order[j];
        const WfaJob job = a.jobs[j];
        unsigned long long t_start = 0;
        if (a.dbg_times) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
        int status = job.status;
        uint32_t score = 0;
        c.n_cmp = c.n_waves = c.n_setops = 0;
        const uint32_t n_nodes = job.n_nodes;
        const uint32_t sw = (n_nodes + 63) / 64;
        if (status == HP_WFA_OK && n_nodes > 1024u * MW) status = HP_WFA_GRAPH_TOO_LARGE;
        if (status == HP_WFA_OK && (n_nodes == 0 || sw > a.set_words_max)) status = HP_WFA_WORKSPACE_OVERFLOW;

        // output row defaults: NoOverlap / 0 (read_parsing.rs:790, 803)
        for (uint32_t i = lane; i < job.row_len; i += 32) { a.out_alleles[job.row_off + i] = HP_ALLELE_NOOVERLAP; a.out_quals[job.row_off + i] = 0; }
        if (a.out_traversed) for (uint32_t i = lane; i < a.trav_words; i += 32) a.out_traversed[(uint64_t)j * a.trav_words + i] = 0;

        if (status == HP_WFA_OK) {
            c.sw = sw; c.stride = wfa_slot_stride(sw);
            c.s = wfa_carve(slab, c.cap, sw);
            c.nodes = a.nodes + job.node_base;
            c.read = a.read_bytes + job.read_off;
            c.read_len = job.read_len;
            c.used = 0; c.n_chunks = 0; c.overflow = false;
            const uint32_t* child_off = a.child_off + job.node_base + j;
            const uint32_t sink = n_nodes - 1;
            const uint32_t max_chunks = c.cap / 8;

            // reads whose probe estimate, scaled to the read, reaches 0.9 max_edit_distance (a try costs ~1 ms, a proof saves ~25):
            // try to prove MaxEditDistance without aligning (never in the counting variant, whose work counters are the
            // reference's)
            if (a.noise && (!a.out_counters || a.dbg_times) && a.noise[j] >= kNoisyProbe &&
                (uint64_t)a.noise[j] * c.read_len * 10ull >= (uint64_t)a.max_edit_distance * 512ull * 9ull &&
                wfa_piece_filter(c, a, child_off, n_nodes)) {
                status = HP_WFA_MAX_EDIT_DISTANCE; score = a.max_edit_distance;
                if (lane == 0) atomicAdd(a.ticket + 1, 1u);
            }
          if (status == HP_WFA_OK) {
            // a fresh epoch instead of a table clear (the clear only happens when the 20-bit counter wraps)
            {
                uint32_t e = 0;
                if (lane == 0) { e = a.epochs[gwarp] + 1u; if (e > kEpochMax) e = 0; a.epochs[gwarp] = e ? e : 1u; }
                e = __shfl_sync(HP_FULL_MASK, e, 0);
                if (e == 0) { for (uint32_t i = lane; i < c.cap; i += 32) c.s.keys[i] = 0ull; e = 1u; }
                c.epoch = e;
            }
            for (uint32_t i = lane; i < n_nodes; i += 32) {
                c.s.seg_len[0][i] = 0; c.s.seg_len[1][i] = 0; c.s.late_head[i] = kNil;
            }
            __syncwarp();

            // node activity masks: word w of lane l owns nodes [1024w + 32l, 1024w + 32l + 32)
            uint32_t act_cur[MW], act_next[MW];
#pragma unroll
            for (int w = 0; w < MW; w++) { act_cur[w] = 0; act_next[w] = 0; }
            uint32_t n_items_next = 0;             // items appended to the next generation's list so far (uniform)

            // ---- initial wave: node 0, diagonal 0, offset 0, set {0} (wfa_graph.rs:366-378) ----
            {
                uint64_t* seed = (uint64_t*)c.s.rowtmp;       // scratch: an all-zero set
                for (uint32_t w = lane; w < sw; w += 32) seed[w] = 0;
                __syncwarp();
                bool first; uint32_t slot;
                wfa_push(c, lane == 0, 0u, 0, 0u, seed, 0u, 0u, 0, first, slot);
                if (lane == 0) { c.s.items[0][0] = slot; c.s.seg_start[0][0] = 0; c.s.seg_len[0][0] = 1; act_cur[0] = 1u; }
                c.used = 1;
                __syncwarp();
            }

            uint32_t ed = 0, farthest = 0, min_prog = 0;
            bool done = false;
            while (!done) {
                const int g = ed & 1, gn = g ^ 1;
                n_items_next = 0;
                // ---- nodes with pending waves, ascending (wfa_graph.rs:406) ----
                for (;;) {
                    uint32_t n = kNil;
#pragma unroll
                    for (int w = 0; w < MW; w++) {
                        if (n != kNil) continue;
                        const uint32_t lanes_with = __ballot_sync(HP_FULL_MASK, act_cur[w] != 0);
                        if (lanes_with == 0) continue;
                        const uint32_t wl = __ffs(lanes_with) - 1;
                        const uint32_t word = __shfl_sync(HP_FULL_MASK, act_cur[w], wl);
                        n = 1024u * w + wl * 32 + (__ffs(word) - 1);
                        if (lane == wl) act_cur[w] &= act_cur[w] - 1;      // clear the lowest set bit
                    }
                    if (n == kNil) break;

                    const WfaNode nd = c.nodes[n];
                    const uint32_t node_len = nd.len;
                    const uint32_t c0 = child_off[n], c1 = child_off[n + 1];
                    const uint32_t own_start = c.s.seg_start[g][n], own_len = c.s.seg_len[g][n];
                    uint32_t late = c.s.late_head[n];
                    const uint32_t next_seg_start = n_items_next;
                    __syncwarp();
                    if (lane == 0) { c.s.seg_len[g][n] = 0; c.s.late_head[n] = kNil; c.s.seg_start[gn][n] = next_seg_start; }

                    // batches of up to 32 (node, diagonal) slots: first the node's own segment, then the late chunks
                    uint32_t own_done = 0;
                    for (;;) {
                        uint32_t item = kNil;
                        if (own_done < own_len) {
                            if (own_done + lane < own_len) item = c.s.items[g][own_start + own_done + lane];
                            own_done += 32;
                        } else if (late != kNil) {
                            const uint32_t* ch = c.s.chunks + (uint64_t)late * 34;
                            const uint32_t cnt = ch[1];
                            if (lane < cnt) item = ch[2 + lane];
                            late = ch[0];
                        } else break;

                        // ---- per-slot step (wfa_graph.rs:443-573) ----
                        const bool act = item != kNil;
                        SlotRef sr{c.s.slots + (uint64_t)(act ? item : 0) * c.stride, sw};
                        int32_t d = 0; uint32_t e = 0;
                        bool alive = false;
                        if (act) {
                            d = sr.diag(); e = sr.end(g);
                            c.n_waves++;
                            const uint32_t rec = sr.record();
                            const bool skip = e < rec || (int64_t)d + (int64_t)e < (int64_t)min_prog;     // :464-469
                            if (!skip) { sr.record() = e; alive = true; }
                        }
                        const uint32_t prog = alive ? (uint32_t)(d + (int32_t)e) : 0u;
                        farthest = max(farthest, __reduce_max_sync(HP_FULL_MASK, prog));                   // :474
                        const bool at_end = (e == node_len);
                        const bool read_left = act && (uint32_t)(d + (int32_t)e) < c.read_len;

                        // finished?  any wave of the sink at (end of node, end of read), skipped or not (:576-629)
                        if (n == sink) {
                            const uint32_t fin = __ballot_sync(HP_FULL_MASK, act && at_end && (uint32_t)(d + (int32_t)e) == c.read_len);
                            if (fin) {
                                const uint32_t fl = __ffs(fin) - 1;
                                const uint32_t fitem = __shfl_sync(HP_FULL_MASK, item, fl);
                                const uint64_t* fs = SlotRef{c.s.slots + (uint64_t)fitem * c.stride, sw}.set(g);
                                // ---- epilogue: traversed nodes -> allele / qual row (read_parsing.rs:790-835) ----
                                for (uint32_t i = lane; i < job.row_len; i += 32) c.s.rowtmp[i] = 0;
                                __syncwarp();
                                const uint32_t* amap_off = a.amap_off + job.node_base + j;
                                for (uint32_t nn = lane; nn < n_nodes; nn += 32) {
                                    if ((fs[nn >> 6] >> (nn & 63)) & 1ull) {
                                        for (uint32_t q = amap_off[nn]; q < amap_off[nn + 1]; q++) {
                                            const uint32_t m = a.amap[q];
                                            atomicOr(&c.s.rowtmp[m >> 1], 1u << (m & 1u));
                                        }
                                    }
                                }
                                __syncwarp();
                                for (uint32_t i = lane; i < job.row_len; i += 32) {
                                    const uint32_t seen = c.s.rowtmp[i];
                                    uint8_t al = HP_ALLELE_NOOVERLAP, ql = 0;
                                    if (seen == 3u) al = HP_ALLELE_AMBIGUOUS;
                                    else if (seen) {
                                        al = (uint8_t)(seen - 1u);
                                        const uint8_t vt = a.vtype[job.het_lo + i];
                                        // doubled global-realignment qualities (read_parsing.rs:18-22, 815-835)
                                        ql = vt == HP_VT_SNV ? 160 : (vt == HP_VT_TANDEM_REPEAT ? 80 :
                                             ((vt == HP_VT_SV_DELETION || vt == HP_VT_SV_INSERTION) ? 40 :
                                             ((vt == HP_VT_DELETION || vt == HP_VT_INSERTION || vt == HP_VT_INDEL) ? 20 : 0)));
                                    }
                                    a.out_alleles[job.row_off + i] = al; a.out_quals[job.row_off + i] = ql;
                                }
                                if (a.out_traversed)
                                    for (uint32_t w = lane; w < min(sw, a.trav_words); w += 32) a.out_traversed[(uint64_t)j * a.trav_words + w] = fs[w];
                                score = ed; done = true;
                            }
                        }
                        if (done) break;

                        const uint64_t* myset = sr.set(g);
                        // ---- pushes into the NEXT edit distance (same node) ----
                        const uint32_t tag_next = ed + 1;
                        const bool inside = alive && !at_end;
                        const bool sink_more = alive && at_end && n == sink && read_left;
#pragma unroll 1
                        for (int phase = 0; phase < 3; phase++) {
                            bool pa; int32_t pd; uint32_t po;
                            if (phase == 0) { pa = inside; pd = d - 1; po = e + 1; }                       // graph advances
                            else if (phase == 1) { pa = inside && read_left; pd = d; po = e + 1; }         // mismatch
                            else { pa = (inside && read_left) || sink_more; pd = d + 1; po = e; }          // read advances
                            if (!__any_sync(HP_FULL_MASK, pa)) continue;
                            bool first; uint32_t slot;
                            wfa_push(c, pa, n, pd, po, myset, kNil, tag_next, gn, first, slot);
                            const uint32_t fm = __ballot_sync(HP_FULL_MASK, first);
                            if (fm) {
                                const uint32_t pos = n_items_next + __popc(fm & ((1u << lane) - 1u));
                                if (first && pos < c.cap) c.s.items[gn][pos] = slot;
                                n_items_next += __popc(fm);
#pragma unroll
                                for (int w = 0; w < MW; w++) if ((n >> 10) == (uint32_t)w && lane == ((n >> 5) & 31u)) act_next[w] |= 1u << (n & 31);
                            }
                            __syncwarp();
                        }
                        // ---- propagation to the children within THIS edit distance (:527-553) ----
                        const bool prop = alive && at_end && n != sink;
                        if (__any_sync(HP_FULL_MASK, prop)) {
                            for (uint32_t q = c0; q < c1; q++) {
                                const uint32_t child = a.child_idx[q];
                                bool first; uint32_t slot;
                                wfa_push(c, prop, child, d + (int32_t)e, 0u, myset, child, ed, g, first, slot);
                                const uint32_t fm = __ballot_sync(HP_FULL_MASK, first);
                                if (fm) {
                                    const uint32_t chunk = c.n_chunks++;
                                    if (chunk < max_chunks) {
                                        uint32_t* ch = c.s.chunks + (uint64_t)chunk * 34;
                                        if (first) ch[2 + __popc(fm & ((1u << lane) - 1u))] = slot;
                                        if (lane == 0) { ch[0] = c.s.late_head[child]; ch[1] = __popc(fm); c.s.late_head[child] = chunk; }
                                    } else c.overflow = true;
#pragma unroll
                                    for (int w = 0; w < MW; w++) if ((child >> 10) == (uint32_t)w && lane == ((child >> 5) & 31u)) act_cur[w] |= 1u << (child & 31);
                                }
                                __syncwarp();
                            }
                        }
                        c.overflow = __any_sync(HP_FULL_MASK, c.overflow);
                        if (c.overflow) break;
                    }
                    if (done || c.overflow) break;
                    if (lane == 0) c.s.seg_len[gn][n] = n_items_next - next_seg_start;
                    __syncwarp();
                }
                if (done) break;
                if (c.overflow) { status = HP_WFA_WORKSPACE_OVERFLOW; break; }
                // ---- next edit distance (:633-648) ----
                ed++;
                uint32_t any_next = 0;
#pragma unroll
                for (int w = 0; w < MW; w++) { act_cur[w] = act_next[w]; any_next |= act_next[w]; act_next[w] = 0; }
                c.n_chunks = 0;
                if ((uint64_t)farthest > a.prune_distance) min_prog = farthest - (uint32_t)a.prune_distance;
                if (ed > a.max_edit_distance) { status = HP_WFA_MAX_EDIT_DISTANCE; score = a.max_edit_distance; break; }
                if (__ballot_sync(HP_FULL_MASK, any_next != 0) == 0) { status = HP_WFA_WORKSPACE_OVERFLOW; break; }   // cannot happen
            }
          }
        }
        if (lane == 0) {
            a.out_status[j] = status;
            a.out_score[j] = (status == HP_WFA_SKIPPED) ? 0xffffffffu : score;
            if (a.out_n_nodes) a.out_n_nodes[j] = n_nodes;
        }
        if (a.out_counters) {
            const uint64_t cmp = __reduce_add_sync(HP_FULL_MASK, (uint32_t)c.n_cmp);
            const uint64_t wv = __reduce_add_sync(HP_FULL_MASK, (uint32_t)c.n_waves);
            const uint64_t so = __reduce_add_sync(HP_FULL_MASK, (uint32_t)c.n_setops);
            if (lane == 0) {
                uint64_t* o = a.out_counters + (uint64_t)j * 4;
                o[0] = cmp; o[1] = wv; o[2] = so; o[3] = n_nodes;
                if (a.dbg_times) { unsigned long long t_end; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end)); o[2] = t_start; o[3] = t_end; }
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// graph construction on the device: from_reference_variants_with_hom (wfa_graph.rs:119-284),
// one thread per job, two passes -- count (nodes / edges / allele tags per job), then fill into exactly sized slices.
// Node order, edges and the allele map are those of the host builder below (and of the reference): variants by
// position (hets before homs on ties), ALT node(s) before the reference node of a locus, allele0 shares the next
// backbone node unless it is itself an ALT (index_allele0 != 0), ALT nodes rejoin the backbone at pos + ref_len.
// ---------------------------------------------------------------------------------------------------------------
struct BuildArgs {
    uint32_t n_sel;                // jobs of this run
    const uint32_t* ids;           // their indices in the caller's batch
    // variant table
    const int64_t* position;
    const uint32_t* ref_len;
    const uint64_t* a0_off;
    const uint32_t* a0_len;
    const uint64_t* a1_off;
    const uint32_t* a1_len;
    const uint8_t* index_allele0;
    const uint8_t* ignored;
    // job windows
    const uint64_t* ref_start;
    const uint64_t* ref_end;
    const uint32_t* het_lo;
    const uint32_t* het_hi;
    const uint32_t* hom_lo;
    const uint32_t* hom_hi;
    uint64_t n_reference;
    // pass 1 output: per selected job {n_nodes, n_edges, n_amap, status}
    uint32_t* counts;
    // pass 2: exactly sized slices
    const WfaJob* jobs;            // node_base set by the host from the counts
    const uint64_t* edge_base;     // [n_sel]
    const uint64_t* amap_base;     // [n_sel]
    WfaNode* nodes;
    uint32_t* child_off;
    uint32_t* child_idx;
    uint32_t* amap_off;
    uint32_t* amap;
    uint32_t* e_parent;            // scratch: edges in creation order (child ascending)
    uint32_t* e_child;
};

constexpr int kBuildCap = 64;      // open ALT branches / parents of one node / pending allele-0 tags
constexpr uint32_t kBuildOk = 0, kBuildInvalid = 1, kBuildOverflow = 2;

template <bool kFill>
__global__ void __launch_bounds__(128) wfa_graph_build_kernel(BuildArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.n_sel) return;
    const uint32_t j = a.ids[slot];
    const uint32_t het_lo = a.het_lo[j];
    uint32_t hi = het_lo, he = a.het_hi[j], mi = a.hom_lo[j], me = a.hom_hi[j];
    const uint64_t w0 = a.ref_start[j], w1 = a.ref_end[j];
    uint32_t n_nodes = 0, n_edges = 0, n_amap = 0, status = kBuildOk;
    uint64_t node_base = 0, ebase = 0, mbase = 0, row = 0;
    if (kFill) {
        if (a.counts[4 * slot + 3] != kBuildOk || (he == het_lo)) return;
        node_base = a.jobs[slot].node_base; ebase = a.edge_base[slot]; mbase = a.amap_base[slot];
        row = node_base + slot;                               // CSR rows: n_nodes + 1 entries per job
    }
    if (he == het_lo) {                                        // skipped job (read_parsing.rs:703-712): no graph
        if (!kFill) { a.counts[4 * slot] = 0; a.counts[4 * slot + 1] = 0; a.counts[4 * slot + 2] = 0; a.counts[4 * slot + 3] = kBuildOk; }
        return;
    }
    uint64_t cursor = w0;
    uint32_t attach[kBuildCap]; int n_attach = 0;              // parents of the next backbone node
    uint32_t pend[kBuildCap]; int n_pend = 0;                  // allele-0 tags waiting for the next backbone node
    uint64_t rpos[kBuildCap]; uint32_t rnode[kBuildCap]; int n_rej = 0;   // open ALT branches: (rejoin position, node)
    if (w1 > a.n_reference || w0 > w1) status = kBuildInvalid;

    auto add = [&](uint32_t src, uint64_t off, uint32_t len) -> uint32_t {      // parents = attach[]
        const uint32_t id = n_nodes++;
        if (kFill) {
            WfaNode nd; nd.seq_off = off; nd.len = len; nd.src = src;
            a.nodes[node_base + id] = nd;
            a.amap_off[row + id] = (uint32_t)(mbase + n_amap);
            for (int q = 0; q < n_attach; q++) { a.e_parent[ebase + n_edges + q] = attach[q]; a.e_child[ebase + n_edges + q] = id; }
        }
        n_edges += (uint32_t)n_attach;
        return id;
    };
    auto tag = [&](uint32_t t) { if (kFill) a.amap[mbase + n_amap] = t; n_amap++; };
    auto backbone_to = [&](uint64_t upto) -> uint32_t {                           // emits reference[cursor, upto)
        const uint32_t id = add(0u, cursor, (uint32_t)(upto - cursor));
        for (int q = 0; q < n_pend; q++) tag(pend[q]);
        n_pend = 0;
        cursor = upto;
        return id;
    };
    auto drain = [&](uint64_t limit) {                                            // ALT nodes rejoining at positions <= limit
        while (status == kBuildOk && n_rej > 0) {
            uint64_t at = rpos[0];
            for (int q = 1; q < n_rej; q++) at = rpos[q] < at ? rpos[q] : at;
            if (at > limit) break;
            if (!(at > cursor)) { status = kBuildInvalid; break; }
            const uint32_t bb = backbone_to(at);
            attach[0] = bb; n_attach = 1;
            for (int q = 0; q < n_rej;) {
                if (rpos[q] == at) {
                    if (n_attach >= kBuildCap) { status = kBuildOverflow; break; }
                    attach[n_attach++] = rnode[q];
                    rpos[q] = rpos[n_rej - 1]; rnode[q] = rnode[n_rej - 1]; n_rej--;
                } else q++;
            }
        }
    };
    while (status == kBuildOk && (hi < he || mi < me)) {
        uint32_t k; int het;
        if (hi < he && (mi >= me || a.position[hi] <= a.position[mi])) { k = hi; het = (int)(hi - het_lo); hi++; }
        else { k = mi; het = -1; mi++; }
        if (a.ignored[k] || a.position[k] < 0) continue;
        const uint64_t pos = (uint64_t)a.position[k], rl = a.ref_len[k];
        if (pos < w0 || pos + rl > w1) continue;
        drain(pos);
        if (status != kBuildOk) break;
        if (cursor < pos || n_nodes == 0) {
            const uint32_t bb = backbone_to(pos);
            attach[0] = bb; n_attach = 1;
        } else if (cursor != pos) { status = kBuildInvalid; break; }
        if (n_rej + 2 > kBuildCap || n_pend + 1 > kBuildCap) { status = kBuildOverflow; break; }
        if (a.index_allele0[k] != 0) {                          // allele0 is an ALT of a multi-allelic site
            const uint32_t alt = add(1u, a.a0_off[k], a.a0_len[k]);
            if (het >= 0) tag(((uint32_t)het << 1) | 0u);
            rpos[n_rej] = pos + rl; rnode[n_rej] = alt; n_rej++;
        } else if (het >= 0) pend[n_pend++] = ((uint32_t)het << 1) | 0u;
        const uint32_t alt = add(1u, a.a1_off[k], a.a1_len[k]);
        if (het >= 0) tag(((uint32_t)het << 1) | 1u);
        rpos[n_rej] = pos + rl; rnode[n_rej] = alt; n_rej++;
    }
    if (status == kBuildOk) drain(UINT64_MAX);
    if (status == kBuildOk && cursor > w1) status = kBuildInvalid;
    if (status == kBuildOk) {
        backbone_to(w1);
        if (n_pend != 0) status = kBuildInvalid;                // cannot happen: backbone_to consumes the pending tags
    }
    if (!kFill) {
        a.counts[4 * slot] = n_nodes; a.counts[4 * slot + 1] = n_edges; a.counts[4 * slot + 2] = n_amap; a.counts[4 * slot + 3] = status;
        return;
    }
    // children CSR: counting sort of the edges by parent (children stay in ascending order)
    a.amap_off[row + n_nodes] = (uint32_t)(mbase + n_amap);
    uint32_t* coff = a.child_off + row;
    for (uint32_t i = 0; i <= n_nodes; i++) coff[i] = 0;
    for (uint32_t e = 0; e < n_edges; e++) coff[a.e_parent[ebase + e] + 1]++;
    for (uint32_t i = 1; i <= n_nodes; i++) coff[i] += coff[i - 1];
    for (uint32_t i = 0; i <= n_nodes; i++) coff[i] += (uint32_t)ebase;
    for (uint32_t e = 0; e < n_edges; e++) a.child_idx[coff[a.e_parent[ebase + e]]++] = a.e_child[ebase + e];
    for (uint32_t i = n_nodes; i >= 1; i--) coff[i] = coff[i - 1];      // undo the fill cursors
    coff[0] = (uint32_t)ebase;
}

// ---------------------------------------------------------------------------------------------------------------
// scheduling probe: jobs that will run to MaxEditDistance take ~10x longer than the others, so they should start first.
// The first base of a job's read is aligned to the first base of its reference window (read_parsing.rs:737-741, 773),
// so the unit-cost edit distance of four 64-base pieces of the read's head against the same pieces of the window (Myers
// bit-vector, one thread per job) separates noisy reads from clean ones.  The estimate only orders the persistent
// kernel's ticket queue: results never depend on it.
// ---------------------------------------------------------------------------------------------------------------
struct ProbeArgs {
    uint32_t n_sel;
    const uint32_t* ids;
    const uint64_t* ref_start;
    const uint64_t* ref_end;
    const uint64_t* read_off;      // [n_jobs + 1] of the caller's batch
    const uint8_t* reference;
    const uint8_t* read_bytes;
    uint32_t* noise;               // [n_sel]
};

// Edit distance of the 64-byte pattern to its best-matching infix of txt[0, n_txt) (Myers' bit-vector search: the first row of the
// matrix is free, the minimum over all end positions is kept).  A pattern that sits a few bases off its expected place -- an
// indel earlier in the read -- still scores its own errors only.
__device__ __forceinline__ uint32_t probe_ed64(const uint8_t* pat, const uint8_t* txt, uint32_t n_txt) {
    uint64_t pA = 0, pC = 0, pG = 0, pT = 0;
    for (uint32_t j = 0; j < 64; j++) {
        const uint8_t c = pat[j];
        const uint64_t bit = 1ull << j;
        if (c == 'A') pA |= bit; else if (c == 'C') pC |= bit; else if (c == 'G') pG |= bit; else if (c == 'T') pT |= bit;
    }
    uint64_t Pv = ~0ull, Mv = 0;
    uint32_t score = 64, best = 64;
    for (uint32_t i = 0; i < n_txt; i++) {
        const uint8_t c = txt[i];
        const uint64_t Eq = c == 'A' ? pA : c == 'C' ? pC : c == 'G' ? pG : c == 'T' ? pT : 0ull;
        const uint64_t Xv = Eq | Mv;
        const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        uint64_t Ph = Mv | ~(Xh | Pv);
        uint64_t Mh = Pv & Xh;
        if (Ph >> 63) score++; else if (Mh >> 63) score--;
        Ph <<= 1; Mh <<= 1;
        Pv = Mh | ~(Xv | Ph); Mv = Ph & Xv;
        best = min(best, score);
    }
    return best;
}

// Scheduling estimate of a job: the first 512 read bases in eight blocks of 64, each matched against the reference window around
// its own offset (32 bases of play on either side), as edits per 512 bases.  ~0 for a clean read (plus the variants it carries),
// ~25 for a read at 5 % error.  Scheduling and the piece-filter gate only: results never depend on it.
__global__ void __launch_bounds__(128) wfa_noise_probe_kernel(ProbeArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.n_sel) return;
    const uint32_t j = a.ids[slot];
    const uint64_t rlen = a.read_off[j + 1] - a.read_off[j], wlen = a.ref_end[j] - a.ref_start[j];
    const uint8_t* rd = a.read_bytes + a.read_off[j];
    const uint8_t* rf = a.reference + a.ref_start[j];
    uint32_t est = 0, blocks = 0;
    for (uint32_t k = 0; k < 8; k++) {
        if ((k + 1) * 64ull > rlen) break;
        const uint64_t t0 = k ? 64ull * k - 32 : 0, t1 = min(wlen, (uint64_t)(64ull * k + 96));
        if (t1 <= t0) break;
        est += probe_ed64(rd + 64 * k, rf + t0, (uint32_t)(t1 - t0));
        blocks++;
    }
    a.noise[slot] = blocks ? est * 8u / blocks : 0u;
}

// ---------------------------------------------------------------------------------------------------------------
// host side: graph flattening (from_reference_variants_with_hom) and the API
// ---------------------------------------------------------------------------------------------------------------
struct FlatGraphs {
    std::vector<WfaJob> jobs;
    std::vector<WfaNode> nodes;
    std::vector<uint32_t> child_off, child_idx, amap_off, amap;
};

// Builds the graph of one job into fg.  Node order, edges and allele map follow wfa_graph.rs:119-284:
// variants by position (hets before homs on ties), ALT node(s) before the reference node of the same locus,
// allele0 shares the next backbone node unless it is itself an ALT (index_allele0 != 0), reconnects at pos+ref_len.
static bool build_job_graph(const hp_wfa_batch* b, uint32_t j, FlatGraphs& fg, WfaJob& job) {
    const hp_variant_table& vt = b->variants;
    const uint64_t w0 = b->ref_start[j], w1 = b->ref_end[j];
    if (w1 > b->n_reference || w0 > w1) return false;
    struct Item { uint32_t k; int32_t het; };
    std::vector<Item> order;
    for (uint32_t k = b->het_lo[j]; k < b->het_hi[j]; k++) order.push_back({k, (int32_t)(k - b->het_lo[j])});
    for (uint32_t k = b->hom_lo[j]; k < b->hom_hi[j]; k++) order.push_back({k, -1});
    std::stable_sort(order.begin(), order.end(), [&](const Item& x, const Item& y) { return vt.position[x.k] < vt.position[y.k]; });

    struct TmpNode { WfaNode n; std::vector<uint32_t> parents; std::vector<uint32_t> alleles; };
    std::vector<TmpNode> tn;
    auto add = [&](uint32_t src, uint64_t off, uint32_t len, const std::vector<uint32_t>& parents) -> uint32_t {
        TmpNode t; t.n.seq_off = off; t.n.len = len; t.n.src = src; t.parents = parents;
        tn.push_back(std::move(t));
        return (uint32_t)tn.size() - 1;
    };
    uint64_t cursor = w0;                                   // reference consumed so far
    std::vector<uint32_t> attach;                           // parents of the next backbone node
    std::vector<uint32_t> ref_alleles;                      // allele-0 tags waiting for the next backbone node
    using Rejoin = std::pair<uint64_t, uint32_t>;           // (rejoin position, ALT node)
    std::priority_queue<Rejoin, std::vector<Rejoin>, std::greater<Rejoin>> rejoin;
    bool ok = true;
    auto backbone_to = [&](uint64_t upto) -> uint32_t {     // emits reference[cursor, upto) as a backbone node
        const uint32_t id = add(0, cursor, (uint32_t)(upto - cursor), attach);
        if (!ref_alleles.empty()) { tn[id].alleles = ref_alleles; ref_alleles.clear(); }
        cursor = upto;
        return id;
    };
    auto drain_rejoins = [&](uint64_t limit) {              // all ALT nodes rejoining at positions <= limit
        while (ok && !rejoin.empty() && rejoin.top().first <= limit) {
            const uint64_t at = rejoin.top().first;
            if (!(at > cursor)) { ok = false; return; }
            const uint32_t backbone = backbone_to(at);
            attach.assign(1, backbone);
            while (!rejoin.empty() && rejoin.top().first == at) { attach.push_back(rejoin.top().second); rejoin.pop(); }
        }
    };
    for (const Item& it : order) {
        const uint32_t k = it.k;
        if (vt.ignored[k] || vt.position[k] < 0) continue;
        const uint64_t pos = (uint64_t)vt.position[k], rl = vt.ref_len[k];
        if (pos < w0 || pos + rl > w1) continue;
        drain_rejoins(pos);
        if (!ok) return false;
        if (cursor < pos || tn.empty()) {
            const uint32_t backbone = backbone_to(pos);
            attach.assign(1, backbone);
        } else if (cursor != pos) return false;
        if (vt.index_allele0[k] != 0) {                     // allele0 is an ALT of a multi-allelic site
            const uint32_t alt = add(1, vt.allele0_off[k], vt.allele0_len[k], attach);
            if (it.het >= 0) tn[alt].alleles.push_back(((uint32_t)it.het << 1) | 0u);
            rejoin.push({pos + rl, alt});
        } else if (it.het >= 0) ref_alleles.push_back(((uint32_t)it.het << 1) | 0u);
        const uint32_t alt = add(1, vt.allele1_off[k], vt.allele1_len[k], attach);
        if (it.het >= 0) tn[alt].alleles.push_back(((uint32_t)it.het << 1) | 1u);
        rejoin.push({pos + rl, alt});
    }
    drain_rejoins(UINT64_MAX);
    if (!ok || cursor > w1) return false;
    backbone_to(w1);
    if (!ref_alleles.empty()) return false;

    // flatten: children CSR from the parent lists
    const uint32_t n = (uint32_t)tn.size();
    job.node_base = fg.nodes.size();
    job.n_nodes = n;
    std::vector<uint32_t> deg(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) for (uint32_t p : tn[i].parents) deg[p + 1]++;
    for (uint32_t i = 0; i < n; i++) deg[i + 1] += deg[i];
    const size_t cbase = fg.child_idx.size();
    fg.child_idx.resize(cbase + deg[n]);
    std::vector<uint32_t> fill(deg.begin(), deg.end() - 1);
    for (uint32_t i = 0; i < n; i++) for (uint32_t p : tn[i].parents) fg.child_idx[cbase + fill[p]++] = i;
    for (uint32_t i = 0; i <= n; i++) fg.child_off.push_back((uint32_t)(cbase + deg[i]));
    for (uint32_t i = 0; i < n; i++) {
        fg.nodes.push_back(tn[i].n);
        fg.amap_off.push_back((uint32_t)fg.amap.size());
        for (uint32_t m : tn[i].alleles) fg.amap.push_back(m);
    }
    fg.amap_off.push_back((uint32_t)fg.amap.size());
    return true;
}

static int wfa_fail(hp_ctx* ctx, int code, const std::string& msg) { if (ctx) ctx->err = msg; return code; }

#define WFA_CUDA(ctx, call)                                                                                          \
    do {                                                                                                             \
        cudaError_t e_ = (call);                                                                                     \
        if (e_ != cudaSuccess) { cudaGetLastError(); return wfa_fail(ctx, HP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } \
    } while (0)

// Graphs (and byte pools) already resident on the device (built by wfa_graph_build_kernel).
struct DevBuilt {
    const WfaJob* jobs = nullptr;
    const WfaNode* nodes = nullptr;
    const uint32_t *child_off = nullptr, *child_idx = nullptr, *amap_off = nullptr, *amap = nullptr;
    const uint8_t *reference = nullptr, *allele_bytes = nullptr, *read_bytes = nullptr, *vtype = nullptr;
    const uint32_t* noise = nullptr;       // probe estimates, one per job of the run
};

struct WfaHostInputs {
    const uint8_t* reference; uint64_t n_reference;
    const uint8_t* allele_bytes; uint64_t n_allele_bytes;
    const uint8_t* seq_pool; uint64_t n_seq_pool;
    const uint8_t* read_bytes; uint64_t n_read_bytes;
    const uint8_t* vtype; uint64_t n_vtype;
};

// Uploads the flattened graphs + byte pools, runs the kernel, downloads the outputs.  `sel` (optional) restricts the
// run to a subset of jobs (retry path); outputs are written at the original job indices.
static int wfa_run(hp_ctx* ctx, const FlatGraphs& fg, const WfaHostInputs& in, uint64_t prune, uint32_t max_ed,
                   uint64_t n_rows, hp_wfa_out* out, uint32_t table_cap, int max_ctas, const DevBuilt* dev = nullptr) {
    const uint32_t nj = (uint32_t)fg.jobs.size();
    if (nj == 0) return HP_OK;
    cudaStream_t st = ctx->stream;
    uint32_t max_nodes = 1;
    for (const WfaJob& j : fg.jobs) if (j.status == HP_WFA_OK) max_nodes = std::max(max_nodes, std::min(j.n_nodes, kWfaMaxNodes));
    const uint32_t sw_max = (max_nodes + 63) / 64;

    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t in_bytes = dev ? al(4ull * nj) + 4096
                          : al(sizeof(WfaJob) * nj) + al(4ull * nj) + al(sizeof(WfaNode) * fg.nodes.size()) + al(4 * fg.child_off.size()) +
                            al(4 * fg.child_idx.size()) + al(4 * fg.amap_off.size()) + al(4 * fg.amap.size()) + al(in.n_reference) +
                            al(in.n_allele_bytes) + al(in.n_seq_pool) + al(in.n_read_bytes) + al(in.n_vtype) + 4096;
    if (!ctx->wfa_in.reserve(in_bytes)) return wfa_fail(ctx, HP_ERR_OUT_OF_MEMORY, "WFA input staging allocation failed");
    uint8_t* p = (uint8_t*)ctx->wfa_in.ptr;
    auto up = [&](const void* src, size_t bytes) -> uint8_t* {
        uint8_t* dst = p; p += al(bytes);
        if (bytes) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
        return dst;
    };
    WfaArgs a;
    a.n_jobs = nj;
    std::vector<uint32_t> order(nj);
    for (uint32_t i = 0; i < nj; i++) order[i] = i;
    // ticket order: jobs the probe marks as noisy first (they run ~10x longer), then longest reads first
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        const bool nx = fg.jobs[x].pad >= kNoisyProbe, ny = fg.jobs[y].pad >= kNoisyProbe;
        if (nx != ny) return nx;
        if (nx) {      // among the noisy: the ones the piece filter is least likely to answer first (they run the longest)
            const uint64_t ex = (uint64_t)fg.jobs[x].pad * fg.jobs[x].read_len, ey = (uint64_t)fg.jobs[y].pad * fg.jobs[y].read_len;
            if (ex != ey) return ex < ey;
        }
        return fg.jobs[x].read_len > fg.jobs[y].read_len;
    });
    if (dev) {
        a.jobs = dev->jobs; a.nodes = dev->nodes; a.child_off = dev->child_off; a.child_idx = dev->child_idx;
        a.amap_off = dev->amap_off; a.amap = dev->amap;
        a.reference = dev->reference; a.allele_bytes = dev->allele_bytes; a.seq_pool = nullptr; a.read_bytes = dev->read_bytes; a.vtype = dev->vtype;
        a.order = (const uint32_t*)up(order.data(), 4ull * nj);
        a.noise = ctx->wfa_no_filter ? nullptr : dev->noise;
    } else {
        a.noise = nullptr;
        a.jobs = (const WfaJob*)up(fg.jobs.data(), sizeof(WfaJob) * nj);
        a.order = (const uint32_t*)up(order.data(), 4ull * nj);
        a.nodes = (const WfaNode*)up(fg.nodes.data(), sizeof(WfaNode) * fg.nodes.size());
        a.child_off = (const uint32_t*)up(fg.child_off.data(), 4 * fg.child_off.size());
        a.child_idx = (const uint32_t*)up(fg.child_idx.data(), 4 * fg.child_idx.size());
        a.amap_off = (const uint32_t*)up(fg.amap_off.data(), 4 * fg.amap_off.size());
        a.amap = (const uint32_t*)up(fg.amap.data(), 4 * fg.amap.size());
        a.reference = up(in.reference, in.n_reference);
        a.allele_bytes = up(in.allele_bytes, in.n_allele_bytes);
        a.seq_pool = up(in.seq_pool, in.n_seq_pool);
        a.read_bytes = up(in.read_bytes, in.n_read_bytes);
        a.vtype = up(in.vtype, in.n_vtype);
    }
    a.prune_distance = prune; a.max_edit_distance = max_ed;

    int n_ctas = std::min<int>((nj + kWfaWarps - 1) / kWfaWarps, ctx->sm_count * HP_WFA_CTAS_PER_SM);
    if (max_ctas > 0) n_ctas = std::min(n_ctas, max_ctas);
    a.table_cap = table_cap; a.set_words_max = sw_max;
    a.slab_bytes = wfa_slab_bytes(table_cap, sw_max);
    const uint64_t n_warps = (uint64_t)n_ctas * kWfaWarps;
    const uint64_t epoch_bytes = (4 * n_warps + 255) & ~255ull;
    if (!ctx->wfa_ws.reserve(a.slab_bytes * n_warps + 256 + epoch_bytes))
        return wfa_fail(ctx, HP_ERR_OUT_OF_MEMORY, "WFA workspace allocation failed");
    a.ticket = (uint32_t*)ctx->wfa_ws.ptr;
    a.epochs = (uint32_t*)((uint8_t*)ctx->wfa_ws.ptr + 256);
    a.slabs = (uint8_t*)ctx->wfa_ws.ptr + 256 + epoch_bytes;
    WFA_CUDA(ctx, cudaMemsetAsync(a.ticket, 0, 256, st));
    // the hash keys are epoch-tagged and never cleared per job: zero them (and the epochs) whenever the tables move
    {
        const uint64_t layout[4] = {(uint64_t)(uintptr_t)ctx->wfa_ws.ptr, a.slab_bytes, (uint64_t)table_cap, n_warps};
        if (memcmp(layout, ctx->wfa_layout, sizeof(layout)) != 0) {
            WFA_CUDA(ctx, cudaMemsetAsync(a.epochs, 0, epoch_bytes, st));
            WFA_CUDA(ctx, cudaMemset2DAsync(a.slabs, a.slab_bytes, 0, (size_t)table_cap * 8, n_warps, st));
            memcpy(ctx->wfa_layout, layout, sizeof(layout));
        }
    }

    const uint32_t tw = out->traversed ? out->trav_words : 0;
    const size_t out_bytes = al(4ull * nj) * 3 + al(n_rows) * 2 + al(8ull * nj * tw) + al(32ull * nj) + 4096;
    if (!ctx->wfa_out.reserve(out_bytes)) return wfa_fail(ctx, HP_ERR_OUT_OF_MEMORY, "WFA output staging allocation failed");
    uint8_t* q = (uint8_t*)ctx->wfa_out.ptr;
    auto carve = [&](size_t bytes) { uint8_t* r = q; q += al(bytes); return r; };
    a.out_status = (int32_t*)carve(4ull * nj); a.out_score = (uint32_t*)carve(4ull * nj);
    a.out_n_nodes = (uint32_t*)carve(4ull * nj);
    a.out_alleles = carve(n_rows); a.out_quals = carve(n_rows);
    a.out_traversed = tw ? (uint64_t*)carve(8ull * nj * tw) : nullptr; a.trav_words = tw;
    a.out_counters = out->counters ? (uint64_t*)carve(32ull * nj) : nullptr;
    a.dbg_times = ctx->wfa_dbg_times ? 1 : 0;

    WFA_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    // node masks: one word per lane up to 1024 nodes, four words (4096 nodes) for batches with a larger graph
    if (max_nodes <= 1024u) wfa_align_kernel<1><<<n_ctas, kWfaWarps * 32, 0, st>>>(a);
    else wfa_align_kernel<4><<<n_ctas, kWfaWarps * 32, 0, st>>>(a);
    WFA_CUDA(ctx, cudaGetLastError());
    WFA_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
    ctx->launches++; ctx->timing_pending = true;

    WFA_CUDA(ctx, cudaMemcpyAsync(&ctx->wfa_filtered_now, a.ticket + 1, 4, cudaMemcpyDeviceToHost, st));
    // download into temporaries, then scatter (the caller's arrays may be indexed by original job ids)
    ctx->wfa_h_status.resize(nj); ctx->wfa_h_score.resize(nj); ctx->wfa_h_nodes.resize(nj);
    ctx->wfa_h_alleles.resize(n_rows); ctx->wfa_h_quals.resize(n_rows);
    ctx->wfa_h_trav.resize((size_t)nj * tw); ctx->wfa_h_ctr.resize(out->counters ? (size_t)nj * 4 : 0);
    WFA_CUDA(ctx, cudaMemcpyAsync(ctx->wfa_h_status.data(), a.out_status, 4ull * nj, cudaMemcpyDeviceToHost, st));
    WFA_CUDA(ctx, cudaMemcpyAsync(ctx->wfa_h_score.data(), a.out_score, 4ull * nj, cudaMemcpyDeviceToHost, st));
    WFA_CUDA(ctx, cudaMemcpyAsync(ctx->wfa_h_nodes.data(), a.out_n_nodes, 4ull * nj, cudaMemcpyDeviceToHost, st));
    if (n_rows) {
        WFA_CUDA(ctx, cudaMemcpyAsync(ctx->wfa_h_alleles.data(), a.out_alleles, n_rows, cudaMemcpyDeviceToHost, st));
        WFA_CUDA(ctx, cudaMemcpyAsync(ctx->wfa_h_quals.data(), a.out_quals, n_rows, cudaMemcpyDeviceToHost, st));
    }
    if (tw) WFA_CUDA(ctx, cudaMemcpyAsync(ctx->wfa_h_trav.data(), a.out_traversed, 8ull * nj * tw, cudaMemcpyDeviceToHost, st));
    if (out->counters) WFA_CUDA(ctx, cudaMemcpyAsync(ctx->wfa_h_ctr.data(), a.out_counters, 32ull * nj, cudaMemcpyDeviceToHost, st));
    WFA_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->wfa_filtered += ctx->wfa_filtered_now;
    return HP_OK;
}

// Builds the graphs of jobs `ids` on the device (two passes: count, then fill exactly sized slices).  On return
// fg.jobs holds the host copy of the job table (node counts, row layout) and `dev` the device pointers.
// The batch-constant inputs (variant table, windows, byte pools) are uploaded once per hp_wfa_align_batch call.
struct DevInputs {
    bool ready = false;
    BuildArgs b{};
    DevBuilt pools;
    const uint64_t* read_off = nullptr;   // device copy of the batch's read offsets (scheduling probe)
    size_t used = 0;               // bytes of ctx->wfa_graph taken by the batch-constant inputs
};

static int wfa_upload_batch(hp_ctx* ctx, const hp_wfa_batch* b, DevInputs& di, size_t graph_bytes_hint) {
    const hp_variant_table& vt = b->variants;
    const uint32_t nj = b->n_jobs, nv = vt.n_variants;
    const uint64_t n_read = b->read_off[nj];
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t bytes = al(8ull * nv) * 3 + al(4ull * nv) * 3 + al(nv) * 3 + al(vt.n_allele_bytes) + al(b->n_reference) + al(n_read) +
                         al(8ull * nj) * 2 + al(4ull * nj) * 4 + al(8ull * (nj + 1)) + 4096;
    if (!ctx->wfa_graph.reserve(bytes + graph_bytes_hint)) return wfa_fail(ctx, HP_ERR_OUT_OF_MEMORY, "WFA graph workspace allocation failed");
    cudaStream_t st = ctx->stream;
    uint8_t* p = (uint8_t*)ctx->wfa_graph.ptr;
    bool ok = true;
    auto up = [&](const void* src, size_t n) -> uint8_t* { uint8_t* d = p; p += al(n); if (n) ok &= cudaMemcpyAsync(d, src, n, cudaMemcpyHostToDevice, st) == cudaSuccess; return d; };
    BuildArgs& a = di.b;
    a.position = (const int64_t*)up(vt.position, 8ull * nv); a.ref_len = (const uint32_t*)up(vt.ref_len, 4ull * nv);
    a.a0_off = (const uint64_t*)up(vt.allele0_off, 8ull * nv); a.a0_len = (const uint32_t*)up(vt.allele0_len, 4ull * nv);
    a.a1_off = (const uint64_t*)up(vt.allele1_off, 8ull * nv); a.a1_len = (const uint32_t*)up(vt.allele1_len, 4ull * nv);
    a.index_allele0 = up(vt.index_allele0, nv); a.ignored = up(vt.ignored, nv);
    di.pools.vtype = up(vt.vtype, nv);
    di.pools.allele_bytes = up(vt.allele_bytes, vt.n_allele_bytes);
    // the two bulk arrays: pageable sources go through pinned staging (hp::upload_large)
    { uint8_t* d = p; p += al(b->n_reference); ok &= upload_large(ctx->pin_ref, d, b->reference, b->n_reference, st); di.pools.reference = d; }
    { uint8_t* d = p; p += al(n_read); ok &= upload_large(ctx->pin_reads, d, b->read_bytes, n_read, st); di.pools.read_bytes = d; }
    a.ref_start = (const uint64_t*)up(b->ref_start, 8ull * nj); a.ref_end = (const uint64_t*)up(b->ref_end, 8ull * nj);
    a.het_lo = (const uint32_t*)up(b->het_lo, 4ull * nj); a.het_hi = (const uint32_t*)up(b->het_hi, 4ull * nj);
    a.hom_lo = (const uint32_t*)up(b->hom_lo, 4ull * nj); a.hom_hi = (const uint32_t*)up(b->hom_hi, 4ull * nj);
    a.n_reference = b->n_reference;
    di.read_off = (const uint64_t*)up(b->read_off, 8ull * (nj + 1));
    if (!ok) { cudaGetLastError(); return wfa_fail(ctx, HP_ERR_CUDA, "WFA batch upload failed"); }
    di.used = (size_t)(p - (uint8_t*)ctx->wfa_graph.ptr);
    di.ready = true;
    return HP_OK;
}

static int wfa_build_on_device(hp_ctx* ctx, const hp_wfa_batch* b, const std::vector<uint32_t>& ids, DevInputs& di,
                               FlatGraphs& fg, DevBuilt& dev, uint64_t& rows) {
    const uint32_t ns = (uint32_t)ids.size();
    cudaStream_t st = ctx->stream;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    // pass 1 needs ids + counts: carve them behind the batch-constant inputs
    size_t need = di.used + al(4ull * ns) + al(16ull * ns) + 4096;
    if (ctx->wfa_graph.cap < need) return wfa_fail(ctx, HP_ERR_INTERNAL, "WFA graph workspace too small for the count pass");
    uint8_t* base = (uint8_t*)ctx->wfa_graph.ptr + di.used;
    uint32_t* d_ids = (uint32_t*)base;
    uint32_t* d_counts = (uint32_t*)(base + al(4ull * ns));
    BuildArgs a = di.b;
    a.n_sel = ns; a.ids = d_ids; a.counts = d_counts;
    WFA_CUDA(ctx, cudaMemcpyAsync(d_ids, ids.data(), 4ull * ns, cudaMemcpyHostToDevice, st));
    const int grid = (int)((ns + 127) / 128);
    wfa_graph_build_kernel<false><<<grid, 128, 0, st>>>(a);
    WFA_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    std::vector<uint32_t> counts(4ull * ns);
    WFA_CUDA(ctx, cudaMemcpyAsync(counts.data(), d_counts, 16ull * ns, cudaMemcpyDeviceToHost, st));
    WFA_CUDA(ctx, cudaStreamSynchronize(st));
    // job table + slice bases
    fg = FlatGraphs();
    fg.jobs.resize(ns);
    std::vector<uint64_t> edge_base(ns), amap_base(ns);
    uint64_t tn = 0, te = 0, ta = 0;
    rows = 0;
    for (uint32_t k = 0; k < ns; k++) {
        const uint32_t j = ids[k];
        WfaJob& job = fg.jobs[k];
        job = WfaJob{};
        job.read_off = b->read_off[j]; job.read_len = (uint32_t)(b->read_off[j + 1] - b->read_off[j]);
        job.row_len = b->het_hi[j] - b->het_lo[j]; job.het_lo = b->het_lo[j];
        job.row_off = rows; rows += job.row_len;
        job.node_base = tn; job.n_nodes = counts[4ull * k];
        job.status = HP_WFA_OK;
        if (job.row_len == 0) job.status = HP_WFA_SKIPPED;                               // read_parsing.rs:703-712
        else if (counts[4ull * k + 3] == kBuildInvalid)
            return wfa_fail(ctx, HP_ERR_INVALID_INPUT, "graph construction failed for job " + std::to_string(j) + " (the reference unwrap()s this, read_parsing.rs:777)");
        else if (counts[4ull * k + 3] == kBuildOverflow) { job.status = HP_WFA_WORKSPACE_OVERFLOW; job.n_nodes = 0; }
        edge_base[k] = te; amap_base[k] = ta;
        if (job.status == HP_WFA_OK) { tn += job.n_nodes; te += counts[4ull * k + 1]; ta += counts[4ull * k + 2]; }
    }
    if (te + ns >= 0xffffffffull || ta + ns >= 0xffffffffull || tn + ns >= 0xffffffffull) return wfa_fail(ctx, HP_ERR_UNSUPPORTED, "WFA batch too large (32-bit CSR offsets)");
    // pass 2 slices
    const size_t fill_bytes = al(sizeof(WfaJob) * ns) + al(8ull * ns) * 2 + al(sizeof(WfaNode) * (tn + 1)) + al(4ull * (tn + ns + 1)) * 2 +
                              al(4ull * (te + 1)) * 3 + al(4ull * (ta + 1));
    need += fill_bytes;
    if (ctx->wfa_graph.cap < need) {
        // grow: the batch-constant inputs have to be uploaded again into the new allocation
        di.ready = false;
        int rc = wfa_upload_batch(ctx, b, di, al(4ull * ns) + al(16ull * ns) + fill_bytes + 8192);
        if (rc != HP_OK) return rc;
        base = (uint8_t*)ctx->wfa_graph.ptr + di.used;
        d_ids = (uint32_t*)base; d_counts = (uint32_t*)(base + al(4ull * ns));
        a = di.b; a.n_sel = ns; a.ids = d_ids; a.counts = d_counts;
        WFA_CUDA(ctx, cudaMemcpyAsync(d_ids, ids.data(), 4ull * ns, cudaMemcpyHostToDevice, st));
        WFA_CUDA(ctx, cudaMemcpyAsync(d_counts, counts.data(), 16ull * ns, cudaMemcpyHostToDevice, st));
    }
    uint8_t* q = base + al(4ull * ns) + al(16ull * ns);
    auto carve = [&](size_t n) { uint8_t* r = q; q += al(n); return r; };
    WfaJob* d_jobs = (WfaJob*)carve(sizeof(WfaJob) * ns);
    uint64_t* d_eb = (uint64_t*)carve(8ull * ns);
    uint64_t* d_ab = (uint64_t*)carve(8ull * ns);
    a.nodes = (WfaNode*)carve(sizeof(WfaNode) * (tn + 1));
    a.child_off = (uint32_t*)carve(4ull * (tn + ns + 1));
    a.amap_off = (uint32_t*)carve(4ull * (tn + ns + 1));
    a.child_idx = (uint32_t*)carve(4ull * (te + 1));
    a.e_parent = (uint32_t*)carve(4ull * (te + 1));
    a.e_child = (uint32_t*)carve(4ull * (te + 1));
    a.amap = (uint32_t*)carve(4ull * (ta + 1));
    a.jobs = d_jobs; a.edge_base = d_eb; a.amap_base = d_ab;
    WFA_CUDA(ctx, cudaMemcpyAsync(d_jobs, fg.jobs.data(), sizeof(WfaJob) * ns, cudaMemcpyHostToDevice, st));
    WFA_CUDA(ctx, cudaMemcpyAsync(d_eb, edge_base.data(), 8ull * ns, cudaMemcpyHostToDevice, st));
    WFA_CUDA(ctx, cudaMemcpyAsync(d_ab, amap_base.data(), 8ull * ns, cudaMemcpyHostToDevice, st));
    wfa_graph_build_kernel<true><<<grid, 128, 0, st>>>(a);
    WFA_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    // scheduling probe (reuses the count-pass slice, which the fill kernel has consumed by stream order)
    std::vector<uint32_t> noise(ns, 0);
    {
        ProbeArgs pa;
        pa.n_sel = ns; pa.ids = d_ids; pa.ref_start = di.b.ref_start; pa.ref_end = di.b.ref_end; pa.read_off = di.read_off;
        pa.reference = di.pools.reference; pa.read_bytes = di.pools.read_bytes; pa.noise = d_counts;
        wfa_noise_probe_kernel<<<grid, 128, 0, st>>>(pa);
        WFA_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
        WFA_CUDA(ctx, cudaMemcpyAsync(noise.data(), d_counts, 4ull * ns, cudaMemcpyDeviceToHost, st));
    }
    // edge_base / amap_base / e_* go out of scope with this function: the fill kernel must have read them
    WFA_CUDA(ctx, cudaStreamSynchronize(st));
    for (uint32_t k = 0; k < ns; k++) fg.jobs[k].pad = noise[k];
    dev = di.pools;
    dev.noise = d_counts;
    dev.jobs = d_jobs; dev.nodes = a.nodes; dev.child_off = a.child_off; dev.child_idx = a.child_idx; dev.amap_off = a.amap_off; dev.amap = a.amap;
    return HP_OK;
}

// scatter results of a run over jobs `ids` (original job indices; rows at the original row offsets)
static void wfa_scatter(hp_ctx* ctx, const FlatGraphs& fg, const std::vector<uint32_t>& ids, const uint64_t* orig_row_off,
                        hp_wfa_out* out) {
    const uint32_t tw = out->traversed ? out->trav_words : 0;
    for (size_t k = 0; k < ids.size(); k++) {
        const uint32_t j = ids[k];
        out->status[j] = ctx->wfa_h_status[k];
        out->score[j] = ctx->wfa_h_score[k];
        if (out->n_nodes) out->n_nodes[j] = ctx->wfa_h_nodes[k];
        const WfaJob& job = fg.jobs[k];
        if (job.row_len) {
            memcpy(out->alleles + orig_row_off[j], ctx->wfa_h_alleles.data() + job.row_off, job.row_len);
            memcpy(out->quals + orig_row_off[j], ctx->wfa_h_quals.data() + job.row_off, job.row_len);
        }
        if (tw) memcpy(out->traversed + (size_t)j * tw, ctx->wfa_h_trav.data() + k * tw, 8ull * tw);
        if (out->counters) memcpy(&out->counters[j], ctx->wfa_h_ctr.data() + k * 4, 32);
    }
}

}  // namespace hp

using namespace hp;

extern "C" {

int hp_wfa_align_batch(hp_ctx* ctx, const hp_wfa_batch* b, hp_wfa_out* out) {
    if (!ctx || !b || !out || !out->status || !out->score || !out->alleles || !out->quals) return HP_ERR_INVALID_INPUT;
    if (b->n_jobs == 0) return HP_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return wfa_fail(ctx, HP_ERR_CUDA, "cudaSetDevice failed");
    ctx->wfa_filtered = 0;
    const hp_variant_table& vt = b->variants;
    for (uint32_t j = 0; j < b->n_jobs; j++) {
        if (b->het_hi[j] < b->het_lo[j] || b->hom_hi[j] < b->hom_lo[j] || b->het_hi[j] > vt.n_variants || b->hom_hi[j] > vt.n_variants ||
            b->read_off[j + 1] < b->read_off[j] || b->row_off[j + 1] - b->row_off[j] != (uint64_t)(b->het_hi[j] - b->het_lo[j]))
            return wfa_fail(ctx, HP_ERR_INVALID_INPUT, "malformed WFA job " + std::to_string(j));
    }
    const uint64_t prune = ctx->params.wfa_prune_distance == 0 ? UINT64_MAX : ctx->params.wfa_prune_distance;   // cli.rs:352-354
    WfaHostInputs in{b->reference, b->n_reference, vt.allele_bytes, vt.n_allele_bytes, nullptr, 0, b->read_bytes,
                     b->read_off[b->n_jobs], vt.vtype, vt.n_variants};

    std::vector<uint32_t> ids(b->n_jobs);
    for (uint32_t j = 0; j < b->n_jobs; j++) ids[j] = j;
    uint32_t cap = ctx->wfa_table_cap;
    // position-sorted het / hom runs are part of the contract (the device builder merges them); the host builder sorts
    for (uint32_t j = 0; j < b->n_jobs && !ctx->wfa_host_build; j++) {
        for (uint32_t k = b->het_lo[j]; k + 1 < b->het_hi[j]; k++) if (vt.position[k] > vt.position[k + 1]) return wfa_fail(ctx, HP_ERR_INVALID_INPUT, "het variants of job " + std::to_string(j) + " are not sorted by position");
        for (uint32_t k = b->hom_lo[j]; k + 1 < b->hom_hi[j]; k++) if (vt.position[k] > vt.position[k + 1]) return wfa_fail(ctx, HP_ERR_INVALID_INPUT, "hom variants of job " + std::to_string(j) + " are not sorted by position");
        if (b->ref_end[j] > b->n_reference || b->ref_start[j] > b->ref_end[j]) return wfa_fail(ctx, HP_ERR_INVALID_INPUT, "reference window of job " + std::to_string(j));
    }
    DevInputs di;
    for (int attempt = 0; attempt < 4 && !ids.empty(); attempt++) {
      FlatGraphs fg;
      uint64_t rows = 0;
      DevBuilt dev;
      // retries (wave table overflow, or a graph the device builder's fixed-size lists cannot hold) build on the host
      const bool on_device = !ctx->wfa_host_build && attempt == 0;
      if (on_device) {
        if (!di.ready) {
            // workspace hint: ~3 nodes per variant in a window (exact sizes come from the count pass)
            uint64_t sv = 0;
            for (uint32_t j : ids) sv += (uint64_t)(b->het_hi[j] - b->het_lo[j]) + (b->hom_hi[j] - b->hom_lo[j]);
            const size_t hint = ctx->wfa_no_hint ? 0 : (size_t)((3 * sv + 2 * ids.size()) * (16 + 8 + 18 + 8) + 64 * ids.size() + (1u << 20));
            int rc0 = wfa_upload_batch(ctx, b, di, hint);
            if (rc0 != HP_OK) return rc0;
        }
        int rc0 = wfa_build_on_device(ctx, b, ids, di, fg, dev, rows);
        if (rc0 != HP_OK) return rc0;
      } else {
        // graph construction is per job and independent: build chunks on host threads, then concatenate
        const size_t n_ids = ids.size();
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const size_t n_chunks = std::max<size_t>(1, std::min<size_t>(hw, n_ids / 64));
        std::vector<FlatGraphs> parts(n_chunks);
        std::vector<int> part_fail(n_chunks, -1);
        auto build_chunk = [&](size_t c) {
            const size_t lo = n_ids * c / n_chunks, hi = n_ids * (c + 1) / n_chunks;
            FlatGraphs& pg = parts[c];
            pg.jobs.reserve(hi - lo);
            for (size_t q = lo; q < hi; q++) {
                const uint32_t j = ids[q];
                WfaJob job{};
                job.read_off = b->read_off[j]; job.read_len = (uint32_t)(b->read_off[j + 1] - b->read_off[j]);
                job.row_len = b->het_hi[j] - b->het_lo[j]; job.het_lo = b->het_lo[j];
                job.status = HP_WFA_OK;
                if (job.row_len == 0) { job.status = HP_WFA_SKIPPED; job.node_base = pg.nodes.size(); job.n_nodes = 0; }   // read_parsing.rs:703-712
                else if (!build_job_graph(b, j, pg, job)) { part_fail[c] = (int)j; return; }
                if (job.status != HP_WFA_OK) {            // keep the CSR row layout: one (empty) row per job
                    pg.child_off.push_back((uint32_t)pg.child_idx.size());
                    pg.amap_off.push_back((uint32_t)pg.amap.size());
                }
                pg.jobs.push_back(job);
            }
        };
        if (n_chunks == 1) build_chunk(0);
        else {
            std::vector<std::thread> pool;
            for (size_t c = 0; c < n_chunks; c++) pool.emplace_back(build_chunk, c);
            for (auto& t : pool) t.join();
        }
        for (size_t c = 0; c < n_chunks; c++)
            if (part_fail[c] >= 0)
                return wfa_fail(ctx, HP_ERR_INVALID_INPUT, "graph construction failed for job " + std::to_string(part_fail[c]) + " (the reference unwrap()s this, read_parsing.rs:777)");
        {
            size_t tn = 0, tc = 0, ta = 0, tj = 0, to = 0;
            for (const FlatGraphs& pg : parts) { tn += pg.nodes.size(); tc += pg.child_idx.size(); ta += pg.amap.size(); tj += pg.jobs.size(); to += pg.child_off.size(); }
            fg.nodes.reserve(tn); fg.child_idx.reserve(tc); fg.amap.reserve(ta); fg.jobs.reserve(tj); fg.child_off.reserve(to); fg.amap_off.reserve(to);
            for (FlatGraphs& pg : parts) {
                const uint64_t nb0 = fg.nodes.size();
                const uint32_t cb0 = (uint32_t)fg.child_idx.size(), ab0 = (uint32_t)fg.amap.size();
                for (WfaJob job : pg.jobs) { job.node_base += nb0; job.row_off = rows; rows += job.row_len; fg.jobs.push_back(job); }
                fg.nodes.insert(fg.nodes.end(), pg.nodes.begin(), pg.nodes.end());
                fg.child_idx.insert(fg.child_idx.end(), pg.child_idx.begin(), pg.child_idx.end());
                fg.amap.insert(fg.amap.end(), pg.amap.begin(), pg.amap.end());
                for (uint32_t x : pg.child_off) fg.child_off.push_back(x + cb0);
                for (uint32_t x : pg.amap_off) fg.amap_off.push_back(x + ab0);
                pg = FlatGraphs();
            }
        }
      }
        // slabs: full occupancy on the first attempt, fewer and larger afterwards
        int max_ctas = 0;
        if (attempt > 0) max_ctas = std::max(1, (int)((8ull << 30) / (wfa_slab_bytes(cap, 16) * kWfaWarps)));
        int rc = wfa_run(ctx, fg, in, prune, ctx->params.wfa_max_edit_distance, rows, out, cap, max_ctas, on_device ? &dev : nullptr);
        if (rc != HP_OK) return rc;
        wfa_scatter(ctx, fg, ids, b->row_off, out);
        std::vector<uint32_t> redo;
        for (uint32_t j : ids) if (out->status[j] == HP_WFA_WORKSPACE_OVERFLOW) redo.push_back(j);
        ids.swap(redo);
        cap = std::min<uint32_t>(cap * 8, 1u << 24);
    }
    return HP_OK;
}

int hp_wfa_graph_align(hp_ctx* ctx, uint32_t n_nodes, const uint8_t* seq, const uint64_t* seq_off, const uint32_t* parent_idx,
                       const uint64_t* parent_off, const uint8_t* read, uint64_t read_len, uint64_t prune_distance,
                       uint32_t max_edit_distance, int32_t* status, uint32_t* score, uint64_t* traversed) {
    if (!ctx || !seq_off || !parent_off || !status || !score) return HP_ERR_INVALID_INPUT;
    if (n_nodes == 0) return wfa_fail(ctx, HP_ERR_INVALID_INPUT, "empty graph");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return wfa_fail(ctx, HP_ERR_CUDA, "cudaSetDevice failed");
    // add_node rules (wfa_graph.rs:298-331): root has no parents, every other node has >= 1, all parents precede
    for (uint32_t i = 0; i < n_nodes; i++) {
        const uint64_t np = parent_off[i + 1] - parent_off[i];
        if ((i == 0) != (np == 0)) return wfa_fail(ctx, HP_ERR_INVALID_INPUT, "node parent rule violated (wfa_graph.rs:302-311)");
        for (uint64_t q = parent_off[i]; q < parent_off[i + 1]; q++)
            if (parent_idx[q] >= i) return wfa_fail(ctx, HP_ERR_INVALID_INPUT, "parents must precede their node (wfa_graph.rs:313-317)");
    }
    FlatGraphs fg;
    WfaJob job{};
    job.node_base = 0; job.n_nodes = n_nodes; job.read_off = 0; job.read_len = (uint32_t)read_len; job.row_off = 0; job.row_len = 0;
    job.status = HP_WFA_OK;
    std::vector<uint32_t> deg(n_nodes + 1, 0);
    for (uint32_t i = 0; i < n_nodes; i++) for (uint64_t q = parent_off[i]; q < parent_off[i + 1]; q++) deg[parent_idx[q] + 1]++;
    for (uint32_t i = 0; i < n_nodes; i++) deg[i + 1] += deg[i];
    fg.child_idx.resize(deg[n_nodes]);
    std::vector<uint32_t> fill(deg.begin(), deg.end() - 1);
    for (uint32_t i = 0; i < n_nodes; i++) for (uint64_t q = parent_off[i]; q < parent_off[i + 1]; q++) fg.child_idx[fill[parent_idx[q]]++] = i;
    for (uint32_t i = 0; i <= n_nodes; i++) { fg.child_off.push_back(deg[i]); fg.amap_off.push_back(0); }
    for (uint32_t i = 0; i < n_nodes; i++) fg.nodes.push_back(WfaNode{seq_off[i], (uint32_t)(seq_off[i + 1] - seq_off[i]), 2u});
    fg.jobs.push_back(job);
    WfaHostInputs in{nullptr, 0, nullptr, 0, seq, seq_off[n_nodes], read, read_len, nullptr, 0};
    const uint32_t tw = (n_nodes + 63) / 64;
    hp_wfa_out o{};
    int32_t st = -1; uint32_t sc = 0;
    std::vector<uint64_t> trav(tw, 0);
    uint8_t dummy = 0;
    o.status = &st; o.score = &sc; o.alleles = &dummy; o.quals = &dummy; o.traversed = trav.data(); o.trav_words = tw;
    uint32_t cap = ctx->wfa_table_cap;
    std::vector<uint32_t> ids{0};
    uint64_t row0[2] = {0, 0};
    for (int attempt = 0; attempt < 4; attempt++) {
        int rc = wfa_run(ctx, fg, in, prune_distance, max_edit_distance, 0, &o, cap, attempt ? 1 : 0);
        if (rc != HP_OK) return rc;
        wfa_scatter(ctx, fg, ids, row0, &o);
        if (st != HP_WFA_WORKSPACE_OVERFLOW) break;
        cap = std::min<uint32_t>(cap * 8, 1u << 24);
    }
    if (st == HP_WFA_WORKSPACE_OVERFLOW) return wfa_fail(ctx, HP_ERR_UNSUPPORTED, "wave table overflow after four workspace enlargements");
    if (st == HP_WFA_GRAPH_TOO_LARGE) return wfa_fail(ctx, HP_ERR_UNSUPPORTED, "graph has more than 4096 nodes (outside the kernel's range)");
    *status = st; *score = sc;
    if (traversed) memcpy(traversed, trav.data(), 8ull * tw);
    return HP_OK;
}

}  // extern "C"
