// hp_device.cuh -- shared device-side definitions for the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define HP_FULL_MASK 0xffffffffu
#define HP_MAX_SEGMENT 40u            // heuristic look-ahead, src/astar_phaser.rs:466
#define HP_PLANE_STRIDE 10u           // words per read-word record: allele bit, non-binary bit, 8 quality bit-planes

namespace hp {

// Per-block metadata produced by astar_prep_kernel.
struct BlkMeta {
    uint64_t var_base;     // first variant of the block in the batch-wide variant arrays
    uint64_t read_base;    // first read
    uint64_t cell_base;    // first cell
    uint32_t n_var;
    uint32_t n_reads;
    uint32_t n_cells;
    uint32_t qgcd;         // gcd of all quality values of the block (>= 1)
    uint32_t n_planes;     // bit planes needed for qual / qgcd
    int32_t  status;       // HP_BLOCK_*
    uint32_t max_span;     // longest read region
    uint32_t max_act;      // largest number of reads covering one variant
};

// Per-read metadata (16 B, one LDG.128).
struct __align__(16) ReadMeta {
    uint32_t start;        // region start (block-relative variant index)
    uint32_t end;          // region end (exclusive)
    uint32_t word_idx;     // index of the read's first 64-variant word record in the plane array
    uint32_t cell_rel;     // offset of the read's first cell relative to the block's cell_base
};

// Everything the solver kernel needs, passed by value.
struct AstarArgs {
    uint32_t n_blocks;
    // batch (reference u8 layout, device pointers)
    const uint8_t* alleles;
    const uint8_t* quals;
    const uint8_t* ignored;
    const uint8_t* is_snv;
    // prep products
    const BlkMeta*  meta;        // [n_blocks]
    const ReadMeta* rmeta;       // [n_reads]
    const uint64_t* planes;      // [n_words * HP_PLANE_STRIDE]
    const uint32_t* act_off;     // [n_vars + n_blocks] per block N+1 offsets (relative to cell_base) into act_idx
    const uint32_t* act_idx;     // [n_cells] block-relative read index of each (variant, covering read) pair
    const uint32_t* col;         // [n_cells] column record of the same pair: qual | allele<<8 | ends<<10 | carry<<16
    const uint32_t* order;       // [n_blocks] processing order: grouped by score-vector class, largest first inside a class
    const uint32_t* class_info;  // [8]: class_count[0..2], pad, class_start[0..2], pad (class 0: K=1, 1: K=2, 2: K=0)
    // scratch
    uint32_t* heur;              // [n_vars + n_blocks] H[] per block (u32 is exact: total quals < 2^31 is enforced)
    uint32_t* ticket;            // work-queue counters, one per class
    uint8_t*  slabs;             // pool of main-queue slabs; a CTA claims one when it gets its first block
    uint64_t  slab_bytes;
    uint32_t* slab_busy;         // [n_slabs] 0 = free, 1 = claimed (the pool is shared by every launch of the context)
    uint32_t  n_slabs;
    uint32_t  slab_seed;         // where this launch starts probing (spreads concurrent launches over the pool)
    uint32_t  qcap;              // main-queue capacity per warp (entries, multiple of 32)
    uint32_t  hap_words;         // 64-bit words per haplotype in a main-queue record (covers the largest block)
    // parameters
    uint32_t min_queue_size;
    uint32_t queue_increment;
    uint32_t sub_capl;           // sub-solver stripe capacity (entries per lane)
    // outputs (device pointers)
    uint8_t*  out_h1;
    uint8_t*  out_h2;
    uint64_t* out_stats;         // [n_blocks * 7]
    int32_t*  out_status;        // [n_blocks]
    uint64_t* out_heur;          // optional [n_vars + n_blocks]
    uint64_t* out_counters;      // optional [n_blocks * 4]
    uint64_t* dbg_cycles;        // optional [n_blocks * 16]: per-block phase cycles / pops / rounds / team, main-loop split (counting variant)
};

// Arguments of astar_prep_kernel.
struct PrepArgs {
    uint32_t n_blocks;
    const uint64_t* var_off;
    const uint64_t* read_off;
    const uint32_t* read_start;
    const uint32_t* read_end;
    const uint64_t* cell_off;
    const uint8_t* alleles;
    const uint8_t* quals;
    const uint8_t* ignored;
    BlkMeta* meta;
    ReadMeta* rmeta;
    uint64_t* planes;
    uint32_t* act_off;   // zero-initialised; first used as per-variant counters, then overwritten with offsets
    uint32_t* act_cur;   // zero-initialised cursors
    uint32_t* act_idx;
    uint32_t* col;
    uint64_t n_cells_total;   // length of alleles / quals (bulk copies never read past it)
    uint32_t smem_cap;        // blocks with at most this many variants keep their coverage counters in shared memory
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// 64-bit mask with bits [lo, hi) set; lo, hi in [0, 64], lo <= hi.
__device__ __forceinline__ uint64_t bit_range(int lo, int hi) {
    uint64_t m_hi = (hi >= 64) ? ~0ull : ((1ull << hi) - 1ull);
    uint64_t m_lo = (lo >= 64) ? ~0ull : ((1ull << lo) - 1ull);
    return m_hi & ~m_lo;
}

// shift with sign: s >= 0 -> x >> s, s < 0 -> x << -s; out-of-range shifts give 0.
__device__ __forceinline__ uint64_t shift_signed(uint64_t x, int s) {
    if (s >= 64 || s <= -64) return 0ull;
    return (s >= 0) ? (x >> s) : (x << (-s));
}

}  // namespace hp
