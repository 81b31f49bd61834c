// assemble_kernels.cu -- matrix assembly on sm_100a: per-mapping rows -> collapsed ReadSegments (row a6 of SURVEY.md 8a).
// Replaces ReadSegment::new (src/data_types/read_segments.rs:40-62), ReadSegment::collapse (:71-121), get_num_set
// (:151-155) and the min-matched-alleles filter (src/read_parsing.rs:612-629).
//
// One warp per read group.  Pass 1: the set-region of every row ([first 0/1 cell, last 0/1 cell + 1): cells outside it read
// as NoOverlap, read_segments.rs:128-143).  Pass 2: lanes over the positions of the group's span, rows applied in push
// order with the reference's rule (first non-NoOverlap wins; equal alleles keep the larger quality; a conflict makes the
// cell Ambiguous with quality 0).  The collapsed cells go to a scratch slice; the host keeps the groups with enough set
// alleles and packs their clipped cells into the block batch.
#include <algorithm>
#include <string>
#include <vector>

#include "hp_host.h"

namespace hp {

struct AssembleArgs {
    uint32_t n_groups;
    const uint64_t* group_row_off;
    const uint32_t* row_start;
    const uint64_t* row_cell_off;
    const uint8_t* alleles;
    const uint8_t* quals;
    const uint64_t* scratch_off;     // [n_groups] slice of the collapsed cells
    const uint32_t* span_lo;         // [n_groups] first block-relative index any row of the group touches
    uint32_t* row_first;             // [n_rows] scratch: set-region of each row (block-relative)
    uint32_t* row_last;
    uint8_t* out_alleles;            // scratch slices
    uint8_t* out_quals;
    uint32_t* g_first;               // [n_groups] region of the collapsed segment (block-relative), g_first >= g_last: empty
    uint32_t* g_last;
    uint32_t* g_num_set;
    uint32_t* g_flags;               // 1 = the reference's assert (:108) would fire
};

constexpr int kAsmWarps = 4;

__global__ void __launch_bounds__(kAsmWarps * 32) assemble_kernel(AssembleArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = blockIdx.x * kAsmWarps + (threadIdx.x >> 5), nw = gridDim.x * kAsmWarps;
    for (uint32_t g = gw; g < a.n_groups; g += nw) {
        const uint64_t r0 = a.group_row_off[g], r1 = a.group_row_off[g + 1];
        // ---- pass 1: ReadSegment::new per row (read_segments.rs:40-62) ----
        uint32_t gmin = 0xffffffffu, gmax = 0;
        for (uint64_t r = r0; r < r1; r++) {
            const uint64_t c0 = a.row_cell_off[r];
            const uint32_t n = (uint32_t)(a.row_cell_off[r + 1] - c0), rs = a.row_start[r];
            uint32_t first = 0xffffffffu, last = 0;
            for (uint32_t i = lane; i < n; i += 32)
                if (a.alleles[c0 + i] < HP_ALLELE_AMBIGUOUS) { first = min(first, i); last = max(last, i + 1); }
            first = __reduce_min_sync(HP_FULL_MASK, first); last = __reduce_max_sync(HP_FULL_MASK, last);
            const bool any = last != 0;
            if (lane == 0) { a.row_first[r] = any ? rs + first : 0u; a.row_last[r] = any ? rs + last : 0u; }
            if (any) { gmin = min(gmin, rs + first); gmax = max(gmax, rs + last); }
        }
        __syncwarp();
        // ---- pass 2: ReadSegment::collapse (:71-121) over [gmin, gmax), then ReadSegment::new of the result ----
        const uint64_t so = a.scratch_off[g];
        const uint32_t lo = a.span_lo[g];
        uint32_t cfirst = 0xffffffffu, clast = 0, nset = 0, flags = 0;
        for (uint32_t pos = (gmin == 0xffffffffu ? 0u : gmin) + lane; pos < gmax; pos += 32) {
            uint8_t al = HP_ALLELE_NOOVERLAP, ql = 0;
            if (r1 - r0 == 1) {                                                       // short circuit (:74-76): the row itself
                const uint64_t c = a.row_cell_off[r0] + (pos - a.row_start[r0]);
                al = a.alleles[c]; ql = a.quals[c];
            } else
            for (uint64_t r = r0; r < r1; r++) {
                if (pos < a.row_first[r] || pos >= a.row_last[r]) continue;         // ReadSegment::allele outside the region
                const uint64_t c = a.row_cell_off[r] + (pos - a.row_start[r]);
                const uint8_t rsa = a.alleles[c], rsq = a.quals[c];
                if (rsa == HP_ALLELE_NOOVERLAP) continue;
                if (al == HP_ALLELE_NOOVERLAP) { al = rsa; ql = rsq; }
                else if (al == HP_ALLELE_AMBIGUOUS) { }
                else if (al == rsa) { ql = max(ql, rsq); if (ql == 0) flags = 1u; }   // assert!(quals[i] > 0), :108
                else { al = HP_ALLELE_AMBIGUOUS; ql = 0; }
            }
            a.out_alleles[so + (pos - lo)] = al; a.out_quals[so + (pos - lo)] = ql;
            if (al < HP_ALLELE_AMBIGUOUS) { cfirst = min(cfirst, pos); clast = max(clast, pos + 1); nset++; }
        }
        cfirst = __reduce_min_sync(HP_FULL_MASK, cfirst); clast = __reduce_max_sync(HP_FULL_MASK, clast);
        nset = __reduce_add_sync(HP_FULL_MASK, nset); flags = __reduce_or_sync(HP_FULL_MASK, flags);
        if (lane == 0) { a.g_first[g] = clast ? cfirst : 0u; a.g_last[g] = clast; a.g_num_set[g] = nset; a.g_flags[g] = flags; }
        __syncwarp();
    }
}

}  // namespace hp

using namespace hp;

extern "C" int hp_assemble_blocks(hp_ctx* ctx, const hp_rows_batch* b, hp_assembled* out) {
    if (!ctx || !b || !out || !out->read_off || !out->read_start || !out->read_end || !out->cell_off || !out->alleles || !out->quals)
        return HP_ERR_INVALID_INPUT;
    auto fail = [&](int code, const std::string& msg) { ctx->err = msg; return code; };
    const uint32_t nb = b->n_blocks;
    out->n_reads = 0; out->n_cells = 0;
    out->read_off[0] = 0; out->cell_off[0] = 0;
    if (nb == 0) return HP_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(HP_ERR_CUDA, "cudaSetDevice failed");
    const uint64_t ng = b->group_off[nb], nr = ng ? b->group_row_off[ng] : 0, nc = nr ? b->row_cell_off[nr] : 0;
    if (ng >= 0xffffffffull) return fail(HP_ERR_UNSUPPORTED, "too many read groups");
    // ---- validation + scratch layout: the span every group can touch ----
    std::vector<uint64_t> scratch_off(ng + 1, 0);
    std::vector<uint32_t> span_lo(ng, 0);
    for (uint32_t blk = 0; blk < nb; blk++) {
        const uint64_t N = b->var_off[blk + 1] - b->var_off[blk];
        if (b->group_off[blk + 1] < b->group_off[blk]) return fail(HP_ERR_INVALID_INPUT, "group_off must be non-decreasing");
        for (uint64_t g = b->group_off[blk]; g < b->group_off[blk + 1]; g++) {
            if (b->group_row_off[g + 1] < b->group_row_off[g]) return fail(HP_ERR_INVALID_INPUT, "group_row_off must be non-decreasing");
            uint64_t lo = UINT64_MAX, hi = 0;
            for (uint64_t r = b->group_row_off[g]; r < b->group_row_off[g + 1]; r++) {
                if (b->row_cell_off[r + 1] < b->row_cell_off[r]) return fail(HP_ERR_INVALID_INPUT, "row_cell_off must be non-decreasing");
                const uint64_t len = b->row_cell_off[r + 1] - b->row_cell_off[r];
                if ((uint64_t)b->row_start[r] + len > N) return fail(HP_ERR_INVALID_INPUT, "row " + std::to_string(r) + " leaves its block");
                if (len) { lo = std::min<uint64_t>(lo, b->row_start[r]); hi = std::max<uint64_t>(hi, b->row_start[r] + len); }
            }
            span_lo[g] = lo == UINT64_MAX ? 0u : (uint32_t)lo;
            scratch_off[g + 1] = scratch_off[g] + (lo == UINT64_MAX ? 0 : hi - lo);
        }
    }
    const uint64_t n_scratch = scratch_off[ng];
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t in_bytes = al(8 * (ng + 1)) * 2 + al(4 * ng) + al(4 * nr) + al(8 * (nr + 1)) + al(nc) * 2 + 4096;
    const size_t out_bytes = al(4 * nr) * 2 + al(n_scratch) * 2 + al(4 * ng) * 4 + 4096;
    if (!ctx->stage_in.reserve(in_bytes) || !ctx->stage_out.reserve(out_bytes)) return fail(HP_ERR_OUT_OF_MEMORY, "staging allocation failed");
    cudaStream_t st = ctx->stream;
    uint8_t* p = (uint8_t*)ctx->stage_in.ptr;
    bool ok = true;
    auto up = [&](const void* src, size_t n) { uint8_t* d = p; p += al(n); if (n) ok &= cudaMemcpyAsync(d, src, n, cudaMemcpyHostToDevice, st) == cudaSuccess; return d; };
    AssembleArgs a;
    a.n_groups = (uint32_t)ng;
    a.group_row_off = (const uint64_t*)up(b->group_row_off, 8 * (ng + 1));
    a.scratch_off = (const uint64_t*)up(scratch_off.data(), 8 * (ng + 1));
    a.span_lo = (const uint32_t*)up(span_lo.data(), 4 * ng);
    a.row_start = (const uint32_t*)up(b->row_start, 4 * nr);
    a.row_cell_off = (const uint64_t*)up(b->row_cell_off, 8 * (nr + 1));
    a.alleles = up(b->alleles, nc); a.quals = up(b->quals, nc);
    uint8_t* q = (uint8_t*)ctx->stage_out.ptr;
    auto carve = [&](size_t n) { uint8_t* d = q; q += al(n); return d; };
    a.row_first = (uint32_t*)carve(4 * nr); a.row_last = (uint32_t*)carve(4 * nr);
    a.out_alleles = carve(n_scratch); a.out_quals = carve(n_scratch);
    a.g_first = (uint32_t*)carve(4 * ng); a.g_last = (uint32_t*)carve(4 * ng);
    a.g_num_set = (uint32_t*)carve(4 * ng); a.g_flags = (uint32_t*)carve(4 * ng);
    std::vector<uint8_t> h_al(n_scratch), h_ql(n_scratch);
    std::vector<uint32_t> h_first(ng), h_last(ng), h_nset(ng), h_flags(ng);
    if (ng) {
        const int grid = (int)std::min<uint64_t>((ng + kAsmWarps - 1) / kAsmWarps, (uint64_t)ctx->sm_count * 8);
        assemble_kernel<<<grid, kAsmWarps * 32, 0, st>>>(a);
        ok &= cudaGetLastError() == cudaSuccess;
        ctx->launches++;
        if (n_scratch) {
            ok &= cudaMemcpyAsync(h_al.data(), a.out_alleles, n_scratch, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            ok &= cudaMemcpyAsync(h_ql.data(), a.out_quals, n_scratch, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        }
        ok &= cudaMemcpyAsync(h_first.data(), a.g_first, 4 * ng, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok &= cudaMemcpyAsync(h_last.data(), a.g_last, 4 * ng, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok &= cudaMemcpyAsync(h_nset.data(), a.g_num_set, 4 * ng, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok &= cudaMemcpyAsync(h_flags.data(), a.g_flags, 4 * ng, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok &= cudaStreamSynchronize(st) == cudaSuccess;
    }
    if (!ok) { cudaGetLastError(); return fail(HP_ERR_CUDA, "matrix assembly launch or copy failed"); }
    // ---- the filter of read_parsing.rs:612-629 and the packing into the block batch ----
    uint64_t n_reads = 0, n_cells = 0;
    for (uint32_t blk = 0; blk < nb; blk++) {
        for (uint64_t g = b->group_off[blk]; g < b->group_off[blk + 1]; g++) {
            uint8_t cls = HP_GROUP_DROPPED;
            if (h_flags[g]) cls = HP_GROUP_ASSERT;
            else if (h_nset[g] >= b->min_matched_alleles && h_nset[g] > 0) cls = HP_GROUP_KEPT;
            else if (h_nset[g] > 0) cls = HP_GROUP_PHASABLE;
            if (out->group_class) out->group_class[g] = cls;
            if (out->group_num_set) out->group_num_set[g] = cls == HP_GROUP_ASSERT ? 0u : h_nset[g];
            if (cls != HP_GROUP_KEPT) continue;
            const uint32_t s0 = h_first[g], e0 = h_last[g];
            if (n_cells + (e0 - s0) > out->cell_capacity) return fail(HP_ERR_INVALID_INPUT, "cell_capacity too small");
            out->read_start[n_reads] = s0; out->read_end[n_reads] = e0;
            const uint64_t src = scratch_off[g] + (s0 - span_lo[g]);
            std::copy(h_al.begin() + src, h_al.begin() + src + (e0 - s0), out->alleles + n_cells);
            std::copy(h_ql.begin() + src, h_ql.begin() + src + (e0 - s0), out->quals + n_cells);
            n_cells += e0 - s0;
            n_reads++;
            out->cell_off[n_reads] = n_cells;
        }
        out->read_off[blk + 1] = n_reads;
    }
    out->n_reads = n_reads; out->n_cells = n_cells;
    return HP_OK;
}
