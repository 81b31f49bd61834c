// post_kernels.cu -- post-solve step on sm_100a: span counts per juncture, sub-block tags and
// read haplotagging.  Replaces get_solution_span_counts (src/phaser.rs:350-388), the block_split / block_tags loop
// (src/phaser.rs:546-569) and haplotag_reads (src/phaser.rs:714-750; score_haplotype read_segments.rs:161-168).
//
// One warp per phase block (grid-stride over blocks): lanes over reads for the juncture ranges (+1/-1 difference
// array, then a warp prefix scan) and for the two haplotype scores; lanes over variants for the tag scan.
#include <string>

#include "hp_host.h"

namespace hp {

struct PostArgs {
    uint32_t n_blocks;
    const uint64_t* var_off;
    const uint64_t* read_off;
    const uint32_t* read_start;
    const uint32_t* read_end;
    const uint64_t* cell_off;
    const uint8_t* alleles;
    const uint8_t* quals;
    const int64_t* var_pos;
    const uint8_t* h1;
    const uint8_t* h2;
    uint32_t* span;        // also the difference-array scratch
    uint64_t* tags;
    uint8_t* read_hap;
    uint64_t* read_tag;
};

constexpr int kPostWarps = 4;

__global__ void __launch_bounds__(kPostWarps * 32) post_solve_kernel(PostArgs a) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = blockIdx.x * kPostWarps + (threadIdx.x >> 5), nw = gridDim.x * kPostWarps;
    for (uint32_t blk = gw; blk < a.n_blocks; blk += nw) {
        const uint64_t v0 = a.var_off[blk];
        const uint32_t N = (uint32_t)(a.var_off[blk + 1] - v0);
        const uint64_t r0 = a.read_off[blk], r1 = a.read_off[blk + 1];
        const uint8_t* h1 = a.h1 + v0;
        const uint8_t* h2 = a.h2 + v0;
        int* diff = (int*)(a.span + v0);
        for (uint32_t i = lane; i < N; i += 32) diff[i] = 0;
        __syncwarp();
        // ---- juncture range of every read (phaser.rs:362-385) ----
        for (uint64_t r = r0 + lane; r < r1; r += 32) {
            uint32_t js = a.read_start[r], je = a.read_end[r];
            if (je == 0) continue;
            je -= 1;
            while (js < je && h1[js] == h2[js]) js++;
            while (js < je && h1[je] == h2[je]) je--;
            if (js < je) { atomicAdd(&diff[js], 1); atomicAdd(&diff[je], -1); }
        }
        __syncwarp();
        // ---- prefix scan -> span counts; running "last split" -> tags (phaser.rs:546-569) ----
        int carry = 0;
        uint32_t split_carry = 0;                      // index of the variant that starts the current sub-block
        for (uint32_t base = 0; base < N; base += 32) {
            const uint32_t i = base + lane;
            int x = (i < N) ? diff[i] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(HP_FULL_MASK, x, o); if ((int)lane >= o) x += y; }
            x += carry;                                // x = span count of juncture i (between variants i and i+1)
            // variant i starts a new sub-block if i == 0 or span[i-1] == 0
            int prev = __shfl_up_sync(HP_FULL_MASK, x, 1);
            if (lane == 0) prev = carry;
            uint32_t st = (i < N && (i == 0 || prev == 0)) ? i : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(HP_FULL_MASK, st, o); if ((int)lane >= o) st = max(st, y); }
            st = max(st, split_carry);
            if (i < N) {
                a.tags[v0 + i] = (uint64_t)a.var_pos[v0 + st];
                a.span[v0 + i] = (i + 1 < N) ? (uint32_t)x : 0u;
            }
            carry = __shfl_sync(HP_FULL_MASK, x, 31);
            split_carry = __shfl_sync(HP_FULL_MASK, st, 31);
        }
        __syncwarp();
        // ---- haplotag_reads (phaser.rs:714-750) ----
        for (uint64_t r = r0 + lane; r < r1; r += 32) {
            const uint32_t s = a.read_start[r], e = a.read_end[r];
            const uint8_t* al = a.alleles + a.cell_off[r];
            const uint8_t* ql = a.quals + a.cell_off[r];
            uint64_t s1 = 0, s2 = 0;
            for (uint32_t i = s; i < e; i++) {
                const uint8_t x = al[i - s], q = ql[i - s];
                if (h1[i] < 2 && x != h1[i]) s1 += q;
                if (h2[i] < 2 && x != h2[i]) s2 += q;
            }
            const uint8_t tagv = s1 < s2 ? 0 : (s1 > s2 ? 1 : 2);
            uint64_t rtag = 0;
            if (tagv != 2) {
                uint32_t fv = s;
                while (fv < e && (h1[fv] == h2[fv] || al[fv - s] >= 2)) fv++;
                rtag = (fv < e) ? a.tags[v0 + fv] : 0;
            }
            a.read_hap[r] = tagv;
            a.read_tag[r] = rtag;
        }
        __syncwarp();
    }
}

}  // namespace hp

using namespace hp;

extern "C" int hp_post_solve_batch(hp_ctx* ctx, const hp_block_batch* b, const int64_t* var_pos, const uint8_t* h1,
                                   const uint8_t* h2, hp_post_out* out) {
    if (!ctx || !b || !var_pos || !h1 || !h2 || !out || !out->span_counts || !out->block_tags || !out->read_haplotag || !out->read_tag)
        return HP_ERR_INVALID_INPUT;
    if (b->n_blocks == 0) return HP_OK;
    auto fail = [&](int code, const std::string& msg) { ctx->err = msg; return code; };
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(HP_ERR_CUDA, "cudaSetDevice failed");
    const uint32_t nb = b->n_blocks;
    const uint64_t nv = b->var_off[nb], nr = b->read_off[nb], nc = b->cell_off[nr];
    for (uint64_t r = 0; r < nr; r++)
        if (b->read_end[r] < b->read_start[r] || b->cell_off[r + 1] - b->cell_off[r] != (uint64_t)(b->read_end[r] - b->read_start[r]))
            return fail(HP_ERR_INVALID_INPUT, "malformed read " + std::to_string(r));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t in_bytes = al(8 * (nb + 1)) * 2 + al(4 * nr) * 2 + al(8 * (nr + 1)) + al(nc) * 2 + al(8 * nv) + al(nv) * 2;
    const size_t out_bytes = al(4 * nv) + al(8 * nv) + al(nr) + al(8 * nr);
    if (!ctx->stage_in.reserve(in_bytes + 4096) || !ctx->stage_out.reserve(out_bytes + 4096)) return fail(HP_ERR_OUT_OF_MEMORY, "staging allocation failed");
    cudaStream_t st = ctx->stream;
    uint8_t* p = (uint8_t*)ctx->stage_in.ptr;
    bool ok = true;
    auto up = [&](const void* src, size_t bytes) { uint8_t* d = p; p += al(bytes); if (bytes) ok &= cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess; return d; };
    PostArgs a;
    a.n_blocks = nb;
    a.var_off = (const uint64_t*)up(b->var_off, 8 * (nb + 1)); a.read_off = (const uint64_t*)up(b->read_off, 8 * (nb + 1));
    a.read_start = (const uint32_t*)up(b->read_start, 4 * nr); a.read_end = (const uint32_t*)up(b->read_end, 4 * nr);
    a.cell_off = (const uint64_t*)up(b->cell_off, 8 * (nr + 1));
    a.alleles = up(b->alleles, nc); a.quals = up(b->quals, nc);
    a.var_pos = (const int64_t*)up(var_pos, 8 * nv);
    a.h1 = up(h1, nv); a.h2 = up(h2, nv);
    uint8_t* q = (uint8_t*)ctx->stage_out.ptr;
    a.span = (uint32_t*)q; q += al(4 * nv);
    a.tags = (uint64_t*)q; q += al(8 * nv);
    a.read_hap = q; q += al(nr);
    a.read_tag = (uint64_t*)q;
    const int grid = (int)std::min<uint64_t>((nb + kPostWarps - 1) / kPostWarps, (uint64_t)ctx->sm_count * 8);
    post_solve_kernel<<<grid, kPostWarps * 32, 0, st>>>(a);
    ok &= cudaGetLastError() == cudaSuccess;
    ctx->launches++;
    ok &= cudaMemcpyAsync(out->span_counts, a.span, 4 * nv, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(out->block_tags, a.tags, 8 * nv, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (nr) {
        ok &= cudaMemcpyAsync(out->read_haplotag, a.read_hap, nr, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok &= cudaMemcpyAsync(out->read_tag, a.read_tag, 8 * nr, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    }
    ok &= cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { cudaGetLastError(); return fail(HP_ERR_CUDA, "post-solve launch or copy failed"); }
    return HP_OK;
}
