// hp_realign.cu -- the realignment pipeline of a batch of phase blocks (host orchestration over the CUDA entry points).
//
// Replaces the read loop of load_full_read_segments (src/read_parsing.rs:545-629): global realignment per mapping, the
// MaxEditDistance -> local_realignment fallback (:564-575), the order-dependent switch-off of global realignment for the rest
// of a block (:593-600), and the collapse / filter that builds the read segments (:612-629).
//
// The rule is sequential per block, but a mapping's result does not depend on the order -- only WHICH of its two results is
// used does.  So the device work is batched: (1) graph-WFA for every mapping, (2) local realignment for the WFA failures,
// (3) a host replay of the counters finds each block's switch-off point, (4) local realignment for every mapping behind a
// switch-off point that does not have a local row yet, (5) rows -> hp_assemble_blocks.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/hiphase_b200.h"
#include "hp_host.h"

using namespace hp;

extern "C" int hp_realign_block_batch(hp_ctx* ctx, const hp_realign_batch* in, hp_realign_out* out) {
    if (!ctx || !in || !out || !out->map_mode || !out->block_disabled_at) return HP_ERR_INVALID_INPUT;
    const uint32_t nb = in->n_blocks;
    if (nb == 0) return HP_OK;
    if (!in->map_off || !in->map_group || !in->n_groups || !in->var_off || !in->wfa_het_base) return HP_ERR_INVALID_INPUT;
    const uint64_t nm64 = in->map_off[nb];
    if (nm64 != in->wfa.n_jobs || nm64 != in->local.n_jobs) return fail(ctx, HP_ERR_INVALID_INPUT, "wfa / local batches must hold one job per mapping");
    const uint32_t nm = (uint32_t)nm64;
    const hp_wfa_batch& W = in->wfa;
    const hp_local_batch& L = in->local;
    for (uint32_t b = 0; b < nb; b++) {
        if (in->map_off[b + 1] < in->map_off[b] || in->var_off[b + 1] < in->var_off[b]) return fail(ctx, HP_ERR_INVALID_INPUT, "offsets must be non-decreasing");
        const uint64_t n_var = in->var_off[b + 1] - in->var_off[b];
        for (uint64_t j = in->map_off[b]; j < in->map_off[b + 1]; j++) {
            if (in->map_group[j] >= in->n_groups[b]) return fail(ctx, HP_ERR_INVALID_INPUT, "read-name group out of range, mapping " + std::to_string(j));
            if ((uint64_t)(L.var_hi[j] - L.var_lo[j]) != n_var) return fail(ctx, HP_ERR_INVALID_INPUT, "local job " + std::to_string(j) + " must cover all het variants of its block");
            if (W.het_lo[j] < in->wfa_het_base[b] || (uint64_t)W.het_hi[j] > in->wfa_het_base[b] + n_var || W.het_hi[j] < W.het_lo[j])
                return fail(ctx, HP_ERR_INVALID_INPUT, "WFA job " + std::to_string(j) + " leaves the het variants of its block");
        }
    }
    if (nm == 0) {
        for (uint32_t b = 0; b < nb; b++) {
            out->block_disabled_at[b] = 0xffffffffu;
            if (out->block_failures) out->block_failures[b] = 0;
            if (out->block_parsed) out->block_parsed[b] = 0;
        }
    }

    const bool timing = getenv("HP_DBG_REALIGN_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    // ---- (1) graph-WFA for every mapping ----
    const uint64_t w_cells = nm ? W.row_off[nm] : 0, l_cells = nm ? L.row_off[nm] : 0;
    std::vector<int32_t> w_status(nm, HP_WFA_SKIPPED), l_status(nm, 0);
    std::vector<uint32_t> w_score(nm, 0);
    std::vector<uint8_t> w_al(w_cells + 1), w_q(w_cells + 1), l_al(l_cells + 1), l_q(l_cells + 1);
    if (nm) {
        hp_wfa_out wo{};
        wo.status = w_status.data(); wo.score = w_score.data(); wo.alleles = w_al.data(); wo.quals = w_q.data();
        int rc = hp_wfa_align_batch(ctx, &W, &wo);
        if (rc != HP_OK) return rc;
    }
    const auto t1 = now();
    std::vector<uint8_t> have_local(nm, 0);
    // Local realignment of the selected mappings only: their jobs are gathered into a compact batch on the host (a few percent
    // of the reads on HiFi data), so neither the copy to the device nor the validation touches the other reads.
    auto run_local = [&](std::vector<uint32_t>& sel) -> int {
        if (sel.empty()) return HP_OK;
        const size_t ns = sel.size();
        std::vector<uint32_t> c_lo(ns), c_hi(ns), c_sd, c_sl;
        std::vector<int64_t> c_pos(ns), c_sr;
        std::vector<uint64_t> c_seg_off(ns + 1, 0), c_read_off(ns + 1, 0), c_row_off(ns + 1, 0);
        uint64_t n_seg = 0, n_rd = 0;
        for (size_t k = 0; k < ns; k++) { const uint32_t j = sel[k]; n_seg += L.seg_off[j + 1] - L.seg_off[j]; n_rd += L.read_off[j + 1] - L.read_off[j]; }
        c_sr.reserve(n_seg); c_sd.reserve(n_seg); c_sl.reserve(n_seg);
        if (!ctx->realign_rb.reserve(n_rd + 1) || !ctx->realign_rq.reserve(n_rd + 1)) return fail(ctx, HP_ERR_OUT_OF_MEMORY, "pinned scratch for local realignment");
        uint8_t* c_rb = (uint8_t*)ctx->realign_rb.ptr;
        uint8_t* c_rq = (uint8_t*)ctx->realign_rq.ptr;
        for (size_t k = 0; k < ns; k++) {
            const uint32_t j = sel[k];
            c_lo[k] = L.var_lo[j]; c_hi[k] = L.var_hi[j]; c_pos[k] = L.read_pos[j];
            for (uint64_t q = L.seg_off[j]; q < L.seg_off[j + 1]; q++) { c_sr.push_back(L.seg_ref_start[q]); c_sd.push_back(L.seg_read_start[q]); c_sl.push_back(L.seg_len[q]); }
            c_seg_off[k + 1] = c_sr.size();
            const uint64_t rl = L.read_off[j + 1] - L.read_off[j];
            if (rl) { memcpy(&c_rb[c_read_off[k]], L.read_bytes + L.read_off[j], rl); memcpy(&c_rq[c_read_off[k]], L.read_quals + L.read_off[j], rl); }
            c_read_off[k + 1] = c_read_off[k] + rl;
            c_row_off[k + 1] = c_row_off[k] + (L.row_off[j + 1] - L.row_off[j]);
        }
        c_sr.push_back(0); c_sd.push_back(0); c_sl.push_back(0);
        hp_local_batch cb = L;
        cb.n_jobs = (uint32_t)ns;
        cb.var_lo = c_lo.data(); cb.var_hi = c_hi.data(); cb.read_pos = c_pos.data(); cb.seg_off = c_seg_off.data();
        cb.seg_ref_start = c_sr.data(); cb.seg_read_start = c_sd.data(); cb.seg_len = c_sl.data();
        cb.read_bytes = c_rb; cb.read_quals = c_rq; cb.read_off = c_read_off.data(); cb.row_off = c_row_off.data();
        std::vector<uint8_t> o_al(c_row_off[ns] + 1), o_q(c_row_off[ns] + 1);
        std::vector<int32_t> o_st(ns, -1);
        hp_local_out co{};
        co.alleles = o_al.data(); co.quals = o_q.data(); co.status = o_st.data();
        const auto tl0 = std::chrono::steady_clock::now();
        int rc = hp_local_realign_batch(ctx, &cb, &co);
        if (rc != HP_OK) return rc;
        if (getenv("HP_DBG_REALIGN_TIMING")) fprintf(stderr, "[hp_realign]   local call on %zu jobs: %.1f ms (kernel %.2f ms)\n", ns, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tl0).count(), hp_last_kernel_ms(ctx));
        for (size_t k = 0; k < ns; k++) {
            const uint32_t j = sel[k];
            have_local[j] = 1;
            l_status[j] = o_st[k];
            if (o_st[k] != HP_LOCAL_OK)
                return fail(ctx, HP_ERR_UNSUPPORTED, "local realignment of mapping " + std::to_string(j) + " failed with job status " +
                            std::to_string(o_st[k]) + " (the reference panics here, read_parsing.rs:320-322, 452-454)");
            const uint64_t n = c_row_off[k + 1] - c_row_off[k];
            if (n) { memcpy(&l_al[L.row_off[j]], &o_al[c_row_off[k]], n); memcpy(&l_q[L.row_off[j]], &o_q[c_row_off[k]], n); }
        }
        return HP_OK;
    };
    // ---- (2) local realignment where graph-WFA gave up (:564-575) ----
    auto wfa_failed = [&](uint32_t j) { return w_status[j] == HP_WFA_MAX_EDIT_DISTANCE || w_status[j] == HP_WFA_GRAPH_TOO_LARGE; };
    {
        std::vector<uint32_t> sel;
        for (uint32_t j = 0; j < nm; j++) {
            if (w_status[j] == HP_WFA_WORKSPACE_OVERFLOW) return fail(ctx, HP_ERR_UNSUPPORTED, "graph-WFA workspace overflow, mapping " + std::to_string(j));
            if (wfa_failed(j)) sel.push_back(j);
        }
        int rc = run_local(sel);
        if (rc != HP_OK) return rc;
    }
    // a local row is "skipped" when no variant got a binary allele (num_overlaps == 0, :492)
    auto local_skipped = [&](uint32_t j) {
        for (uint64_t c = L.row_off[j]; c < L.row_off[j + 1]; c++) if (l_al[c] < 2) return false;
        return true;
    };
    const auto t2 = now();
    // ---- (3) replay up to each block's switch-off point (:583-600) ----
    std::vector<uint32_t> late;
    for (uint32_t b = 0; b < nb; b++) {
        double failures = 0.0, parsed = 0.0;
        uint32_t disabled_at = 0xffffffffu;
        for (uint64_t j = in->map_off[b]; j < in->map_off[b + 1]; j++) {
            if (disabled_at != 0xffffffffu) { if (!have_local[j]) late.push_back((uint32_t)j); continue; }
            bool skipped, local;
            if (wfa_failed((uint32_t)j)) { local = true; skipped = local_skipped((uint32_t)j); }
            else { local = false; skipped = w_status[j] == HP_WFA_SKIPPED; }
            if (skipped) continue;
            parsed += 1.0;
            if (local) failures += 1.0;
            if (failures >= (double)in->global_failure_minimum && failures / parsed >= in->global_failure_ratio)
                disabled_at = (uint32_t)(j - in->map_off[b]);
        }
        out->block_disabled_at[b] = disabled_at;
    }
    // ---- (4) local realignment behind the switch-off points (:551-554) ----
    {
        int rc = run_local(late);
        if (rc != HP_OK) return rc;
    }
    const auto t3 = now();
    // ---- (5) final modes, counters, rows ----
    const uint32_t max_ed = ctx->params.wfa_max_edit_distance;
    std::vector<uint64_t> group_off(nb + 1, 0);
    for (uint32_t b = 0; b < nb; b++) group_off[b + 1] = group_off[b] + in->n_groups[b];
    const uint64_t n_groups = group_off[nb];
    std::vector<uint64_t> group_rows(n_groups + 1, 0);
    for (uint32_t b = 0; b < nb; b++) {
        double failures = 0.0, parsed = 0.0;
        const uint32_t dis = out->block_disabled_at[b];
        for (uint64_t j = in->map_off[b]; j < in->map_off[b + 1]; j++) {
            const uint32_t k = (uint32_t)(j - in->map_off[b]);
            uint8_t mode;
            uint32_t score;
            if (dis != 0xffffffffu && k > dis) { mode = HP_MAP_LOCAL_DISABLED; score = max_ed; }
            else if (wfa_failed((uint32_t)j)) { mode = HP_MAP_LOCAL_FAILED; score = w_score[j]; }
            else { mode = HP_MAP_GLOBAL; score = w_score[j]; }
            const bool skipped = mode == HP_MAP_GLOBAL ? w_status[j] == HP_WFA_SKIPPED : local_skipped((uint32_t)j);
            if (skipped) { mode = HP_MAP_SKIPPED; if (w_status[j] == HP_WFA_SKIPPED && !(dis != 0xffffffffu && k > dis)) score = 0xffffffffu; }   // usize::MAX, :712
            else {
                parsed += 1.0;
                if (mode != HP_MAP_GLOBAL) failures += 1.0;
                group_rows[group_off[b] + in->map_group[j] + 1]++;
            }
            out->map_mode[j] = mode;
            if (out->map_score) out->map_score[j] = score;
        }
        if (out->block_failures) out->block_failures[b] = (uint32_t)failures;
        if (out->block_parsed) out->block_parsed[b] = (uint32_t)parsed;
    }
    for (uint64_t g = 0; g < n_groups; g++) group_rows[g + 1] += group_rows[g];
    const uint64_t n_rows = group_rows[n_groups];
    std::vector<uint32_t> row_start(n_rows + 1);
    std::vector<uint64_t> row_src(n_rows + 1), row_len(n_rows + 1), row_cell_off(n_rows + 1, 0);
    std::vector<uint8_t> row_is_local(n_rows + 1);
    {
        std::vector<uint64_t> cur(group_rows.begin(), group_rows.end() - 1);
        for (uint32_t b = 0; b < nb; b++)
            for (uint64_t j = in->map_off[b]; j < in->map_off[b + 1]; j++) {
                if (out->map_mode[j] == HP_MAP_SKIPPED) continue;
                const uint64_t r = cur[group_off[b] + in->map_group[j]]++;          // mappings of a group stay in BAM order
                if (out->map_mode[j] == HP_MAP_GLOBAL) {
                    row_start[r] = W.het_lo[j] - in->wfa_het_base[b]; row_src[r] = W.row_off[j]; row_len[r] = W.row_off[j + 1] - W.row_off[j]; row_is_local[r] = 0;
                } else {
                    row_start[r] = 0; row_src[r] = L.row_off[j]; row_len[r] = L.row_off[j + 1] - L.row_off[j]; row_is_local[r] = 1;
                }
            }
    }
    for (uint64_t r = 0; r < n_rows; r++) row_cell_off[r + 1] = row_cell_off[r] + row_len[r];
    std::vector<uint8_t> r_al(row_cell_off[n_rows] + 1), r_q(row_cell_off[n_rows] + 1);
    for (uint64_t r = 0; r < n_rows; r++) {
        if (!row_len[r]) continue;
        memcpy(&r_al[row_cell_off[r]], (row_is_local[r] ? l_al.data() : w_al.data()) + row_src[r], row_len[r]);
        memcpy(&r_q[row_cell_off[r]], (row_is_local[r] ? l_q.data() : w_q.data()) + row_src[r], row_len[r]);
    }
    hp_rows_batch rows{};
    rows.n_blocks = nb; rows.var_off = in->var_off; rows.group_off = group_off.data(); rows.group_row_off = group_rows.data();
    rows.row_start = row_start.data(); rows.row_cell_off = row_cell_off.data(); rows.alleles = r_al.data(); rows.quals = r_q.data();
    rows.min_matched_alleles = in->min_matched_alleles;
    const auto t4 = now();
    const int rc_asm = hp_assemble_blocks(ctx, &rows, &out->assembled);
    if (timing) fprintf(stderr, "[hp_realign] %u mappings: wfa %.1f ms, local(failed) %.1f, replay+local(late) %.1f, rows %.1f, assemble %.1f\n", nm, ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, now()));
    return rc_asm;
}


// =================================================================================================================
// CIGAR projection of global realignment (read_parsing.rs:672-742) on the device: one thread per mapping.
// =================================================================================================================
namespace {

__global__ void wfa_plan_kernel(hp_plan_batch b, hp_plan_out o, int* bad) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= b.n_maps) return;
    const uint64_t s0 = b.seg_off[j], s1 = b.seg_off[j + 1];
    const uint32_t blk = b.map_block[j];
    if (s1 <= s0 || blk >= b.n_blocks) { atomicExch(bad, (int)j + 1); return; }              // assert!(max_position >= min_position)
    const int64_t min_position = b.seg_ref_start[s0];                                         // :677-685 (segments ascend)
    const int64_t max_position = b.seg_ref_start[s1 - 1] + (int64_t)b.seg_len[s1 - 1] - 1;
    auto lower = [](const int64_t* p, uint32_t lo, uint32_t hi, int64_t v) {                  // first index with p[i] >= v
        while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (p[mid] < v) lo = mid + 1; else hi = mid; }
        return lo;
    };
    // aligned_range.contains(position): min <= position <= max (:692-701)
    const uint32_t h0 = b.het_first[blk], h1 = b.het_first[blk + 1];
    const uint32_t het_lo = lower(b.het_pos, h0, h1, min_position), het_hi = lower(b.het_pos, h0, h1, max_position + 1);
    const uint32_t m0 = b.hom_first[blk], m1 = b.hom_first[blk + 1];
    const uint32_t hom_lo = lower(b.hom_pos, m0, m1, min_position), hom_hi = lower(b.hom_pos, m0, m1, max_position + 1);
    const bool no_het = het_hi <= het_lo, no_hom = hom_hi <= hom_lo;       // no overlap: the empty range at the block's first call
    o.ref_start[j] = (uint64_t)min_position; o.ref_end[j] = (uint64_t)(max_position + 1);
    o.het_lo[j] = no_het ? h0 : het_lo; o.het_hi[j] = no_het ? h0 : het_hi;
    o.hom_lo[j] = no_hom ? m0 : hom_lo; o.hom_hi[j] = no_hom ? m0 : hom_hi;                   // unwrap_or(0) / last stays 0 (:718-729)
    o.read_start[j] = b.seg_read_start[s0];                                                   // :737-741
    o.read_end[j] = b.seg_read_start[s1 - 1] + b.seg_len[s1 - 1];
}

}  // namespace

extern "C" int hp_wfa_plan_batch(hp_ctx* ctx, const hp_plan_batch* in, hp_plan_out* out) {
    if (!ctx || !in || !out || !out->ref_start || !out->ref_end || !out->het_lo || !out->het_hi || !out->hom_lo || !out->hom_hi ||
        !out->read_start || !out->read_end) return HP_ERR_INVALID_INPUT;
    const uint32_t nm = in->n_maps, nb = in->n_blocks;
    if (nm == 0) return HP_OK;
    if (!in->map_block || !in->seg_off || !in->het_first || !in->hom_first) return HP_ERR_INVALID_INPUT;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, HP_ERR_CUDA, "cudaSetDevice failed");
    const uint64_t ns = in->seg_off[nm];
    const uint64_t nh = in->het_first[nb], nhm = in->hom_first[nb];
    for (uint32_t b = 0; b < nb; b++) {
        if (in->het_first[b + 1] < in->het_first[b] || in->hom_first[b + 1] < in->hom_first[b]) return fail(ctx, HP_ERR_INVALID_INPUT, "call ranges must be non-decreasing");
        for (uint32_t k = in->het_first[b]; k + 1 < in->het_first[b + 1]; k++) if (in->het_pos[k] > in->het_pos[k + 1]) return fail(ctx, HP_ERR_INVALID_INPUT, "het positions of a block must ascend");
        for (uint32_t k = in->hom_first[b]; k + 1 < in->hom_first[b + 1]; k++) if (in->hom_pos[k] > in->hom_pos[k + 1]) return fail(ctx, HP_ERR_INVALID_INPUT, "hom positions of a block must ascend");
    }
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t in_bytes = al(4ull * nm) + al(8ull * (nm + 1)) + al(8 * ns) + al(4 * ns) * 2 + al(4ull * (nb + 1)) * 2 + al(8 * nh) + al(8 * nhm) + 4096;
    const size_t out_bytes = al(8ull * nm) * 2 + al(4ull * nm) * 6 + 4096;
    if (!ctx->stage_in.reserve(in_bytes) || !ctx->stage_out.reserve(out_bytes)) return fail(ctx, HP_ERR_OUT_OF_MEMORY, "staging allocation failed");
    cudaStream_t st = ctx->stream;
    uint8_t* p = (uint8_t*)ctx->stage_in.ptr;
    bool ok = true;
    auto up = [&](const void* src, size_t bytes) { uint8_t* d = p; p += al(bytes); if (bytes) ok &= cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess; return d; };
    hp_plan_batch d = *in;
    d.map_block = (const uint32_t*)up(in->map_block, 4ull * nm); d.seg_off = (const uint64_t*)up(in->seg_off, 8ull * (nm + 1));
    d.seg_ref_start = (const int64_t*)up(in->seg_ref_start, 8 * ns); d.seg_read_start = (const uint32_t*)up(in->seg_read_start, 4 * ns);
    d.seg_len = (const uint32_t*)up(in->seg_len, 4 * ns);
    d.het_first = (const uint32_t*)up(in->het_first, 4ull * (nb + 1)); d.het_pos = (const int64_t*)up(in->het_pos, 8 * nh);
    d.hom_first = (const uint32_t*)up(in->hom_first, 4ull * (nb + 1)); d.hom_pos = (const int64_t*)up(in->hom_pos, 8 * nhm);
    uint8_t* q = (uint8_t*)ctx->stage_out.ptr;
    auto carve = [&](size_t bytes) { uint8_t* r = q; q += al(bytes); return r; };
    hp_plan_out o;
    o.ref_start = (uint64_t*)carve(8ull * nm); o.ref_end = (uint64_t*)carve(8ull * nm);
    o.het_lo = (uint32_t*)carve(4ull * nm); o.het_hi = (uint32_t*)carve(4ull * nm); o.hom_lo = (uint32_t*)carve(4ull * nm); o.hom_hi = (uint32_t*)carve(4ull * nm);
    o.read_start = (uint32_t*)carve(4ull * nm); o.read_end = (uint32_t*)carve(4ull * nm);
    int* bad = (int*)carve(4);
    ok &= cudaMemsetAsync(bad, 0, 4, st) == cudaSuccess;
    wfa_plan_kernel<<<(nm + 255) / 256, 256, 0, st>>>(d, o, bad);
    ok &= cudaGetLastError() == cudaSuccess;
    ctx->launches++;
    int h_bad = 0;
    ok &= cudaMemcpyAsync(out->ref_start, o.ref_start, 8ull * nm, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(out->ref_end, o.ref_end, 8ull * nm, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(out->het_lo, o.het_lo, 4ull * nm, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(out->het_hi, o.het_hi, 4ull * nm, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(out->hom_lo, o.hom_lo, 4ull * nm, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(out->hom_hi, o.hom_hi, 4ull * nm, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(out->read_start, o.read_start, 4ull * nm, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(out->read_end, o.read_end, 4ull * nm, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { cudaGetLastError(); return fail(ctx, HP_ERR_CUDA, "CIGAR projection launch or copy failed"); }
    if (h_bad) return fail(ctx, HP_ERR_INVALID_INPUT, "mapping " + std::to_string(h_bad - 1) + " has no aligned segment or an invalid block (the reference asserts, read_parsing.rs:686)");
    return HP_OK;
}
