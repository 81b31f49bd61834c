// realign_oracle.cpp -- CPU ORACLE (test infrastructure only, see hp_oracle.h): the read loop of load_full_read_segments,
// src/read_parsing.rs:545-629, restated statement by statement on top of the oracle's global_realignment
// (hpo_wfa_align_batch on a one-job view = read_parsing.rs:653-867) and local_realignment (hpo_local_realign_batch on a
// one-job view = read_parsing.rs:121-503).  One mapping at a time, in BAM order, exactly like the reference: nothing is batched
// or reordered here, so the order-dependent switch-off rule (:593-600) is exercised the way the reference runs it.
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "hp_oracle.h"

namespace {

// view of job j as a one-job batch (per-job arrays advanced by j; the CSR offsets stay absolute)
hp_wfa_batch wfa_view(const hp_wfa_batch& b, uint32_t j) {
    hp_wfa_batch v = b;
    v.n_jobs = 1;
    v.ref_start += j; v.ref_end += j; v.het_lo += j; v.het_hi += j; v.hom_lo += j; v.hom_hi += j; v.read_off += j; v.row_off += j;
    return v;
}
hp_local_batch local_view(const hp_local_batch& b, uint32_t j) {
    hp_local_batch v = b;
    v.n_jobs = 1;
    v.var_lo += j; v.var_hi += j; v.read_pos += j; v.seg_off += j; v.read_off += j; v.row_off += j;
    return v;
}

}  // namespace

extern "C" int hpo_realign_block_batch(const hp_params* params, const hp_realign_batch* in, hp_realign_out* out) {
    const uint32_t nb = in->n_blocks;
    const uint64_t nm = in->map_off[nb];
    const hp_wfa_batch& W = in->wfa;
    const hp_local_batch& L = in->local;
    std::vector<uint8_t> w_al(W.row_off[nm] + 1), w_q(W.row_off[nm] + 1), l_al(L.row_off[nm] + 1), l_q(L.row_off[nm] + 1);
    std::vector<int32_t> w_status(nm + 1), l_status(nm + 1);
    std::vector<uint32_t> w_score(nm + 1);

    // rows of the block batch handed to the assembly: read_groups (:528), one entry per read name
    std::vector<uint64_t> group_off(nb + 1, 0);
    for (uint32_t b = 0; b < nb; b++) group_off[b + 1] = group_off[b] + in->n_groups[b];
    struct Row { uint32_t start; std::vector<uint8_t> al, q; };
    std::vector<std::vector<Row>> groups(group_off[nb]);

    for (uint32_t b = 0; b < nb; b++) {
        bool global_disabled = false;                                                        // :534
        double num_global_failures = 0.0, total_parsed = 0.0;                                // :535-536
        out->block_disabled_at[b] = 0xffffffffu;
        for (uint64_t j = in->map_off[b]; j < in->map_off[b + 1]; j++) {
            std::vector<uint8_t> alleles, quals;
            uint32_t row_start = 0;
            uint64_t skipped_reads = 0, local_aligned = 0;
            uint32_t wfa_score = 0;
            uint8_t mode;
            auto run_local = [&]() {                                                         // local_realignment(&read, variant_calls)
                hp_local_batch v = local_view(L, (uint32_t)j);
                hp_local_out lo{};
                lo.alleles = l_al.data(); lo.quals = l_q.data(); lo.status = l_status.data() + j;
                const int rc = hpo_local_realign_batch(&v, &lo);
                if (rc != 0 || l_status[j] != HP_LOCAL_OK) return false;
                alleles.assign(l_al.begin() + L.row_off[j], l_al.begin() + L.row_off[j + 1]);
                quals.assign(l_q.begin() + L.row_off[j], l_q.begin() + L.row_off[j + 1]);
                row_start = 0;
                uint64_t num_overlaps = 0;
                for (uint8_t a : alleles) if (a < 2) num_overlaps++;                          // :474-475
                skipped_reads = num_overlaps == 0 ? 1 : 0;                                   // :492
                local_aligned = 1 - skipped_reads;                                           // :493
                return true;
            };
            if (global_disabled) {                                                           // :551-554
                if (!run_local()) return 1;
                wfa_score = params->wfa_max_edit_distance;
                mode = HP_MAP_LOCAL_DISABLED;
            } else {                                                                         // :555-579
                hp_wfa_batch v = wfa_view(W, (uint32_t)j);
                hp_wfa_out wo{};
                wo.status = w_status.data() + j; wo.score = w_score.data() + j; wo.alleles = w_al.data(); wo.quals = w_q.data();
                if (hpo_wfa_align_batch(params, &v, &wo, 1) != 0) return 1;
                if (w_status[j] == HP_WFA_MAX_EDIT_DISTANCE) {                               // :564-575
                    if (!run_local()) return 1;
                    wfa_score = w_score[j];
                    mode = HP_MAP_LOCAL_FAILED;
                } else {
                    skipped_reads = w_status[j] == HP_WFA_SKIPPED ? 1 : 0;                   // :703-712
                    alleles.assign(w_al.begin() + W.row_off[j], w_al.begin() + W.row_off[j + 1]);
                    quals.assign(w_q.begin() + W.row_off[j], w_q.begin() + W.row_off[j + 1]);
                    row_start = W.het_lo[j] - in->wfa_het_base[b];
                    wfa_score = w_score[j];
                    mode = HP_MAP_GLOBAL;
                }
            }
            if (skipped_reads == 0) {                                                        // :584
                groups[group_off[b] + in->map_group[j]].push_back(Row{row_start, alleles, quals});   // :587-588
                num_global_failures += (double)local_aligned;                                // :593
                total_parsed += 1.0;                                                         // :594
                if (!global_disabled && num_global_failures >= (double)in->global_failure_minimum &&
                    num_global_failures / total_parsed >= in->global_failure_ratio) {        // :597-600
                    global_disabled = true;
                    out->block_disabled_at[b] = (uint32_t)(j - in->map_off[b]);
                }
            } else mode = HP_MAP_SKIPPED;                                                    // :601-604
            out->map_mode[j] = mode;
            if (out->map_score) out->map_score[j] = wfa_score;
        }
        if (out->block_failures) out->block_failures[b] = (uint32_t)num_global_failures;
        if (out->block_parsed) out->block_parsed[b] = (uint32_t)total_parsed;
    }
    // :612-629 collapse + filter, through the oracle's assembly
    std::vector<uint64_t> group_row_off(1, 0), row_cell_off(1, 0);
    std::vector<uint32_t> row_start;
    std::vector<uint8_t> al, q;
    for (const auto& g : groups) {
        for (const Row& r : g) {
            row_start.push_back(r.start);
            al.insert(al.end(), r.al.begin(), r.al.end()); q.insert(q.end(), r.q.begin(), r.q.end());
            row_cell_off.push_back(al.size());
        }
        group_row_off.push_back(row_start.size());
    }
    al.push_back(0); q.push_back(0); row_start.push_back(0);
    hp_rows_batch rows{};
    rows.n_blocks = nb; rows.var_off = in->var_off; rows.group_off = group_off.data(); rows.group_row_off = group_row_off.data();
    rows.row_start = row_start.data(); rows.row_cell_off = row_cell_off.data(); rows.alleles = al.data(); rows.quals = q.data();
    rows.min_matched_alleles = in->min_matched_alleles;
    return hpo_assemble_blocks(&rows, &out->assembled);
}


// The pre-WFA half of global_realignment, src/read_parsing.rs:672-742, per mapping: aligned pairs -> min / max position, the
// overlapped het / hom calls and the aligned slice of the read.  Loops over every aligned pair and every call like the reference.
extern "C" int hpo_wfa_plan_batch(const hp_plan_batch* b, hp_plan_out* o) {
    for (uint32_t j = 0; j < b->n_maps; j++) {
        int64_t min_position = INT64_MAX, max_position = INT64_MIN;                          // :675-676
        int64_t read_at_min = 0, read_at_max = 0;
        for (uint64_t s = b->seg_off[j]; s < b->seg_off[j + 1]; s++)
            for (uint32_t k = 0; k < b->seg_len[s]; k++) {                                   // for bp in read.aligned_pairs() (:677-684)
                const int64_t ref_index = b->seg_ref_start[s] + k, segment_index = (int64_t)b->seg_read_start[s] + k;
                if (ref_index < min_position) { min_position = ref_index; read_at_min = segment_index; }
                if (ref_index > max_position) { max_position = ref_index; read_at_max = segment_index; }
            }
        if (!(max_position >= min_position)) return 1;                                       // :686
        const uint32_t blk = b->map_block[j];
        bool have_first = false; uint32_t first_overlap = b->het_first[blk], last_overlap = b->het_first[blk];
        for (uint32_t i = b->het_first[blk]; i < b->het_first[blk + 1]; i++)                 // :692-701
            if (b->het_pos[i] >= min_position && b->het_pos[i] < max_position + 1) {
                if (!have_first) { first_overlap = i; have_first = true; }
                last_overlap = i + 1;
            }
        if (!have_first) last_overlap = first_overlap;
        bool have_hom = false; uint32_t first_hom = b->hom_first[blk], last_hom = b->hom_first[blk];
        for (uint32_t i = b->hom_first[blk]; i < b->hom_first[blk + 1]; i++)                 // :718-729
            if (b->hom_pos[i] >= min_position && b->hom_pos[i] < max_position + 1) {
                if (!have_hom) { first_hom = i; have_hom = true; }
                last_hom = i + 1;
            }
        if (!have_hom) last_hom = first_hom;
        o->ref_start[j] = (uint64_t)min_position; o->ref_end[j] = (uint64_t)(max_position + 1);
        o->het_lo[j] = first_overlap; o->het_hi[j] = last_overlap; o->hom_lo[j] = first_hom; o->hom_hi[j] = last_hom;
        o->read_start[j] = (uint32_t)read_at_min; o->read_end[j] = (uint32_t)read_at_max + 1;  // :737-741
    }
    return 0;
}
