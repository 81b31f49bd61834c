// post_oracle.cpp -- CPU ORACLE (test infrastructure only, see hp_oracle.h) for the post-solve step:
//   get_solution_span_counts   src/phaser.rs:350-388
//   block_split / block_tags   src/phaser.rs:546-569
//   haplotag_reads             src/phaser.rs:714-750   (score_haplotype: src/data_types/read_segments.rs:161-168)
#include "hp_oracle.h"

#include <vector>

extern "C" int hpo_post_solve_batch(const hp_block_batch* b, const int64_t* var_pos, const uint8_t* h1g, const uint8_t* h2g,
                                    hp_post_out* out) {
    for (uint32_t blk = 0; blk < b->n_blocks; blk++) {
        const uint64_t v0 = b->var_off[blk];
        const size_t N = (size_t)(b->var_off[blk + 1] - v0);
        const uint8_t* h1 = h1g + v0;
        const uint8_t* h2 = h2g + v0;
        std::vector<uint64_t> span(N > 0 ? N - 1 : 0, 0);
        // ---- get_solution_span_counts ----
        for (uint64_t r = b->read_off[blk]; r < b->read_off[blk + 1]; r++) {
            size_t js = b->read_start[r], je = b->read_end[r];
            if (je == 0) continue;
            je -= 1;                                                   // junctures, not alleles (:366)
            while (js < je && h1[js] == h2[js]) js++;                  // :370-373
            while (js < je && h1[je] == h2[je]) je--;                  // :376-379
            for (size_t i = js; i < je; i++) span[i]++;
        }
        for (size_t i = 0; i + 1 < N; i++) out->span_counts[v0 + i] = (uint32_t)span[i];
        if (N) out->span_counts[v0 + N - 1] = 0;
        // ---- block tags (:546-569): a juncture without spanning reads starts a new sub-block ----
        uint64_t tag = N ? (uint64_t)var_pos[v0] : 0;
        for (size_t i = 0; i < N; i++) {
            if (i > 0 && span[i - 1] == 0) tag = (uint64_t)var_pos[v0 + i];
            out->block_tags[v0 + i] = tag;
        }
        // ---- haplotag_reads (:714-750) ----
        for (uint64_t r = b->read_off[blk]; r < b->read_off[blk + 1]; r++) {
            const size_t s = b->read_start[r], e = b->read_end[r];
            const uint8_t* al = b->alleles + b->cell_off[r];
            const uint8_t* ql = b->quals + b->cell_off[r];
            uint64_t a1 = 0, a2 = 0;
            for (size_t i = s; i < e; i++) {
                if (h1[i] < 2 && al[i - s] != h1[i]) a1 += ql[i - s];
                if (h2[i] < 2 && al[i - s] != h2[i]) a2 += ql[i - s];
            }
            uint8_t tagv = a1 < a2 ? 0 : (a1 > a2 ? 1 : 2);
            uint64_t rtag = 0;
            if (tagv != 2) {
                size_t fv = s;
                while (fv < e && (h1[fv] == h2[fv] || al[fv - s] >= 2)) fv++;   // unequal scores imply such a variant exists
                if (fv >= e) return 1;                                           // the reference would index out of bounds
                rtag = out->block_tags[v0 + fv];
            }
            out->read_haplotag[r] = tagv;
            out->read_tag[r] = rtag;
        }
    }
    return 0;
}
