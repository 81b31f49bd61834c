// local_oracle.cpp -- CPU ORACLE (test infrastructure only, see hp_oracle.h) for the local-realignment fallback:
//   sequence_alignment::edit_distance        src/sequence_alignment.rs:6-38   (full grid, two rows)
//   Variant::match_allele                    src/data_types/variants.rs:598-606
//   Variant::closest_allele_clip             src/data_types/variants.rs:624-641
//   local_realignment                        src/read_parsing.rs:121-503
// Written as a literal restatement: a hash map from reference coordinate to read index filled from the aligned
// pairs (:137-146), coordinate-by-coordinate scans, the f64 harmonic-mean quality in iteration order (:293-327).
#include "hp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

// sequence_alignment.rs:6-38
size_t edit_distance(const uint8_t* v1, size_t l1, const uint8_t* v2, size_t l2) {
    std::vector<size_t> row(l1 + 1, 0), prev(l1 + 1);
    for (size_t j = 0; j <= l1; j++) prev[j] = j;
    for (size_t i = 0; i < l2; i++) {
        row[0] = i + 1;
        for (size_t j = 0; j < l1; j++) {
            const size_t a = prev[j + 1] + 1, b = row[j] + 1, c = prev[j] + (v1[j] == v2[i] ? 0 : 1);
            row[j + 1] = std::min(a, std::min(b, c));
        }
        std::swap(row, prev);
    }
    return prev[l1];
}

struct VarView {
    int64_t pos;
    size_t ref_len, prefix_len, postfix_len;
    const uint8_t *a0, *a1;
    size_t l0, l1;
    uint8_t vtype, ignored;
};

VarView view(const hp_local_batch* b, uint32_t k) {
    const hp_variant_table& t = b->variants;
    VarView v;
    v.pos = t.position[k]; v.ref_len = t.ref_len[k]; v.prefix_len = b->prefix_len[k]; v.postfix_len = b->postfix_len[k];
    v.a0 = t.allele_bytes + t.allele0_off[k]; v.l0 = t.allele0_len[k];
    v.a1 = t.allele_bytes + t.allele1_off[k]; v.l1 = t.allele1_len[k];
    v.vtype = t.vtype[k]; v.ignored = t.ignored[k];
    return v;
}

// variants.rs:598-606
uint8_t match_allele(const VarView& v, const uint8_t* s, size_t n) {
    if (n == v.l0 && std::memcmp(s, v.a0, n) == 0) return 0;
    if (n == v.l1 && std::memcmp(s, v.a1, n) == 0) return 1;
    return 2;
}

// variants.rs:624-641
uint8_t closest_allele_clip(const VarView& v, const uint8_t* s, size_t n, size_t head, size_t tail, size_t* d0o, size_t* d1o) {
    const size_t d0 = edit_distance(s, n, v.a0 + head, v.l0 - tail - head);
    const size_t d1 = edit_distance(s, n, v.a1 + head, v.l1 - tail - head);
    *d0o = d0; *d1o = d1;
    return d0 < d1 ? 0 : (d0 > d1 ? 1 : 2);
}

constexpr double SNV_QUAL = 80, TR_QUAL = 40, SV_INDEL_QUAL = 20, INDEL_QUAL = 10;   // read_parsing.rs:18-21

// Rust `f as u8`: saturating, NaN -> 0
uint8_t as_u8(double f) {
    if (!(f == f)) return 0;
    if (f <= 0.0) return 0;
    if (f >= 255.0) return 255;
    return (uint8_t)f;
}
// Rust f64::min / f64::max: the non-NaN operand wins
double rmin(double a, double b) { return std::fmin(a, b); }
double rmax(double a, double b) { return std::fmax(a, b); }

}  // namespace

extern "C" uint64_t hpo_edit_distance(const uint8_t* a, uint64_t la, const uint8_t* b, uint64_t lb) {
    return (uint64_t)edit_distance(a, (size_t)la, b, (size_t)lb);
}

/* match_allele / closest_allele_clip of variant k of the batch's table against an observed sequence. */
extern "C" uint8_t hpo_match_allele(const hp_local_batch* b, uint32_t k, const uint8_t* s, uint64_t n) {
    return match_allele(view(b, k), s, (size_t)n);
}
extern "C" uint8_t hpo_closest_allele_clip(const hp_local_batch* b, uint32_t k, const uint8_t* s, uint64_t n, uint64_t head,
                                          uint64_t tail, uint64_t* d0, uint64_t* d1) {
    size_t x0, x1;
    const uint8_t r = closest_allele_clip(view(b, k), s, (size_t)n, (size_t)head, (size_t)tail, &x0, &x1);
    // the reference returns (allele, min, other); report both distances and let the caller order them
    *d0 = x0; *d1 = x1;
    return r;
}

extern "C" int hpo_local_realign_batch(const hp_local_batch* b, hp_local_out* out) {
    for (uint32_t j = 0; j < b->n_jobs; j++) {
        int32_t status = HP_LOCAL_OK;
        // :137-146 reference coordinate -> read index
        std::unordered_map<int64_t, int64_t> lookup;
        const int64_t min_position = b->read_pos[j];
        int64_t max_position = min_position;
        for (uint64_t s = b->seg_off[j]; s < b->seg_off[j + 1]; s++)
            for (uint32_t i = 0; i < b->seg_len[s]; i++) {
                const int64_t rc = b->seg_ref_start[s] + i;
                lookup[rc] = (int64_t)b->seg_read_start[s] + i;
                max_position = std::max(max_position, rc);
            }
        const int64_t range_start = min_position, range_end = max_position + 1;        // :150
        auto in_range = [&](int64_t x) { return x >= range_start && x < range_end; };
        auto get = [&](int64_t rc, int64_t* v) { auto it = lookup.find(rc); if (it == lookup.end()) return false; *v = it->second; return true; };
        const uint8_t* seq = b->read_bytes + b->read_off[j];
        const uint8_t* rq = b->read_quals + b->read_off[j];
        const size_t read_len = (size_t)(b->read_off[j + 1] - b->read_off[j]);

        size_t last_deletion_end = 0;
        const uint64_t row = b->row_off[j];
        for (uint32_t k = b->var_lo[j]; k < b->var_hi[j]; k++) {
            const VarView v = view(b, k);
            uint8_t allele = HP_ALLELE_NOOVERLAP, qual = 0;
            bool exact = false, overlaps = false;
            size_t d0 = 0, d1 = 0;
            if (v.ignored) {                                                          // :179-185
                allele = HP_ALLELE_NOOVERLAP;
            } else if (v.pos < (int64_t)last_deletion_end) {                          // :186-193
                allele = HP_ALLELE_AMBIGUOUS; overlaps = true;
            } else if (v.vtype == HP_VT_SNV || v.vtype == HP_VT_INSERTION || v.vtype == HP_VT_DELETION || v.vtype == HP_VT_INDEL ||
                       v.vtype == HP_VT_SV_INSERTION || v.vtype == HP_VT_TANDEM_REPEAT) {
                const size_t first_start = (size_t)v.pos - v.prefix_len, last_start = (size_t)v.pos + 1;       // :208-211
                const size_t first_end = (size_t)v.pos + v.ref_len, last_end = (size_t)v.pos + v.ref_len + v.postfix_len + 1;
                bool has_cs = false, has_ce = false;
                size_t closest_start = 0, closest_end = 0;
                int64_t t;
                for (size_t sc = last_start; sc-- > first_start;)                     // :214-220 (reverse scan)
                    if (get((int64_t)sc, &t)) { closest_start = (size_t)t; has_cs = true; break; }
                for (size_t ec = first_end; ec < last_end; ec++)                      // :223-229
                    if (get((int64_t)ec, &t)) { closest_end = (size_t)t; has_ce = true; break; }
                bool has_s = false, has_e = false;
                size_t ss = 0, se = 0, start_clip = 0, end_clip = 0;
                if (has_cs && has_ce) {                                               // :237-270
                    for (size_t sc = first_start; sc < last_start; sc++) {
                        start_clip++;
                        if (get((int64_t)sc, &t)) {
                            if (closest_start - (size_t)t > 2 * v.prefix_len) continue;
                            ss = (size_t)t; has_s = true;
                            for (size_t ec = last_end; ec-- > first_end;) {
                                end_clip++;
                                int64_t u;
                                if (get((int64_t)ec, &u)) {
                                    if ((size_t)u - closest_end > 2 * v.postfix_len) continue;
                                    se = (size_t)u; has_e = true;
                                    break;
                                }
                            }
                            break;
                        }
                    }
                }
                if (has_s) {
                    if (has_e) {
                        if (se < ss || se > read_len) { status = HP_LOCAL_BAD_SLICE; allele = HP_ALLELE_AMBIGUOUS; overlaps = true; }
                        else {
                            allele = match_allele(v, seq + ss, se - ss);               // :280
                            if (allele == HP_ALLELE_AMBIGUOUS) {
                                allele = closest_allele_clip(v, seq + ss, se - ss, start_clip - 1, end_clip - 1, &d0, &d1);   // :283
                                exact = false;
                            } else exact = true;
                            double sum = 0.0;                                         // :293-296
                            for (size_t i = ss; i < se; i++) sum += 1.0 / (double)rq[i];
                            const double harmonic = (double)(se - ss) / sum;
                            const double factor = rmin(harmonic / 40.0, 1.0);         // :299
                            const double base = v.vtype == HP_VT_SNV ? SNV_QUAL
                                              : v.vtype == HP_VT_TANDEM_REPEAT ? TR_QUAL
                                              : v.vtype == HP_VT_SV_INSERTION ? SV_INDEL_QUAL : INDEL_QUAL;    // :302-323
                            qual = as_u8(rmax(base * factor, 1.0));                   // :327
                            overlaps = true;
                        }
                    } else { allele = HP_ALLELE_AMBIGUOUS; overlaps = true; }          // :331-337
                } else if (in_range(v.pos)) { allele = HP_ALLELE_AMBIGUOUS; overlaps = true; }   // :340-343
                else { allele = HP_ALLELE_NOOVERLAP; overlaps = false; }              // :344-349
            } else if (v.vtype == HP_VT_SV_DELETION) {                                // :354-451
                if (in_range(v.pos)) {
                    const size_t last_start = (size_t)v.pos + 1, first_end = (size_t)v.pos + v.ref_len;
                    if (in_range((int64_t)first_end)) {
                        const size_t expected_deleted = first_end - last_start;
                        size_t start_anchor = last_start;
                        while (!lookup.count((int64_t)start_anchor)) {
                            if (start_anchor <= (size_t)range_start) break;
                            start_anchor--;
                        }
                        size_t end_anchor = first_end;
                        while (!lookup.count((int64_t)end_anchor)) {
                            end_anchor++;
                            if (end_anchor >= (size_t)range_end) break;
                        }
                        size_t deleted = 0;
                        for (size_t dc = start_anchor; dc < end_anchor; dc++)
                            if (!lookup.count((int64_t)dc)) deleted++;
                        const double win = 0.33;
                        const double ratio = (double)deleted / (double)expected_deleted;
                        if (ratio < win) {
                            allele = HP_ALLELE_REFERENCE;
                            qual = as_u8(rmax(SV_INDEL_QUAL * (1.0 - ratio), 1.0));
                            exact = ratio == 0.0;
                        } else if (std::fabs(1.0 - ratio) < win) {
                            allele = HP_ALLELE_ALTERNATE;
                            qual = as_u8(rmax(SV_INDEL_QUAL * (1.0 - std::fabs(1.0 - ratio)), 1.0));
                            exact = ratio == 1.0;
                            last_deletion_end = first_end;                            // :428
                        } else allele = HP_ALLELE_AMBIGUOUS;
                        overlaps = true;
                    } else { allele = HP_ALLELE_AMBIGUOUS; overlaps = true; }          // :436-442
                } else { allele = HP_ALLELE_NOOVERLAP; overlaps = false; }            // :443-449
            } else {
                status = HP_LOCAL_UNHANDLED_TYPE;                                     // :452-454 panic!
                allele = HP_ALLELE_NOOVERLAP;
            }
            const uint64_t c = row + (k - b->var_lo[j]);
            out->alleles[c] = allele; out->quals[c] = qual;
            if (out->match_class) out->match_class[c] = (uint8_t)((overlaps ? HP_LOCAL_OVERLAPS : 0) | (exact ? HP_LOCAL_EXACT : 0));
            if (out->edit_distance) { out->edit_distance[2 * c] = (uint32_t)d0; out->edit_distance[2 * c + 1] = (uint32_t)d1; }
        }
        out->status[j] = status;
    }
    return 0;
}
