// local_kernels.cu -- local realignment on sm_100a.  Replaces local_realignment
// (src/read_parsing.rs:121-503), Variant::match_allele / closest_allele_clip (src/data_types/variants.rs:598-641) and
// sequence_alignment::edit_distance (src/sequence_alignment.rs:6-38).
//
// One warp per job (read mapping), persistent, jobs pulled from an atomic ticket.  Three passes over the job's row:
//   A  lanes over variants: anchor search on the aligned segments (binary search instead of the reference's hash map of
//      every aligned pair), exact allele match, the f64 harmonic-mean quality (summed in read order like the reference,
//      so the rounding is identical), SV-deletion ratio.  Inexact cells are left pending.
//   B  edit distances of the pending cells.  Unit-cost Levenshtein distance does not depend on how it is computed, so
//      instead of the reference's full grid the kernel runs Myers' bit-vector recurrence (global variant: the horizontal
//      delta entering row 0 is +1): one lane per comparison when the shorter sequence has <= 64 bases, and a
//      warp-systolic multi-word version (lane b owns 64-base blocks, text characters flow from lane to lane with the
//      horizontal carries) for long x long comparisons (SV insertions).
//   C  the order-dependent part (:186-193, :428): an SV deletion called ALT masks every later variant that starts inside
//      it.  Resolved with ballots over the few SV-deletion cells, in variant order.
#include <algorithm>
#include <string>

#include "hp_host.h"

namespace hp {

struct LocalArgs {
    uint32_t n_jobs;
    // variant table
    const int64_t* position;
    const uint32_t* ref_len;
    const uint32_t* prefix_len;
    const uint32_t* postfix_len;
    const uint64_t* a0_off;
    const uint32_t* a0_len;
    const uint64_t* a1_off;
    const uint32_t* a1_len;
    const uint8_t* vtype;
    const uint8_t* ignored;
    const uint8_t* allele_bytes;
    // jobs
    const uint32_t* var_lo;
    const uint32_t* var_hi;
    const int64_t* read_pos;
    const uint64_t* seg_off;
    const int64_t* seg_ref;
    const uint32_t* seg_read;
    const uint32_t* seg_len;
    const uint8_t* read_bytes;
    const uint8_t* read_quals;
    const uint64_t* read_off;
    const uint64_t* row_off;
    // outputs
    uint8_t* alleles;
    uint8_t* quals;
    uint8_t* mclass;       // HP_LOCAL_OVERLAPS | HP_LOCAL_EXACT (+ internal bits while the job is in flight)
    uint32_t* ed;          // [2 * cells]
    int32_t* status;
    // scratch per cell
    uint32_t* t_ss;        // pending: slice start / end in the read, head / tail clip
    uint32_t* t_se;
    uint32_t* t_clip;      // head | tail << 16 ... widened below
    uint32_t* t_clip2;
    int64_t* del_end;      // SV deletion called ALT: first_end_coordinate
    uint32_t* ticket;
    int8_t* ed_scratch;    // per-warp row of horizontal deltas for patterns beyond one panel (nullptr: none needed)
    uint64_t ed_stride;
};

constexpr uint32_t kPending = 4u, kUnhandled = 8u, kBadSlice = 16u, kTooLong = 32u, kSvDelAlt = 64u;
constexpr int kLocalWarps = 4;
constexpr int kCoopK = 8;                    // 64-base blocks per lane in the systolic path: 32*8*64 = 16384 pattern rows per panel
constexpr uint32_t kEdTooLong = 0xffffffffu;

__device__ __forceinline__ uint32_t base_code(uint8_t c) {
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
}

// reference coordinate -> read index through the aligned segments (the HashMap of read_parsing.rs:137-146)
struct Segs {
    const int64_t* ref;
    const uint32_t* rd;
    const uint32_t* len;
    uint32_t n;
    // index of the last segment with ref_start <= rc, or -1
    __device__ __forceinline__ int last_le(int64_t rc) const {
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ref[mid] <= rc) lo = mid + 1; else hi = mid;
        }
        return (int)lo - 1;
    }
    __device__ __forceinline__ bool get(int64_t rc, int64_t& out) const {
        const int s = last_le(rc);
        if (s < 0) return false;
        const int64_t d = rc - ref[s];
        if (d >= (int64_t)len[s]) return false;
        out = (int64_t)rd[s] + d;
        return true;
    }
    __device__ __forceinline__ bool mapped(int64_t rc) const { int64_t t; return get(rc, t); }
    // number of mapped coordinates in [a, b)
    __device__ __forceinline__ int64_t mapped_in(int64_t a, int64_t b) const {
        if (b <= a) return 0;
        int s = last_le(a);
        if (s < 0) s = 0;
        int64_t cnt = 0;
        for (; s < (int)n && ref[s] < b; s++) {
            const int64_t lo = ref[s] > a ? ref[s] : a;
            const int64_t e = ref[s] + (int64_t)len[s];
            const int64_t hi = e < b ? e : b;
            if (hi > lo) cnt += hi - lo;
        }
        return cnt;
    }
};

// Myers / Hyyro bit-vector Levenshtein distance, pattern of m <= 64 bytes held by one thread (global distance: the
// horizontal delta entering row 0 is +1).
__device__ uint32_t ed_small(const uint8_t* pat, uint32_t m, const uint8_t* txt, uint32_t n) {
    if (m == 0) return n;
    uint64_t pA = 0, pC = 0, pG = 0, pT = 0;
    for (uint32_t j = 0; j < m; j++) {
        const uint32_t c = base_code(pat[j]);
        const uint64_t bit = 1ull << j;
        if (c == 0) pA |= bit; else if (c == 1) pC |= bit; else if (c == 2) pG |= bit; else if (c == 3) pT |= bit;
    }
    uint64_t Pv = ~0ull, Mv = 0;
    uint32_t score = m;
    const uint64_t top = 1ull << (m - 1);
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t ch = txt[i];
        const uint32_t c = base_code(ch);
        uint64_t Eq;
        if (c == 0) Eq = pA; else if (c == 1) Eq = pC; else if (c == 2) Eq = pG; else if (c == 3) Eq = pT;
        else { Eq = 0; for (uint32_t j = 0; j < m; j++) if (pat[j] == ch) Eq |= 1ull << j; }
        const uint64_t Xv = Eq | Mv;
        const uint64_t Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
        uint64_t Ph = Mv | ~(Xh | Pv);
        uint64_t Mh = Pv & Xh;
        if (Ph & top) score++; else if (Mh & top) score--;
        Ph = (Ph << 1) | 1ull;
        Mh <<= 1;
        Pv = Mh | ~(Xv | Ph);
        Mv = Ph & Xv;
    }
    return score;
}

// Warp-systolic multi-word version: lane b owns the 64-base pattern blocks [b*K, b*K+K); at step t it processes text
// character t-b, taking the horizontal carry of its first block from lane b-1's last block of the previous step.
// peq: shared memory, [(k*4 + letter)*32 + lane].  Every lane returns the distance.  m > 64 (else use ed_small).
// Patterns beyond one panel (kPanelRows = 32 lanes x kCoopK blocks x 64 rows) are processed panel after panel: the horizontal
// deltas leaving a panel's last row (one per text character, in hbuf[0..n)) enter the next panel's first row where the first
// panel takes the constant +1 of the global distance.  hbuf == nullptr: no scratch, such a pattern returns kEdTooLong.
constexpr uint32_t kPanelRows = 32u * kCoopK * 64u;

// One panel: pattern rows pat[0..m), m <= kPanelRows.  kFirst: the horizontal delta entering row 0 is the constant +1, else
// hbuf[j]; kLast: the distance is tracked at the panel's last row (score), else the deltas leaving it go to hbuf[j].
template <bool kFirst, bool kLast>
__device__ __forceinline__ int ed_panel(const uint8_t* pat, uint32_t m, const uint8_t* txt, uint32_t n, uint64_t* peq, uint32_t lane,
                                        volatile int8_t* hbuf, int score, uint32_t& last_lane) {
    const uint32_t nblk = (m + 63) >> 6;
    const uint32_t K = (nblk + 31) >> 5;
    const uint32_t b0 = lane * K;                                  // first block of this lane
    last_lane = (nblk - 1) / K;
    // pattern masks
    for (uint32_t k = 0; k < K; k++) {
        uint64_t pA = 0, pC = 0, pG = 0, pT = 0;
        const uint32_t g = b0 + k;
        if (g < nblk) {
            const uint32_t lo = g << 6, hi = min(m, lo + 64u);
            for (uint32_t j = lo; j < hi; j++) {
                const uint32_t c = base_code(pat[j]);
                const uint64_t bit = 1ull << (j - lo);
                if (c == 0) pA |= bit; else if (c == 1) pC |= bit; else if (c == 2) pG |= bit; else if (c == 3) pT |= bit;
            }
        }
        peq[(k * 4 + 0) * 32 + lane] = pA; peq[(k * 4 + 1) * 32 + lane] = pC;
        peq[(k * 4 + 2) * 32 + lane] = pG; peq[(k * 4 + 3) * 32 + lane] = pT;
    }
    __syncwarp();
    uint64_t Pv[kCoopK], Mv[kCoopK];
#pragma unroll
    for (int k = 0; k < kCoopK; k++) { Pv[k] = ~0ull; Mv[k] = 0; }
    const uint32_t top_bit = (m - 1) & 63u;
    int hout_prev = 0;
    const uint32_t steps = n + last_lane;
    uint8_t ch_next = (lane == 0 && n > 0) ? txt[0] : 0;
    for (uint32_t t = 0; t < steps; t++) {
        int hin = __shfl_up_sync(HP_FULL_MASK, hout_prev, 1);
        const int64_t j = (int64_t)t - (int64_t)lane;
        const bool active = j >= 0 && j < (int64_t)n && lane <= last_lane;
        if (lane == 0) hin = (kFirst || !active) ? 1 : (int)hbuf[j];
        const uint8_t ch = ch_next;
        // the next step's character: t + 1 - lane
        {
            const int64_t jn = j + 1;
            ch_next = (jn >= 0 && jn < (int64_t)n && lane <= last_lane) ? txt[jn] : 0;
        }
        if (active) {
            const uint32_t c = base_code(ch);
#pragma unroll
            for (int k = 0; k < kCoopK; k++) {
                const uint32_t g = b0 + (uint32_t)k;
                if ((uint32_t)k < K && g < nblk) {
                    uint64_t Eq;
                    if (c < 4u) Eq = peq[((uint32_t)k * 4 + c) * 32 + lane];
                    else {
                        Eq = 0;
                        const uint32_t lo = g << 6, hi = min(m, lo + 64u);
                        for (uint32_t q = lo; q < hi; q++) if (pat[q] == ch) Eq |= 1ull << (q - lo);
                    }
                    const uint64_t Xv = Eq | Mv[k];
                    if (hin < 0) Eq |= 1ull;
                    const uint64_t Xh = (((Eq & Pv[k]) + Pv[k]) ^ Pv[k]) | Eq;
                    uint64_t Ph = Mv[k] | ~(Xh | Pv[k]);
                    uint64_t Mh = Pv[k] & Xh;
                    int hout = 0;
                    if (kLast && g == nblk - 1) {                   // the distance is tracked at the pattern's last row
                        if ((Ph >> top_bit) & 1ull) score++; else if ((Mh >> top_bit) & 1ull) score--;
                    }
                    if (Ph >> 63) hout = 1; else if (Mh >> 63) hout = -1;
                    Ph <<= 1; Mh <<= 1;
                    if (hin < 0) Mh |= 1ull; else if (hin > 0) Ph |= 1ull;
                    Pv[k] = Mh | ~(Xv | Ph);
                    Mv[k] = Ph & Xv;
                    hin = hout;
                }
            }
            hout_prev = hin;
            // a panel that is not the last one is full and ends on a block boundary: the delta leaving its last row feeds the
            // next panel (lane 0 read hbuf[j] of this panel last_lane steps ago, so the row is reused in place)
            if (!kLast && lane == last_lane) hbuf[j] = (int8_t)hin;
        }
    }
    __syncwarp();
    return score;
}

// patterns of several panels (rare: an SV allele and a read slice both beyond 16384 bases)
__device__ __noinline__ uint32_t ed_coop_long(const uint8_t* pat, uint32_t m, const uint8_t* txt, uint32_t n, uint64_t* peq,
                                              uint32_t lane, volatile int8_t* hbuf) {
    uint32_t last_lane = 0;
    int score = ed_panel<true, false>(pat, kPanelRows, txt, n, peq, lane, hbuf, (int)m, last_lane);
    uint32_t p0 = kPanelRows;
    for (; m - p0 > kPanelRows; p0 += kPanelRows)
        score = ed_panel<false, false>(pat + p0, kPanelRows, txt, n, peq, lane, hbuf, score, last_lane);
    score = ed_panel<false, true>(pat + p0, m - p0, txt, n, peq, lane, hbuf, score, last_lane);
    return (uint32_t)__shfl_sync(HP_FULL_MASK, score, last_lane);
}

__device__ __forceinline__ uint32_t ed_coop(const uint8_t* pat, uint32_t m, const uint8_t* txt, uint32_t n, uint64_t* peq, uint32_t lane,
                                            volatile int8_t* hbuf = nullptr) {
    if (m > kPanelRows) return hbuf ? ed_coop_long(pat, m, txt, n, peq, lane, hbuf) : kEdTooLong;
    uint32_t last_lane = 0;
    const int score = ed_panel<true, true>(pat, m, txt, n, peq, lane, nullptr, (int)m, last_lane);
    return (uint32_t)__shfl_sync(HP_FULL_MASK, score, last_lane);
}

// edit distance of (a, la) vs (b, lb); symmetric, so the shorter one is the bit-vector pattern.  Lane-local part:
// returns true and the distance when the shorter side has <= 64 bytes.
__device__ __forceinline__ bool ed_lane(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb, uint32_t& d) {
    const bool a_short = la <= lb;
    const uint8_t* pat = a_short ? a : b;
    const uint8_t* txt = a_short ? b : a;
    const uint32_t m = a_short ? la : lb, n = a_short ? lb : la;
    if (m > 64u) return false;
    d = ed_small(pat, m, txt, n);
    return true;
}

// Rust `x as u8` for a finite or NaN f64: saturating, NaN -> 0
__device__ __forceinline__ uint8_t f64_as_u8(double f) {
    if (!(f == f) || f <= 0.0) return 0;
    if (f >= 255.0) return 255;
    return (uint8_t)f;
}

__global__ void __launch_bounds__(kLocalWarps * 32) local_realign_kernel(LocalArgs a) {
    __shared__ uint64_t peq_s[kLocalWarps][kCoopK * 4 * 32];
    __shared__ uint32_t job_s[kLocalWarps];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint64_t* peq = peq_s[warp];
    volatile int8_t* hbuf = a.ed_scratch ? a.ed_scratch + (uint64_t)(blockIdx.x * kLocalWarps + warp) * a.ed_stride : nullptr;
    for (;;) {
        if (lane == 0) job_s[warp] = atomicAdd(a.ticket, 1u);
        __syncwarp();
        const uint32_t j = job_s[warp];
        __syncwarp();
        if (j >= a.n_jobs) break;
        Segs sg;
        const uint64_t s0 = a.seg_off[j];
        sg.ref = a.seg_ref + s0; sg.rd = a.seg_read + s0; sg.len = a.seg_len + s0; sg.n = (uint32_t)(a.seg_off[j + 1] - s0);
        const int64_t range_start = a.read_pos[j];                                                 // :133-150
        int64_t max_position = range_start;
        if (sg.n) max_position = max(max_position, sg.ref[sg.n - 1] + (int64_t)sg.len[sg.n - 1] - 1);
        const int64_t range_end = max_position + 1;
        const uint8_t* seq = a.read_bytes + a.read_off[j];
        const uint8_t* rq = a.read_quals + a.read_off[j];
        const uint64_t read_len = a.read_off[j + 1] - a.read_off[j];
        const uint32_t v_lo = a.var_lo[j], nv = a.var_hi[j] - v_lo;
        const uint64_t row = a.row_off[j];

        // ---------------- pass A ----------------
        for (uint32_t i = lane; i < nv; i += 32) {
            const uint32_t k = v_lo + i;
            const uint64_t cell = row + i;
            const int64_t vpos = a.position[k];
            const uint32_t vt = a.vtype[k];
            uint8_t allele = HP_ALLELE_NOOVERLAP, qual = 0;
            uint32_t cls = 0;
            if (a.ignored[k]) {
                // :179-185
            } else if (vt == HP_VT_SNV || vt == HP_VT_INSERTION || vt == HP_VT_DELETION || vt == HP_VT_INDEL ||
                       vt == HP_VT_SV_INSERTION || vt == HP_VT_TANDEM_REPEAT) {
                const int64_t pl = a.prefix_len[k], ql = a.postfix_len[k], rl = a.ref_len[k];
                const int64_t first_start = vpos - pl, last_start = vpos + 1;                      // :208-211
                const int64_t first_end = vpos + rl, last_end = vpos + rl + ql + 1;
                // closest start: the largest mapped coordinate in [first_start, last_start) (:214-220)
                bool has_cs = false, has_ce = false;
                int64_t closest_start = 0, closest_end = 0;
                {
                    const int s = sg.last_le(vpos);
                    if (s >= 0) {
                        const int64_t e = sg.ref[s] + (int64_t)sg.len[s] - 1;                      // last mapped coordinate of s
                        const int64_t c = e < vpos ? e : vpos;
                        if (c >= first_start) { has_cs = true; closest_start = (int64_t)sg.rd[s] + (c - sg.ref[s]); }
                    }
                }
                // closest end: the smallest mapped coordinate in [first_end, last_end) (:223-229)
                {
                    const int s = sg.last_le(first_end);
                    if (s >= 0 && first_end - sg.ref[s] < (int64_t)sg.len[s]) { has_ce = true; closest_end = (int64_t)sg.rd[s] + (first_end - sg.ref[s]); }
                    else if (s + 1 < (int)sg.n && sg.ref[s + 1] < last_end) { has_ce = true; closest_end = sg.rd[s + 1]; }
                }
                bool has_s = false, has_e = false;
                int64_t ss = 0, se = 0;
                uint32_t start_clip = 0, end_clip = 0;
                if (has_cs && has_ce) {                                                            // :237-270
                    for (int64_t sc = first_start; sc < last_start; sc++) {
                        start_clip++;
                        int64_t si;
                        if (sg.get(sc, si)) {
                            if (closest_start - si > 2 * pl) continue;
                            ss = si; has_s = true;
                            for (int64_t ec = last_end - 1; ec >= first_end; ec--) {
                                end_clip++;
                                int64_t ni;
                                if (sg.get(ec, ni)) {
                                    if (ni - closest_end > 2 * ql) continue;
                                    se = ni; has_e = true;
                                    break;
                                }
                            }
                            break;
                        }
                    }
                }
                if (has_s) {
                    cls = HP_LOCAL_OVERLAPS;
                    if (has_e) {
                        if (se < ss || (uint64_t)se > read_len) { allele = HP_ALLELE_AMBIGUOUS; cls |= kBadSlice; }
                        else {
                            const uint32_t n = (uint32_t)(se - ss);
                            const uint8_t* obs = seq + ss;
                            const uint8_t* a0 = a.allele_bytes + a.a0_off[k];
                            const uint8_t* a1 = a.allele_bytes + a.a1_off[k];
                            const uint32_t l0 = a.a0_len[k], l1 = a.a1_len[k];
                            allele = HP_ALLELE_AMBIGUOUS;                                          // match_allele (variants.rs:598-606)
                            if (n == l0) { uint32_t q = 0; while (q < n && obs[q] == a0[q]) q++; if (q == n) allele = HP_ALLELE_REFERENCE; }
                            if (allele == HP_ALLELE_AMBIGUOUS && n == l1) { uint32_t q = 0; while (q < n && obs[q] == a1[q]) q++; if (q == n) allele = HP_ALLELE_ALTERNATE; }
                            if (allele == HP_ALLELE_AMBIGUOUS) {
                                cls |= kPending;
                                a.t_ss[cell] = (uint32_t)ss; a.t_se[cell] = (uint32_t)se;
                                a.t_clip[cell] = start_clip - 1; a.t_clip2[cell] = end_clip - 1;
                            } else cls |= HP_LOCAL_EXACT;
                            // harmonic mean of the base qualities, summed in read order (:293-299)
                            double sum = 0.0;
                            for (uint32_t q = 0; q < n; q++) sum = __dadd_rn(sum, __ddiv_rn(1.0, (double)rq[ss + q]));
                            const double harmonic = __ddiv_rn((double)n, sum);
                            const double factor = fmin(__ddiv_rn(harmonic, 40.0), 1.0);
                            const double base = vt == HP_VT_SNV ? 80.0 : vt == HP_VT_TANDEM_REPEAT ? 40.0 : vt == HP_VT_SV_INSERTION ? 20.0 : 10.0;   // :18-21, :302-323
                            qual = f64_as_u8(fmax(__dmul_rn(base, factor), 1.0));                   // :327
                        }
                    } else allele = HP_ALLELE_AMBIGUOUS;                                           // :331-337
                } else if (vpos >= range_start && vpos < range_end) { allele = HP_ALLELE_AMBIGUOUS; cls = HP_LOCAL_OVERLAPS; }   // :340-343
            } else if (vt == HP_VT_SV_DELETION) {                                                  // :354-451
                if (vpos >= range_start && vpos < range_end) {
                    cls = HP_LOCAL_OVERLAPS;
                    allele = HP_ALLELE_AMBIGUOUS;
                    const int64_t last_start = vpos + 1, first_end = vpos + (int64_t)a.ref_len[k];
                    if (first_end >= range_start && first_end < range_end) {
                        const int64_t expected = first_end - last_start;
                        // start anchor: walk down from last_start to the first mapped coordinate, stopping at the
                        // range start (:370-379); last_start > range_start here
                        int64_t start_anchor = last_start;
                        if (!sg.mapped(last_start)) {
                            const int s = sg.last_le(last_start);
                            const int64_t pm = s >= 0 ? sg.ref[s] + (int64_t)sg.len[s] - 1 : range_start - 1;
                            start_anchor = pm > range_start ? pm : range_start;
                        }
                        // end anchor: walk up from first_end to the first mapped coordinate or the range end (:380-388)
                        int64_t end_anchor = first_end;
                        if (!sg.mapped(first_end)) {
                            const int s = sg.last_le(first_end);
                            const int64_t sm = (s + 1 < (int)sg.n) ? sg.ref[s + 1] : range_end;
                            end_anchor = sm < range_end ? sm : range_end;
                            if (end_anchor < first_end + 1) end_anchor = first_end + 1;
                        }
                        const int64_t deleted = (end_anchor - start_anchor) - sg.mapped_in(start_anchor, end_anchor);   // :391-396
                        const double ratio = __ddiv_rn((double)deleted, (double)expected);
                        if (ratio < 0.33) {
                            allele = HP_ALLELE_REFERENCE;
                            qual = f64_as_u8(fmax(__dmul_rn(20.0, __dsub_rn(1.0, ratio)), 1.0));
                            if (ratio == 0.0) cls |= HP_LOCAL_EXACT;
                        } else if (fabs(__dsub_rn(1.0, ratio)) < 0.33) {
                            allele = HP_ALLELE_ALTERNATE;
                            qual = f64_as_u8(fmax(__dmul_rn(20.0, __dsub_rn(1.0, fabs(__dsub_rn(1.0, ratio)))), 1.0));
                            if (ratio == 1.0) cls |= HP_LOCAL_EXACT;
                            cls |= kSvDelAlt;
                            a.del_end[cell] = first_end;                                           // :428
                        }
                    }
                }
            } else cls = kUnhandled;                                                               // :452-454 panic!
            a.alleles[cell] = allele; a.quals[cell] = qual; a.mclass[cell] = (uint8_t)cls;
            if (a.ed) { a.ed[2 * cell] = 0; a.ed[2 * cell + 1] = 0; }
        }
        __syncwarp();

        // ---------------- pass B: closest_allele_clip of the pending cells (variants.rs:624-641) ----------------
        for (uint32_t base = 0; base < nv; base += 32) {
            const uint32_t i = base + lane;
            const uint64_t cell = row + i;
            uint32_t cls = i < nv ? a.mclass[cell] : 0u;
            const bool pending = (cls & kPending) != 0;
            const uint8_t *obs = nullptr, *c0 = nullptr, *c1 = nullptr;
            uint32_t n = 0, l0 = 0, l1 = 0, d0 = 0, d1 = 0;
            bool need0 = false, need1 = false;
            if (pending) {
                const uint32_t k = v_lo + i;
                const uint32_t ss = a.t_ss[cell], se = a.t_se[cell], head = a.t_clip[cell], tail = a.t_clip2[cell];
                obs = seq + ss; n = se - ss;
                c0 = a.allele_bytes + a.a0_off[k] + head; l0 = a.a0_len[k] - tail - head;
                c1 = a.allele_bytes + a.a1_off[k] + head; l1 = a.a1_len[k] - tail - head;
                need0 = !ed_lane(obs, n, c0, l0, d0);
                need1 = !ed_lane(obs, n, c1, l1, d1);
            }
            // long x long comparisons, one at a time on the whole warp
#pragma unroll 1
            for (int which = 0; which < 2; which++) {
                uint32_t m = __ballot_sync(HP_FULL_MASK, which ? need1 : need0);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint8_t* xa = (const uint8_t*)__shfl_sync(HP_FULL_MASK, (unsigned long long)obs, src);
                    const uint8_t* xb = (const uint8_t*)__shfl_sync(HP_FULL_MASK, (unsigned long long)(which ? c1 : c0), src);
                    const uint32_t la = __shfl_sync(HP_FULL_MASK, n, src), lb = __shfl_sync(HP_FULL_MASK, which ? l1 : l0, src);
                    const bool a_short = la <= lb;
                    const uint32_t d = ed_coop(a_short ? xa : xb, a_short ? la : lb, a_short ? xb : xa, a_short ? lb : la, peq, lane, hbuf);
                    if ((int)lane == src) { if (which) d1 = d; else d0 = d; }
                }
            }
            if (pending) {
                uint8_t allele;
                if (d0 == kEdTooLong || d1 == kEdTooLong) { allele = HP_ALLELE_AMBIGUOUS; cls |= kTooLong; d0 = d1 = 0; }
                else allele = d0 < d1 ? HP_ALLELE_REFERENCE : (d0 > d1 ? HP_ALLELE_ALTERNATE : HP_ALLELE_AMBIGUOUS);
                a.alleles[cell] = allele;
                a.mclass[cell] = (uint8_t)(cls & ~kPending);
                if (a.ed) { a.ed[2 * cell] = d0; a.ed[2 * cell + 1] = d1; }
            }
        }
        __syncwarp();

        // ---------------- pass C: SV deletions called ALT mask the variants they cover (:186-193, :428) ----------------
        int64_t lde = 0;                                           // last_deletion_end
        uint32_t job_flags = 0;
        for (uint32_t base = 0; base < nv; base += 32) {
            const uint32_t i = base + lane;
            const uint64_t cell = row + i;
            const bool in = i < nv;
            const uint32_t k = v_lo + (in ? i : 0);
            uint32_t cls = in ? a.mclass[cell] : 0u;
            const bool ign = in && a.ignored[k] != 0;
            const int64_t vpos = in ? a.position[k] : 0;
            const int64_t dend = (cls & kSvDelAlt) ? a.del_end[cell] : 0;
            int64_t my_lde = lde;                                  // last_deletion_end as seen by this variant
            bool cand = in && !ign && (cls & kSvDelAlt);
            for (;;) {
                const uint32_t m = __ballot_sync(HP_FULL_MASK, cand && vpos >= my_lde);
                if (m == 0) break;
                const int f = __ffs(m) - 1;
                const int64_t nl = __shfl_sync(HP_FULL_MASK, dend, f);
                if ((int)lane == f) cand = false;                  // this deletion stands and sets last_deletion_end
                if ((int)lane > f) my_lde = nl;
                lde = nl;
            }
            if (in && !ign) {
                if (vpos < my_lde) {                               // :186-193
                    a.alleles[cell] = HP_ALLELE_AMBIGUOUS; a.quals[cell] = 0;
                    if (a.ed) { a.ed[2 * cell] = 0; a.ed[2 * cell + 1] = 0; }
                    cls = HP_LOCAL_OVERLAPS;
                }
                job_flags |= cls & (kUnhandled | kBadSlice | kTooLong);
            }
            if (in) a.mclass[cell] = (uint8_t)(cls & (HP_LOCAL_OVERLAPS | HP_LOCAL_EXACT));
        }
        job_flags = __reduce_or_sync(HP_FULL_MASK, job_flags);
        if (lane == 0)
            a.status[j] = (job_flags & kUnhandled) ? HP_LOCAL_UNHANDLED_TYPE : (job_flags & kBadSlice) ? HP_LOCAL_BAD_SLICE
                        : (job_flags & kTooLong) ? HP_LOCAL_ALLELE_TOO_LONG : HP_LOCAL_OK;
        __syncwarp();
    }
}

// sequence_alignment::edit_distance for a batch of pairs: one warp per pair
struct EdArgs {
    uint32_t n_pairs;
    const uint8_t* bytes;
    const uint64_t* a_off;
    const uint32_t* a_len;
    const uint64_t* b_off;
    const uint32_t* b_len;
    uint32_t* dist;
    int8_t* ed_scratch;
    uint64_t ed_stride;
};

__global__ void __launch_bounds__(kLocalWarps * 32) edit_distance_kernel(EdArgs a) {
    __shared__ uint64_t peq_s[kLocalWarps][kCoopK * 4 * 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * kLocalWarps + warp, nw = gridDim.x * kLocalWarps;
    for (uint32_t p = gw; p < a.n_pairs; p += nw) {
        const uint8_t* x = a.bytes + a.a_off[p];
        const uint8_t* y = a.bytes + a.b_off[p];
        const uint32_t lx = a.a_len[p], ly = a.b_len[p];
        const bool x_short = lx <= ly;
        const uint8_t* pat = x_short ? x : y;
        const uint8_t* txt = x_short ? y : x;
        const uint32_t m = x_short ? lx : ly, n = x_short ? ly : lx;
        uint32_t d;
        if (m <= 64u) { d = 0; if (lane == 0) d = ed_small(pat, m, txt, n); d = __shfl_sync(HP_FULL_MASK, d, 0); }
        else d = ed_coop(pat, m, txt, n, peq_s[warp], lane, a.ed_scratch ? a.ed_scratch + (uint64_t)gw * a.ed_stride : nullptr);
        if (lane == 0) a.dist[p] = d;
        __syncwarp();
    }
}

}  // namespace hp

using namespace hp;

namespace {
inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
}

extern "C" int hp_local_realign_batch(hp_ctx* ctx, const hp_local_batch* b, hp_local_out* out) {
    if (!ctx || !b || !out || !out->alleles || !out->quals || !out->status) return HP_ERR_INVALID_INPUT;
    if (b->n_jobs == 0) return HP_OK;
    auto fail = [&](int code, const std::string& msg) { ctx->err = msg; return code; };
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(HP_ERR_CUDA, "cudaSetDevice failed");
    const hp_variant_table& t = b->variants;
    const uint32_t nj = b->n_jobs, nvt = t.n_variants;
    const uint64_t n_segs = b->seg_off[nj], n_read = b->read_off[nj], n_cells = b->row_off[nj];
    // ---- validation (conditions the reference answers with a panic or that would read out of bounds) ----
    for (uint32_t k = 0; k < nvt; k++) {
        if (t.allele0_off[k] + t.allele0_len[k] > t.n_allele_bytes || t.allele1_off[k] + t.allele1_len[k] > t.n_allele_bytes)
            return fail(HP_ERR_INVALID_INPUT, "allele bytes out of range, variant " + std::to_string(k));
        const uint64_t fix = (uint64_t)b->prefix_len[k] + b->postfix_len[k];
        if (fix > t.allele0_len[k] || fix > t.allele1_len[k])
            return fail(HP_ERR_INVALID_INPUT, "prefix + postfix longer than an allele, variant " + std::to_string(k));
        if (t.position[k] < (int64_t)b->prefix_len[k]) return fail(HP_ERR_INVALID_INPUT, "prefix reaches below coordinate 0, variant " + std::to_string(k));
    }
    for (uint32_t j = 0; j < nj; j++) {
        if (b->var_lo[j] > b->var_hi[j] || b->var_hi[j] > nvt) return fail(HP_ERR_INVALID_INPUT, "variant range of job " + std::to_string(j));
        if (b->row_off[j + 1] - b->row_off[j] != (uint64_t)(b->var_hi[j] - b->var_lo[j])) return fail(HP_ERR_INVALID_INPUT, "row_off of job " + std::to_string(j));
        if (b->seg_off[j + 1] < b->seg_off[j] || b->read_off[j + 1] < b->read_off[j]) return fail(HP_ERR_INVALID_INPUT, "offsets of job " + std::to_string(j));
        const uint64_t rl = b->read_off[j + 1] - b->read_off[j];
        int64_t prev_ref = INT64_MIN; uint64_t prev_rd = 0;
        for (uint64_t s = b->seg_off[j]; s < b->seg_off[j + 1]; s++) {
            if (b->seg_len[s] == 0 || b->seg_ref_start[s] < prev_ref || b->seg_read_start[s] < prev_rd || b->seg_ref_start[s] < b->read_pos[j] ||
                (uint64_t)b->seg_read_start[s] + b->seg_len[s] > rl)
                return fail(HP_ERR_INVALID_INPUT, "aligned segments of job " + std::to_string(j) + " are not ascending / inside the read");
            prev_ref = b->seg_ref_start[s] + b->seg_len[s]; prev_rd = (uint64_t)b->seg_read_start[s] + b->seg_len[s];
        }
    }
    const size_t in_bytes = al256(8 * (size_t)nvt) + al256(4 * (size_t)nvt) * 5 + al256(8 * (size_t)nvt) * 2 + al256(nvt) * 2 + al256(t.n_allele_bytes) +
                            al256(4 * (size_t)nj) * 2 + al256(8 * (size_t)nj) + al256(8 * ((size_t)nj + 1)) * 3 +
                            al256(8 * n_segs) + al256(4 * n_segs) * 2 + al256(n_read) * 2 + 4096;
    const size_t out_bytes = al256(n_cells) * 3 + al256(8 * n_cells) + al256(4 * (size_t)nj) + al256(4 * n_cells) * 4 + al256(8 * n_cells) + 4096;
    if (!ctx->stage_in.reserve(in_bytes) || !ctx->stage_out.reserve(out_bytes) || !ctx->ticket.reserve(256))
        return fail(HP_ERR_OUT_OF_MEMORY, "staging allocation failed");
    cudaStream_t st = ctx->stream;
    uint8_t* p = (uint8_t*)ctx->stage_in.ptr;
    bool ok = true;
    auto up = [&](const void* src, size_t bytes) { uint8_t* d = p; p += al256(bytes); if (bytes) ok &= cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess; return d; };
    LocalArgs a;
    a.n_jobs = nj;
    a.position = (const int64_t*)up(t.position, 8 * (size_t)nvt);
    a.ref_len = (const uint32_t*)up(t.ref_len, 4 * (size_t)nvt);
    a.prefix_len = (const uint32_t*)up(b->prefix_len, 4 * (size_t)nvt);
    a.postfix_len = (const uint32_t*)up(b->postfix_len, 4 * (size_t)nvt);
    a.a0_off = (const uint64_t*)up(t.allele0_off, 8 * (size_t)nvt); a.a0_len = (const uint32_t*)up(t.allele0_len, 4 * (size_t)nvt);
    a.a1_off = (const uint64_t*)up(t.allele1_off, 8 * (size_t)nvt); a.a1_len = (const uint32_t*)up(t.allele1_len, 4 * (size_t)nvt);
    a.vtype = up(t.vtype, nvt); a.ignored = up(t.ignored, nvt);
    a.allele_bytes = up(t.allele_bytes, t.n_allele_bytes);
    a.var_lo = (const uint32_t*)up(b->var_lo, 4 * (size_t)nj); a.var_hi = (const uint32_t*)up(b->var_hi, 4 * (size_t)nj);
    a.read_pos = (const int64_t*)up(b->read_pos, 8 * (size_t)nj);
    a.seg_off = (const uint64_t*)up(b->seg_off, 8 * ((size_t)nj + 1));
    a.seg_ref = (const int64_t*)up(b->seg_ref_start, 8 * n_segs);
    a.seg_read = (const uint32_t*)up(b->seg_read_start, 4 * n_segs);
    a.seg_len = (const uint32_t*)up(b->seg_len, 4 * n_segs);
    // bulk arrays: pageable sources go through pinned staging (hp::upload_large)
    { uint8_t* d = p; p += al256(n_read); ok &= hp::upload_large(ctx->pin_reads, d, b->read_bytes, n_read, st); a.read_bytes = d; }
    { uint8_t* d = p; p += al256(n_read); ok &= hp::upload_large(ctx->pin_quals, d, b->read_quals, n_read, st); a.read_quals = d; }
    a.read_off = (const uint64_t*)up(b->read_off, 8 * ((size_t)nj + 1));
    a.row_off = (const uint64_t*)up(b->row_off, 8 * ((size_t)nj + 1));
    uint8_t* q = (uint8_t*)ctx->stage_out.ptr;
    auto carve = [&](size_t bytes) { uint8_t* d = q; q += al256(bytes); return d; };
    a.alleles = carve(n_cells); a.quals = carve(n_cells); a.mclass = carve(n_cells);
    a.ed = out->edit_distance ? (uint32_t*)carve(8 * n_cells) : nullptr;
    a.status = (int32_t*)carve(4 * (size_t)nj);
    a.t_ss = (uint32_t*)carve(4 * n_cells); a.t_se = (uint32_t*)carve(4 * n_cells);
    a.t_clip = (uint32_t*)carve(4 * n_cells); a.t_clip2 = (uint32_t*)carve(4 * n_cells);
    a.del_end = (int64_t*)carve(8 * n_cells);
    a.ticket = (uint32_t*)ctx->ticket.ptr;
    ok &= cudaMemsetAsync(a.ticket, 0, 4, st) == cudaSuccess;
    int grid = (int)std::min<uint64_t>(((uint64_t)nj + kLocalWarps - 1) / kLocalWarps, (uint64_t)ctx->sm_count * 8);
    // comparisons with more than one panel of pattern rows (both the read slice and an allele beyond 16384 bases) carry a row of
    // horizontal deltas per warp from panel to panel
    a.ed_scratch = nullptr; a.ed_stride = 0;
    {
        uint64_t max_allele = 0, max_read = 0;
        for (uint32_t k = 0; k < nvt; k++) max_allele = std::max<uint64_t>(max_allele, std::max(t.allele0_len[k], t.allele1_len[k]));
        for (uint32_t j = 0; j < nj; j++) max_read = std::max<uint64_t>(max_read, b->read_off[j + 1] - b->read_off[j]);
        if (max_allele > kPanelRows && max_read > kPanelRows) {
            grid = std::min(grid, ctx->sm_count * 2);
            a.ed_stride = al256(std::max(max_allele, max_read));
            if (!ctx->ed_scratch.reserve((size_t)grid * kLocalWarps * a.ed_stride)) return fail(HP_ERR_OUT_OF_MEMORY, "edit distance scratch allocation failed");
            a.ed_scratch = (int8_t*)ctx->ed_scratch.ptr;
        }
    }
    if (ctx->ev0) { cudaEventRecord(ctx->ev0, st); }
    local_realign_kernel<<<grid, kLocalWarps * 32, 0, st>>>(a);
    if (ctx->ev1) { cudaEventRecord(ctx->ev1, st); ctx->timing_pending = true; }
    ok &= cudaGetLastError() == cudaSuccess;
    ctx->launches++;
    if (n_cells) {
        ok &= cudaMemcpyAsync(out->alleles, a.alleles, n_cells, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        ok &= cudaMemcpyAsync(out->quals, a.quals, n_cells, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        if (out->match_class) ok &= cudaMemcpyAsync(out->match_class, a.mclass, n_cells, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        if (out->edit_distance) ok &= cudaMemcpyAsync(out->edit_distance, a.ed, 8 * n_cells, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    }
    ok &= cudaMemcpyAsync(out->status, a.status, 4 * (size_t)nj, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { cudaGetLastError(); return fail(HP_ERR_CUDA, "local realignment launch or copy failed"); }
    return HP_OK;
}

extern "C" int hp_edit_distance_batch(hp_ctx* ctx, uint32_t n_pairs, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* a_off,
                                      const uint32_t* a_len, const uint64_t* b_off, const uint32_t* b_len, uint32_t* dist) {
    if (!ctx || (n_pairs && (!a_off || !a_len || !b_off || !b_len || !dist))) return HP_ERR_INVALID_INPUT;
    if (n_pairs == 0) return HP_OK;
    auto fail = [&](int code, const std::string& msg) { ctx->err = msg; return code; };
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(HP_ERR_CUDA, "cudaSetDevice failed");
    for (uint32_t i = 0; i < n_pairs; i++)
        if (a_off[i] + a_len[i] > n_bytes || b_off[i] + b_len[i] > n_bytes) return fail(HP_ERR_INVALID_INPUT, "pair " + std::to_string(i) + " out of range");
    const size_t in_bytes = al256(n_bytes) + al256(8 * (size_t)n_pairs) * 2 + al256(4 * (size_t)n_pairs) * 2 + 4096;
    if (!ctx->stage_in.reserve(in_bytes) || !ctx->stage_out.reserve(al256(4 * (size_t)n_pairs) + 4096)) return fail(HP_ERR_OUT_OF_MEMORY, "staging allocation failed");
    cudaStream_t st = ctx->stream;
    uint8_t* p = (uint8_t*)ctx->stage_in.ptr;
    bool ok = true;
    auto up = [&](const void* src, size_t n) { uint8_t* d = p; p += al256(n); if (n) ok &= cudaMemcpyAsync(d, src, n, cudaMemcpyHostToDevice, st) == cudaSuccess; return d; };
    EdArgs a;
    a.n_pairs = n_pairs;
    a.bytes = up(bytes, n_bytes);
    a.a_off = (const uint64_t*)up(a_off, 8 * (size_t)n_pairs); a.a_len = (const uint32_t*)up(a_len, 4 * (size_t)n_pairs);
    a.b_off = (const uint64_t*)up(b_off, 8 * (size_t)n_pairs); a.b_len = (const uint32_t*)up(b_len, 4 * (size_t)n_pairs);
    a.dist = (uint32_t*)ctx->stage_out.ptr;
    int grid = (int)std::min<uint64_t>(((uint64_t)n_pairs + kLocalWarps - 1) / kLocalWarps, (uint64_t)ctx->sm_count * 8);
    a.ed_scratch = nullptr; a.ed_stride = 0;
    {
        uint64_t longest = 0;                                            // text length of the pairs whose pattern spans several panels
        for (uint32_t i = 0; i < n_pairs; i++)
            if (std::min(a_len[i], b_len[i]) > kPanelRows) longest = std::max<uint64_t>(longest, std::max(a_len[i], b_len[i]));
        if (longest) {
            grid = std::min(grid, ctx->sm_count * 2);
            a.ed_stride = al256(longest);
            if (!ctx->ed_scratch.reserve((size_t)grid * kLocalWarps * a.ed_stride)) return fail(HP_ERR_OUT_OF_MEMORY, "edit distance scratch allocation failed");
            a.ed_scratch = (int8_t*)ctx->ed_scratch.ptr;
        }
    }
    edit_distance_kernel<<<grid, kLocalWarps * 32, 0, st>>>(a);
    ok &= cudaGetLastError() == cudaSuccess;
    ctx->launches++;
    ok &= cudaMemcpyAsync(dist, a.dist, 4 * (size_t)n_pairs, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok &= cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { cudaGetLastError(); return fail(HP_ERR_CUDA, "edit distance launch or copy failed"); }
    return HP_OK;
}
