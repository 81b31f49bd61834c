// hp_synth.cpp -- multi-threaded generator of the synthetic HG002-scale phase-block stream.
//
// Bench / test infrastructure (libhp_synth.so), NOT part of the product ABI: it only produces inputs in the layout of
// hp_block_batch (the reference's clipped ReadSegments, src/data_types/read_segments.rs:40-62).  Block b of the stream
// depends on (config_id, b) alone -- never on the thread count, the shard or the rank -- so every rank of a multi-GPU
// run can regenerate exactly its own blocks, and the CPU arm sees the very same blocks.
//
// Distribution (the same as hiphase_b200/synth.py: c3_blocks, with its own counter-seeded generator instead of numpy's):
//   N log-uniform in [n_lo, n_hi]; a block is "noisy" with p_noisy (p_err_noisy instead of p_err); R = max(2, coverage*N/12)
//   reads with span = clip(rint(Normal(12, 4)), 2, 40), uniform start, random haplotype; every variant left uncovered gets
//   a 3-variant read; cell = truth ^ hap, flipped with p_err, Ambiguous with p_amb, NoOverlap with p_gap, NoOverlap on
//   ignored variants (p_ignored; astar_phaser.rs:435-442); quality by variant type {SNV 160, indel 20, TR 80, SV 40}
//   (read_parsing.rs:815-835, doubled), 0 for non-binary cells; rows clipped as ReadSegment::new, reads with fewer than
//   two set alleles dropped (--min-matched-alleles 2, read_parsing.rs:617).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/hiphase_b200.h"

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) {
        uint64_t z = (x += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {                       // xoshiro256**
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

struct SynthParams {
    uint64_t config_id;
    uint32_t n_lo, n_hi;
    double coverage, p_noisy, p_err, p_err_noisy, p_amb, p_gap, p_ignored;
};

inline uint64_t block_seed(uint64_t config_id, uint64_t block) { return 0xB200ull + config_id * 1000003ull + block * 0x9e3779b97f4a7c15ull; }

struct Header { uint32_t n_var; uint8_t noisy; };

inline Header draw_header(Rng& g, const SynthParams& p) {
    Header h;
    const double ln = std::log((double)p.n_lo) + g.uni() * (std::log((double)p.n_hi) - std::log((double)p.n_lo));
    h.n_var = (uint32_t)std::max<double>(1.0, std::floor(std::exp(ln)));
    h.noisy = g.uni() < p.p_noisy;
    return h;
}

struct BlockOut {
    uint32_t n_var = 0;
    std::vector<uint32_t> rs, re;
    std::vector<uint8_t> al, ql, ign, snv;
};

void gen_block(uint64_t block, const SynthParams& p, BlockOut& o) {
    Rng g(block_seed(p.config_id, block));
    const Header h = draw_header(g, p);
    const uint32_t N = h.n_var;
    const double p_err = h.noisy ? p.p_err_noisy : p.p_err;
    o.n_var = N;
    std::vector<uint8_t> truth(N), vq(N);
    o.ign.assign(N, 0); o.snv.assign(N, 0);
    bool all_ign = true;
    for (uint32_t i = 0; i < N; i++) {
        truth[i] = (uint8_t)(g.next() >> 63);
        const double u = g.uni();
        const int t = u < 0.80 ? 0 : (u < 0.95 ? 1 : (u < 0.98 ? 2 : 3));
        static const uint8_t kQ[4] = {160, 20, 80, 40};
        vq[i] = kQ[t]; o.snv[i] = t == 0;
        o.ign[i] = g.uni() < p.p_ignored;
        all_ign = all_ign && o.ign[i];
    }
    if (N >= 2 && all_ign) std::fill(o.ign.begin(), o.ign.end(), 0);

    uint32_t R = (uint32_t)std::max<double>(2.0, std::floor(p.coverage * N / 12.0));
    std::vector<uint32_t> st(R), sp(R);
    std::vector<uint8_t> hp(R);
    std::vector<int32_t> cover(N + 1, 0);
    for (uint32_t r = 0; r < R; r++) {
        double s = std::nearbyint(12.0 + 4.0 * g.normal());
        s = std::min(40.0, std::max(2.0, s));
        sp[r] = (uint32_t)std::min<double>(s, N);
        st[r] = (uint32_t)std::floor(g.uni() * (double)(N - sp[r] + 1));
        hp[r] = (uint8_t)(g.next() >> 63);
        cover[st[r]]++; cover[st[r] + sp[r]]--;
    }
    if (N >= 3) {
        int32_t run = 0;
        for (uint32_t i = 0; i < N; i++) {
            run += cover[i];
            if (run == 0) {
                const uint32_t s = (uint32_t)std::min<int64_t>(std::max<int64_t>((int64_t)i - 1, 0), (int64_t)N - 3);
                st.push_back(s); sp.push_back(3); hp.push_back((uint8_t)(g.next() >> 63));
            }
        }
        R = (uint32_t)st.size();
    }
    o.rs.clear(); o.re.clear(); o.al.clear(); o.ql.clear();
    uint8_t a[64], q[64];
    for (uint32_t r = 0; r < R; r++) {
        int lo = -1, hi = -1, nset = 0;
        for (uint32_t k = 0; k < sp[r]; k++) {
            const uint32_t pos = st[r] + k;
            uint8_t x = truth[pos] ^ hp[r];
            if (g.uni() < p_err) x ^= 1;
            if (g.uni() < p.p_amb) x = 2;
            if (g.uni() < p.p_gap) x = 3;
            if (o.ign[pos]) x = 3;
            a[k] = x; q[k] = x < 2 ? vq[pos] : 0;
            if (x < 2) { if (lo < 0) lo = (int)k; hi = (int)k; nset++; }
        }
        if (nset < 2) continue;
        o.rs.push_back(st[r] + (uint32_t)lo); o.re.push_back(st[r] + (uint32_t)hi + 1);
        o.al.insert(o.al.end(), a + lo, a + hi + 1);
        o.ql.insert(o.ql.end(), q + lo, q + hi + 1);
    }
}

}  // namespace

struct hp_synth_batch {
    std::vector<uint64_t> var_off, read_off, cell_off;
    std::vector<uint32_t> read_start, read_end;
    std::vector<uint8_t> alleles, quals, ignored, is_snv;
};

extern "C" {

struct hp_synth_params {
    uint64_t config_id;
    uint32_t n_lo, n_hi;
    double coverage, p_noisy, p_err, p_err_noisy, p_amb, p_gap, p_ignored;
};

void hp_synth_default_params(hp_synth_params* p) {
    p->config_id = 3; p->n_lo = 20; p->n_hi = 2000; p->coverage = 30.0; p->p_noisy = 0.02; p->p_err = 0.02;
    p->p_err_noisy = 0.15; p->p_amb = 0.03; p->p_gap = 0.01; p->p_ignored = 0.01;
}

static SynthParams conv(const hp_synth_params* p) {
    SynthParams s;
    s.config_id = p->config_id; s.n_lo = p->n_lo; s.n_hi = p->n_hi; s.coverage = p->coverage; s.p_noisy = p->p_noisy;
    s.p_err = p->p_err; s.p_err_noisy = p->p_err_noisy; s.p_amb = p->p_amb; s.p_gap = p->p_gap; s.p_ignored = p->p_ignored;
    return s;
}

// Variant count and noisy flag of blocks [first, first + n) without generating them (what a cost-sorted partition needs).
int hp_synth_headers(const hp_synth_params* params, uint64_t first, uint64_t n, uint32_t* n_var, uint8_t* noisy) {
    const SynthParams p = conv(params);
    for (uint64_t i = 0; i < n; i++) {
        Rng g(block_seed(p.config_id, first + i));
        const Header h = draw_header(g, p);
        if (n_var) n_var[i] = h.n_var;
        if (noisy) noisy[i] = h.noisy;
    }
    return 0;
}

// Generates the blocks ids[0..n) (in that order) on `threads` host threads.
int hp_synth_generate(const hp_synth_params* params, const uint64_t* ids, uint64_t n, int threads, hp_synth_batch** out) {
    const SynthParams p = conv(params);
    std::vector<BlockOut> blocks(n);
    if (threads < 1) threads = 1;
    std::atomic<uint64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const uint64_t i = next.fetch_add(16);
            if (i >= n) break;
            for (uint64_t k = i; k < std::min<uint64_t>(n, i + 16); k++) gen_block(ids[k], p, blocks[k]);
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(worker);
        worker();
        for (auto& t : th) t.join();
    }
    hp_synth_batch* b = new hp_synth_batch();
    b->var_off.resize(n + 1); b->read_off.resize(n + 1);
    uint64_t nv = 0, nr = 0, nc = 0;
    std::vector<uint64_t> cbase(n + 1);
    for (uint64_t i = 0; i < n; i++) {
        b->var_off[i] = nv; b->read_off[i] = nr; cbase[i] = nc;
        nv += blocks[i].n_var; nr += blocks[i].rs.size(); nc += blocks[i].al.size();
    }
    b->var_off[n] = nv; b->read_off[n] = nr; cbase[n] = nc;
    b->cell_off.resize(nr + 1); b->read_start.resize(nr); b->read_end.resize(nr);
    b->alleles.resize(nc); b->quals.resize(nc); b->ignored.resize(nv); b->is_snv.resize(nv);
    next = 0;
    auto copier = [&]() {
        for (;;) {
            const uint64_t i0 = next.fetch_add(64);
            if (i0 >= n) break;
            for (uint64_t i = i0; i < std::min<uint64_t>(n, i0 + 64); i++) {
                BlockOut& o = blocks[i];
                const uint64_t r0 = b->read_off[i];
                uint64_t c = cbase[i];
                for (size_t r = 0; r < o.rs.size(); r++) {
                    b->read_start[r0 + r] = o.rs[r]; b->read_end[r0 + r] = o.re[r];
                    b->cell_off[r0 + r] = c; c += o.re[r] - o.rs[r];
                }
                if (!o.al.empty()) { memcpy(&b->alleles[cbase[i]], o.al.data(), o.al.size()); memcpy(&b->quals[cbase[i]], o.ql.data(), o.ql.size()); }
                memcpy(&b->ignored[b->var_off[i]], o.ign.data(), o.n_var); memcpy(&b->is_snv[b->var_off[i]], o.snv.data(), o.n_var);
                std::vector<uint32_t>().swap(o.rs); std::vector<uint32_t>().swap(o.re); std::vector<uint8_t>().swap(o.al); std::vector<uint8_t>().swap(o.ql);
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; t++) th.emplace_back(copier);
        copier();
        for (auto& t : th) t.join();
    }
    b->cell_off[nr] = nc;
    *out = b;
    return 0;
}

// View of the generated batch (pointers into memory owned by *b).
void hp_synth_view(const hp_synth_batch* b, hp_block_batch* v) {
    v->n_blocks = (uint32_t)(b->var_off.size() - 1);
    v->var_off = b->var_off.data(); v->read_off = b->read_off.data(); v->read_start = b->read_start.data();
    v->read_end = b->read_end.data(); v->cell_off = b->cell_off.data(); v->alleles = b->alleles.data();
    v->quals = b->quals.data(); v->ignored = b->ignored.data(); v->is_snv = b->is_snv.data();
}

void hp_synth_free(hp_synth_batch* b) { delete b; }

}  // extern "C"
