// hp_pack.cu -- packed phase-block container: the wire / on-disk form of hp_block_batch, so that a
// front end that still owns VCF / BAM decoding (the unmodified HiPhase Rust code, src/phaser.rs:27-323 +
// src/read_parsing.rs:520-637, or any other reader) can hand whole batches of ready-to-solve blocks to this library,
// and a stats writer with the solver-side columns of HiPhase's --stats-file (src/writers/phase_stats.rs:207-254).
//
// Layout (little endian):
//   0   char  magic[8]  "HPB200\0\1"
//   8   u32   kind      1 = phase blocks (hp_block_batch)
//   12  u32   n_sections
//   16  u64   n_blocks
//   24  section table: n_sections x { char name[16]; u32 elem_size; u32 reserved; u64 count; u64 offset }
//   ... section data, each 64-byte aligned (so a reader may mmap the file and point hp_block_batch into it)
// Sections of kind 1: var_off, read_off, read_start, read_end, cell_off, alleles, quals, ignored, is_snv and, optionally,
// var_pos (Variant::position() per variant: needed for the block tags of hp_post_solve_batch and the stats writer).
#include <cinttypes>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hiphase_b200.h"

namespace {

constexpr char kMagic[8] = {'H', 'P', 'B', '2', '0', '0', '\0', '\1'};

struct Section {
    char name[16];
    uint32_t elem_size;
    uint32_t reserved;
    uint64_t count;
    uint64_t offset;
};
static_assert(sizeof(Section) == 40, "section table entry is 40 bytes");

struct Header {
    char magic[8];
    uint32_t kind;
    uint32_t n_sections;
    uint64_t n_blocks;
};
static_assert(sizeof(Header) == 24, "header is 24 bytes");

inline uint64_t align64(uint64_t x) { return (x + 63) & ~63ull; }

}  // namespace

struct hp_packed {
    std::vector<uint8_t> bytes;
    hp_block_batch batch{};
    const int64_t* var_pos = nullptr;
    std::string err;
};

static thread_local std::string g_pack_err;

extern "C" const char* hp_pack_last_error(void) { return g_pack_err.c_str(); }

extern "C" int hp_pack_write_blocks(const char* path, const hp_block_batch* b, const int64_t* var_pos) {
    if (!path || !b || (b->n_blocks && (!b->var_off || !b->read_off))) { g_pack_err = "null argument"; return HP_ERR_INVALID_INPUT; }
    const uint32_t nb = b->n_blocks;
    const uint64_t nv = nb ? b->var_off[nb] : 0, nr = nb ? b->read_off[nb] : 0, nc = nr ? b->cell_off[nr] : 0;
    struct Src { const char* name; uint32_t es; uint64_t count; const void* data; };
    const uint64_t zero64 = 0;
    std::vector<Src> src = {
        {"var_off", 8, (uint64_t)nb + 1, nb ? (const void*)b->var_off : &zero64}, {"read_off", 8, (uint64_t)nb + 1, nb ? (const void*)b->read_off : &zero64},
        {"read_start", 4, nr, b->read_start}, {"read_end", 4, nr, b->read_end},
        {"cell_off", 8, nr + 1, nr ? (const void*)b->cell_off : &zero64},
        {"alleles", 1, nc, b->alleles}, {"quals", 1, nc, b->quals}, {"ignored", 1, nv, b->ignored}, {"is_snv", 1, nv, b->is_snv}};
    if (var_pos) src.push_back({"var_pos", 8, nv, var_pos});
    Header h;
    memcpy(h.magic, kMagic, 8);
    h.kind = 1; h.n_sections = (uint32_t)src.size(); h.n_blocks = nb;
    std::vector<Section> tab(src.size());
    uint64_t off = align64(sizeof(Header) + sizeof(Section) * src.size());
    for (size_t i = 0; i < src.size(); i++) {
        memset(&tab[i], 0, sizeof(Section));
        strncpy(tab[i].name, src[i].name, 15);
        tab[i].elem_size = src[i].es; tab[i].count = src[i].count; tab[i].offset = off;
        off = align64(off + src[i].es * src[i].count);
    }
    FILE* f = fopen(path, "wb");
    if (!f) { g_pack_err = std::string("cannot open ") + path; return HP_ERR_INVALID_INPUT; }
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(tab.data(), sizeof(Section), tab.size(), f) == tab.size();
    uint64_t pos = sizeof(Header) + sizeof(Section) * src.size();
    static const uint8_t pad[64] = {0};
    for (size_t i = 0; ok && i < src.size(); i++) {
        if (tab[i].offset > pos) { ok = fwrite(pad, 1, tab[i].offset - pos, f) == tab[i].offset - pos; pos = tab[i].offset; }
        const uint64_t n = (uint64_t)src[i].es * src[i].count;
        if (ok && n) { ok = src[i].data && fwrite(src[i].data, 1, n, f) == n; pos += n; }
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { g_pack_err = std::string("write failed: ") + path; return HP_ERR_INTERNAL; }
    return HP_OK;
}

extern "C" int hp_pack_open(const char* path, hp_packed** out) {
    if (!path || !out) { g_pack_err = "null argument"; return HP_ERR_INVALID_INPUT; }
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) { g_pack_err = std::string("cannot open ") + path; return HP_ERR_INVALID_INPUT; }
    hp_packed* p = new hp_packed();
    auto fail = [&](const std::string& m) { g_pack_err = m; fclose(f); delete p; return HP_ERR_INVALID_INPUT; };
    if (fseek(f, 0, SEEK_END) != 0) return fail("seek failed");
    const long sz = ftell(f);
    if (sz < (long)sizeof(Header)) return fail("file too short");
    rewind(f);
    p->bytes.resize((size_t)sz);
    if (fread(p->bytes.data(), 1, (size_t)sz, f) != (size_t)sz) return fail("read failed");
    Header h;
    memcpy(&h, p->bytes.data(), sizeof(h));
    if (memcmp(h.magic, kMagic, 8) != 0) return fail("bad magic (not an HPB200 v1 file)");
    if (h.kind != 1) return fail("unsupported container kind");
    if (h.n_sections > 64 || sizeof(Header) + sizeof(Section) * (uint64_t)h.n_sections > (uint64_t)sz) return fail("bad section table");
    if (h.n_blocks > 0xffffffffull) return fail("too many blocks");
    const Section* tab = (const Section*)(p->bytes.data() + sizeof(Header));
    auto find = [&](const char* name, uint32_t es, uint64_t want, bool required, const void** ptr) -> bool {
        for (uint32_t i = 0; i < h.n_sections; i++) {
            if (strncmp(tab[i].name, name, 16) != 0) continue;
            // (division, not es * want: a corrupt count must not wrap the product)
            if (tab[i].elem_size != es || tab[i].count != want || (tab[i].offset & 63) || tab[i].offset > (uint64_t)sz ||
                want > ((uint64_t)sz - tab[i].offset) / es) return false;
            *ptr = p->bytes.data() + tab[i].offset;
            return true;
        }
        *ptr = nullptr;
        return !required;
    };
    const uint32_t nb = (uint32_t)h.n_blocks;
    const void *vo, *ro, *rs, *re, *co, *al, *ql, *ig, *sn, *vp;
    if (!find("var_off", 8, (uint64_t)nb + 1, true, &vo) || !find("read_off", 8, (uint64_t)nb + 1, true, &ro)) return fail("var_off / read_off section");
    const uint64_t* var_off = (const uint64_t*)vo;
    const uint64_t* read_off = (const uint64_t*)ro;
    if (var_off[0] != 0 || read_off[0] != 0) return fail("offsets must start at 0");
    for (uint32_t i = 0; i < nb; i++) if (var_off[i + 1] < var_off[i] || read_off[i + 1] < read_off[i]) return fail("offsets must be non-decreasing");
    const uint64_t nv = var_off[nb], nr = read_off[nb];
    if (nv > (uint64_t)sz || nr > (uint64_t)sz / 4) return fail("variant / read counts exceed the file size");
    if (!find("read_start", 4, nr, true, &rs) || !find("read_end", 4, nr, true, &re) || !find("cell_off", 8, nr + 1, true, &co)) return fail("read sections");
    const uint64_t* cell_off = (const uint64_t*)co;
    if (cell_off[0] != 0) return fail("cell_off must start at 0");
    const uint32_t* rstart = (const uint32_t*)rs;
    const uint32_t* rend = (const uint32_t*)re;
    for (uint32_t i = 0; i < nb; i++) {
        const uint64_t n_var = var_off[i + 1] - var_off[i];
        for (uint64_t r = read_off[i]; r < read_off[i + 1]; r++)
            if (rend[r] < rstart[r] || rend[r] > n_var || cell_off[r + 1] < cell_off[r] || cell_off[r + 1] - cell_off[r] != (uint64_t)(rend[r] - rstart[r]))
                return fail("read " + std::to_string(r) + ": region outside its block, or region and cell range disagree");
    }
    const uint64_t nc = cell_off[nr];
    if (!find("alleles", 1, nc, true, &al) || !find("quals", 1, nc, true, &ql) || !find("ignored", 1, nv, true, &ig) || !find("is_snv", 1, nv, true, &sn)) return fail("cell / variant sections");
    if (!find("var_pos", 8, nv, false, &vp)) return fail("var_pos section");
    fclose(f);
    p->batch.n_blocks = nb;
    p->batch.var_off = var_off; p->batch.read_off = read_off; p->batch.read_start = rstart; p->batch.read_end = rend; p->batch.cell_off = cell_off;
    p->batch.alleles = (const uint8_t*)al; p->batch.quals = (const uint8_t*)ql; p->batch.ignored = (const uint8_t*)ig; p->batch.is_snv = (const uint8_t*)sn;
    p->var_pos = (const int64_t*)vp;
    *out = p;
    return HP_OK;
}

extern "C" int hp_pack_get_blocks(const hp_packed* p, hp_block_batch* batch, const int64_t** var_pos) {
    if (!p || !batch) return HP_ERR_INVALID_INPUT;
    *batch = p->batch;
    if (var_pos) *var_pos = p->var_pos;
    return HP_OK;
}

extern "C" void hp_pack_close(hp_packed* p) { delete p; }

// One row per block with the solver-side columns of HiPhase's stats file (writers/phase_stats.rs:207-254; tab separated,
// or comma separated when the path ends in .csv like StatsWriter::new, :262-271).  start / end are the positions of the
// block's first / last variant when var_pos is given, else the variant index range.
extern "C" int hp_write_phase_stats(const char* path, const hp_block_batch* b, const int64_t* var_pos, const hp_astar_out* out,
                                    uint64_t first_block_index) {
    if (!path || !b || !out || !out->stats || !out->status) { g_pack_err = "null argument"; return HP_ERR_INVALID_INPUT; }
    FILE* f = fopen(path, "w");
    if (!f) { g_pack_err = std::string("cannot open ") + path; return HP_ERR_INVALID_INPUT; }
    const size_t L = strlen(path);
    const char d = (L >= 4 && strcmp(path + L - 4, ".csv") == 0) ? ',' : '\t';
    fprintf(f, "block_index%cstart%cend%cnum_variants%cnum_reads%cpruned_solutions%cestimated_cost%cactual_cost%ccost_ratio%cphased_variants%chomozygous_variants%cskipped_variants%csolver_status\n",
            d, d, d, d, d, d, d, d, d, d, d, d);
    for (uint32_t i = 0; i < b->n_blocks; i++) {
        const uint64_t v0 = b->var_off[i], v1 = b->var_off[i + 1];
        const hp_phase_stats& s = out->stats[i];
        const long long start = v1 > v0 ? (var_pos ? (long long)var_pos[v0] : (long long)v0) : 0;
        const long long end = v1 > v0 ? (var_pos ? (long long)var_pos[v1 - 1] : (long long)(v1 - 1)) : 0;
        const double ratio = s.actual_cost == 0 ? 1.0 : (double)s.estimated_cost / (double)s.actual_cost;   // get_cost_ratio, :184-198
        fprintf(f, "%" PRIu64 "%c%lld%c%lld%c%" PRIu64 "%c%" PRIu64 "%c%" PRIu64 "%c%" PRIu64 "%c%" PRIu64 "%c%.17g%c%" PRIu64 "%c%" PRIu64 "%c%" PRIu64 "%c%d\n",
                first_block_index + i, d, start, d, end, d, v1 - v0, d, b->read_off[i + 1] - b->read_off[i], d, s.pruned_solutions, d,
                s.estimated_cost, d, s.actual_cost, d, ratio, d, s.phased_variants, d, s.homozygous_variants, d, s.skipped_variants, d, out->status[i]);
    }
    if (fclose(f) != 0) { g_pack_err = std::string("write failed: ") + path; return HP_ERR_INTERNAL; }
    return HP_OK;
}
