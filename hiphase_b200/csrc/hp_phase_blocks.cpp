// hp_phase_blocks -- batch runner over packed phase blocks.
//   hp_phase_blocks <blocks.hpb> <out_prefix> [device]
// Reads an HPB200 container (hp_pack_open), solves every block on the GPU (hp_astar_solve_batch), runs the post-solve
// step when variant positions are present (hp_post_solve_batch), and writes
//   <out_prefix>.stats.tsv   one row per block, the solver-side columns of HiPhase's --stats-file
//   <out_prefix>.haps.tsv    block_index, variant index in block, position, haplotype_1, haplotype_2, phase block tag
// There is no CPU solver in here: without a B200 the run fails with the library's HP_ERR_NO_DEVICE.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/hiphase_b200.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <blocks.hpb> <out_prefix> [device]\n", argv[0]); return 2; }
    const int device = argc > 3 ? atoi(argv[3]) : 0;
    hp_packed* pk = nullptr;
    if (hp_pack_open(argv[1], &pk) != HP_OK) { fprintf(stderr, "error: %s\n", hp_pack_last_error()); return 1; }
    hp_block_batch b;
    const int64_t* var_pos = nullptr;
    hp_pack_get_blocks(pk, &b, &var_pos);
    const uint64_t nv = b.n_blocks ? b.var_off[b.n_blocks] : 0, nr = b.n_blocks ? b.read_off[b.n_blocks] : 0;
    hp_params prm;
    hp_default_params(&prm);
    hp_ctx* ctx = nullptr;
    if (hp_ctx_create(&prm, device, &ctx) != HP_OK) { fprintf(stderr, "error: %s\n", hp_last_error(nullptr)); hp_pack_close(pk); return 1; }
    std::vector<uint8_t> h1(nv + 1), h2(nv + 1);
    std::vector<hp_phase_stats> stats(b.n_blocks + 1);
    std::vector<int32_t> status(b.n_blocks + 1, -1);
    hp_astar_out out{};
    out.h1 = h1.data(); out.h2 = h2.data(); out.stats = stats.data(); out.status = status.data();
    int rc = hp_astar_solve_batch(ctx, &b, &out);
    if (rc != HP_OK) { fprintf(stderr, "error %d: %s\n", rc, hp_last_error(ctx)); return 1; }
    std::vector<uint64_t> tags(nv + 1, 0);
    if (var_pos && nv) {
        std::vector<uint32_t> span(nv + 1);
        std::vector<uint8_t> rhap(nr + 1);
        std::vector<uint64_t> rtag(nr + 1);
        hp_post_out po{span.data(), tags.data(), rhap.data(), rtag.data()};
        rc = hp_post_solve_batch(ctx, &b, var_pos, h1.data(), h2.data(), &po);
        if (rc != HP_OK) { fprintf(stderr, "error %d: %s\n", rc, hp_last_error(ctx)); return 1; }
    }
    const std::string prefix = argv[2];
    if (hp_write_phase_stats((prefix + ".stats.tsv").c_str(), &b, var_pos, &out, 0) != HP_OK) { fprintf(stderr, "error: %s\n", hp_pack_last_error()); return 1; }
    FILE* f = fopen((prefix + ".haps.tsv").c_str(), "w");
    if (!f) { fprintf(stderr, "error: cannot write %s.haps.tsv\n", prefix.c_str()); return 1; }
    fprintf(f, "block_index\tvariant\tposition\thaplotype_1\thaplotype_2\tphase_block\n");
    uint64_t failed = 0;
    for (uint32_t i = 0; i < b.n_blocks; i++) {
        if (status[i] != HP_BLOCK_OK) { failed++; continue; }
        for (uint64_t v = b.var_off[i]; v < b.var_off[i + 1]; v++)
            fprintf(f, "%u\t%" PRIu64 "\t%lld\t%u\t%u\t%" PRIu64 "\n", i, v - b.var_off[i], var_pos ? (long long)var_pos[v] : (long long)(v - b.var_off[i]),
                    h1[v], h2[v], var_pos ? tags[v] : 0);
    }
    fclose(f);
    printf("%u blocks, %" PRIu64 " variants, %" PRIu64 " reads: %" PRIu64 " blocks with a non-OK status; solver kernels %.3f ms, %" PRIu64 " launches\n",
           b.n_blocks, nv, nr, failed, hp_last_kernel_ms(ctx), hp_launch_count(ctx));
    hp_ctx_destroy(ctx);
    hp_pack_close(pk);
    return failed ? 3 : 0;
}
