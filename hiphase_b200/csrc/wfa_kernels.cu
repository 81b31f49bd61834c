// wfa_kernels.cu -- graph-WFA realignment kernels (placeholder until K2 lands in this round).
#include "hp_host.h"

extern "C" {

int hp_wfa_align_batch(hp_ctx* ctx, const hp_wfa_batch*, hp_wfa_out*) {
    if (ctx) ctx->err = "hp_wfa_align_batch: not built yet";
    return HP_ERR_INTERNAL;
}

int hp_wfa_graph_align(hp_ctx* ctx, uint32_t, const uint8_t*, const uint64_t*, const uint32_t*, const uint64_t*,
                       const uint8_t*, uint64_t, uint64_t, uint32_t, int32_t*, uint32_t*, uint64_t*) {
    if (ctx) ctx->err = "hp_wfa_graph_align: not built yet";
    return HP_ERR_INTERNAL;
}

}
