// (i) compact long reads and (ii) backbone edge table.
#include "common.cuh"
struct K12State {};
void k12_state_destroy(K12State* s) { delete s; }

extern "C" int hgpu_compact_lr(hgpu_t* ctx, const hgpu_hits_t*, const uint32_t*, uint32_t, const double*, uint32_t,
                               const hgpu_k1_params*, hgpu_cl_elem*, uint32_t*, uint64_t*) {
    if (!ctx) return HGPU_E_INVALID;
    HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "hgpu_compact_lr: not built yet");
}
extern "C" int hgpu_backbone_edges(hgpu_t* ctx, const uint32_t*, const uint8_t*, const uint32_t*, uint32_t, uint32_t,
                                   uint64_t*, uint32_t*, hgpu_edge_supp*, uint8_t*, uint64_t*) {
    if (!ctx) return HGPU_E_INVALID;
    HGPU_FAIL(ctx, HGPU_E_UNSUPPORTED, "hgpu_backbone_edges: not built yet");
}
