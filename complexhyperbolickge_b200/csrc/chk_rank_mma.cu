// K2 (tensor-core tier) — placeholder until the tcgen05 kernel lands; fails loudly, never falls back.
#include "chk_common.cuh"

int chk_rank_counts_mma(int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                        const void* target, const void* entity, const void* hn, const void* bt,
                        int64_t n_rows, const void* shadow, void* workspace, int64_t workspace_bytes,
                        int64_t* counts, cudaStream_t st) {
    chk_set_error("CHK_RANK_MMA not built yet");
    return CHK_EUNSUPPORTED;
}
extern "C" int64_t chk_entity_shadow_bytes(int rank, int64_t n_rows) { return 0; }
extern "C" int chk_entity_shadow_build(int rank, int64_t n_rows, const void* entity_f32, void* shadow, void* stream) {
    chk_set_error("CHK_RANK_MMA not built yet");
    return CHK_EUNSUPPORTED;
}
extern "C" int64_t chk_rank_mma_workspace_bytes(int rank, int64_t b) { return 0; }
