// K7 (product mode) placeholder: tcgen05 implicit-GEMM conv stack -- filled in by the next milestone.
#include "embed_common.cuh"

namespace hb {
int tc_prepare(hb_embed_model*, const float*) { return HB_OK; }
void tc_release(hb_embed_model*) {}
int64_t tc_workspace_bytes(int, int) { return 0; }
int tc_embed_clips(const hb_embed_model*, const float*, int, int, const int32_t*, int, float*, void*, int64_t, cudaStream_t) {
    set_error("HB_EMBED_F16 is not built yet");
    return HB_ERR_UNSUPPORTED;
}
int64_t tc_activation(const hb_embed_model*, const float*, int, int, int, float*, int64_t, void*, int64_t, cudaStream_t) {
    set_error("HB_EMBED_F16 is not built yet");
    return HB_ERR_UNSUPPORTED;
}
}  // namespace hb
