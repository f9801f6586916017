// K8 placeholder -- filled in by the classifier milestone.
#include "hb_common.cuh"
#define NYI { hb::set_error("classifier not built yet"); return HB_ERR_UNSUPPORTED; }
extern "C" int64_t hb_mlp_num_params(void) { return 256417; }
extern "C" int hb_mlp_create(hb_mlp_model**, const float*, int64_t) NYI
extern "C" int hb_mlp_destroy(hb_mlp_model*) NYI
extern "C" int hb_mlp_get_params(const hb_mlp_model*, float*, int64_t) NYI
extern "C" int hb_mlp_set_params(hb_mlp_model*, const float*, int64_t) NYI
extern "C" int64_t hb_mlp_workspace_bytes(int, int) NYI
extern "C" int hb_mlp_forward(const hb_mlp_model*, const float*, float*, int, void*, int64_t, void*) NYI
extern "C" int hb_mlp_train_step(hb_mlp_model*, const float*, const int64_t*, int, float, float, float, int, float*, float*, void*, int64_t, void*) NYI
extern "C" int hb_mlp_get_grads(const hb_mlp_model*, float*, int64_t) NYI
extern "C" int hb_mlp_forward_multi(hb_mlp_model* const*, int, const float*, float*, int, void*, int64_t, void*) NYI
