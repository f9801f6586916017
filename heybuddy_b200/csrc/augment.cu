// K1-K4 placeholder -- filled in by the augmentation milestone.
#include "hb_common.cuh"
extern "C" int hb_rir_spectrum(const float*, float*, int, void*) { hb::set_error("augment not built yet"); return HB_ERR_UNSUPPORTED; }
extern "C" int hb_augment_f32(const float*, const float*, const float*, const float*, const float*, const hb_augment_draws*, float*, int, int, void*) { hb::set_error("augment not built yet"); return HB_ERR_UNSUPPORTED; }
extern "C" int hb_fix_length_i16(const int16_t*, const int64_t*, const int32_t*, float*, int, int, void*) { hb::set_error("augment not built yet"); return HB_ERR_UNSUPPORTED; }
