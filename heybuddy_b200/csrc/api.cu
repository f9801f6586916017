// Library-level C-ABI entry points: version and the thread-local error string.
#include "hb_common.cuh"

#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

namespace hb {
static thread_local char g_error[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
}  // namespace hb

extern "C" int hb_abi_version(void) { return HB_ABI_VERSION; }
extern "C" const char* hb_last_error(void) { return hb::g_error; }
extern "C" int64_t hb_launch_count(void) { return hb::g_launches.load(); }

// ---- fused featurization: ragged int16 clips -> embeddings, one call -----------------------------------------------
// a1 + K1-K4 + K6 (hb_augment_mel_i16: one kernel, audio x 32767) -> K7 (hb_embed_clips) through one workspace.
static int64_t featurize_align(int64_t n) { return (n + 255) & ~255ll; }

extern "C" int64_t hb_featurize_workspace_bytes(int n, int T, int mode) {
    if (n < 0 || T < 0) return HB_ERR_INVALID;
    const int F = hb_mel_frames(T);
    const int64_t ws = hb_embed_clips_workspace_bytes(n, F, mode);
    if (ws < 0) return ws;
    return featurize_align((int64_t)n * T * 4) + featurize_align((int64_t)n * F * hb::kMels * 4) + featurize_align(ws);
}

extern "C" int hb_featurize_i16(const hb_embed_model* m, int mode, const int16_t* samples_dev, const int64_t* offsets_dev,
                                const int32_t* pad_before_dev, const float* noise_bank_dev, const float* colored_bases_dev,
                                const float* rir_spec_bank_dev, const hb_clip_aug* params_dev, const int32_t* slot_offsets_host,
                                int n_slots, float* out_dev, int n, int T, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && out_dev && workspace_dev && n >= 0, "hb_featurize_i16: bad argument");
    HB_REQUIRE(workspace_bytes >= hb_featurize_workspace_bytes(n, T, mode), "hb_featurize_i16: workspace too small");
    if (n == 0) return HB_OK;
    const int F = hb_mel_frames(T);
    unsigned char* base = reinterpret_cast<unsigned char*>(workspace_dev);
    float* audio = reinterpret_cast<float*>(base);
    float* mel = reinterpret_cast<float*>(base + featurize_align((int64_t)n * T * 4));
    void* ews = base + featurize_align((int64_t)n * T * 4) + featurize_align((int64_t)n * F * hb::kMels * 4);
    // production mode: augmentation and mel in ONE kernel (the augmented clip stays in shared memory); HB_FEATURIZE_STAGED=1 runs the
    // bit-identical parity pair hb_augment_clips_i16 -> hb_mel_f32 through the f32 [n][T] intermediate instead
    static const bool staged = getenv("HB_FEATURIZE_STAGED") != nullptr && atoi(getenv("HB_FEATURIZE_STAGED")) != 0;
    int rc;
    if (staged) {
        rc = hb_augment_clips_i16(samples_dev, offsets_dev, pad_before_dev, noise_bank_dev, colored_bases_dev, rir_spec_bank_dev,
                                  params_dev, audio, n, T, stream);
        if (rc) return rc;
        if ((rc = hb_mel_f32(audio, T, 32767.0f, mel, n, T, stream))) return rc;
    } else if ((rc = hb_augment_mel_i16(samples_dev, offsets_dev, pad_before_dev, noise_bank_dev, colored_bases_dev, rir_spec_bank_dev,
                                        params_dev, 32767.0f, mel, n, T, stream))) {
        return rc;
    }
    return hb_embed_clips(m, mode, mel, n, F, slot_offsets_host, n_slots, out_dev, ews,
                          hb_embed_clips_workspace_bytes(n, F, mode), stream);
}
