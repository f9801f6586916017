// Speech-embedding conv stack: layer table + model handle shared by the fp32 and tcgen05 paths.
#pragma once
#include "hb_common.cuh"
#include <cuda_fp16.h>

namespace hb {

struct ConvLayer {
    int kh, kw, cin, cout;
    int same;        // 1: SAME padding on the freq axis (kw/2 each side); 0: VALID
    int leaky;       // LeakyReLU(0.2) after the conv
    int pool_t, pool_f;  // max-pool after the layer (1,1 = none)
};

// Mirrors heybuddy_b200/spec.py EMBEDDING_LAYERS.
static const ConvLayer kLayers[kNumConv] = {
    {1, 3, 1, 24, 1, 1, 1, 1},  {3, 1, 24, 24, 0, 1, 1, 1}, {1, 3, 24, 24, 1, 1, 1, 1}, {3, 1, 24, 24, 0, 1, 2, 2},
    {1, 3, 24, 48, 1, 1, 1, 1}, {3, 1, 48, 48, 0, 1, 1, 1}, {1, 3, 48, 48, 1, 1, 1, 1}, {3, 1, 48, 48, 0, 1, 1, 2},
    {1, 3, 48, 72, 1, 1, 1, 1}, {3, 1, 72, 72, 0, 1, 1, 1}, {1, 3, 72, 72, 1, 1, 1, 1}, {3, 1, 72, 72, 0, 1, 2, 2},
    {1, 3, 72, 96, 1, 1, 1, 1}, {3, 1, 96, 96, 0, 1, 1, 1}, {1, 3, 96, 96, 1, 1, 1, 1}, {3, 1, 96, 96, 0, 1, 2, 2},
    {1, 2, 96, 96, 0, 1, 1, 1}, {3, 1, 96, 96, 0, 1, 1, 1}, {1, 1, 96, 96, 0, 1, 1, 1}, {3, 1, 96, 96, 0, 0, 1, 1},
};

inline int64_t layer_weight_floats(const ConvLayer& l) { return (int64_t)l.kh * l.kw * l.cin * l.cout; }
inline int64_t total_weight_floats() {
    int64_t n = 0;
    for (int i = 0; i < kNumConv; ++i) n += layer_weight_floats(kLayers[i]) + kLayers[i].cout;
    return n;
}

}  // namespace hb

struct hb_embed_model {
    float* w32 = nullptr;              // packed fp32 [layer: kernel HWIO, bias]
    int64_t w_off[hb::kNumConv];       // float offset of each layer's kernel
    int64_t b_off[hb::kNumConv];       // float offset of each layer's bias
    void* tcg = nullptr;               // blocks 1-4: Toeplitz-packed weights (embed_tcg.cu)
    void* tail = nullptr;              // conv2d_16..19: per-layer B operands (embed_tail.cu)
    int device = 0;
};

namespace hb {
// embed_fp32.cu
int64_t fp32_workspace_bytes(int B, int F);
int fp32_tail_from_l15(const hb_embed_model* m, const float* pre_pool, int B, int T15, const int32_t* slot_offsets_host,
                       int n_slots, float* out, float* scratch, int64_t scratch_floats, cudaStream_t st);
int fp32_gather_slots(const float* tmp0, const float* tmp1, int J0, int J1, const int* slot_m_dev, int n_slots, float* out, int B,
                      cudaStream_t st);
int fp32_pool_public(const float* x, float* y, int n, int T, int F, int C, int pt, int pf, int phase, cudaStream_t st);
// embed_tc.cu
int tc_prepare(hb_embed_model* m, const float* weights_host);
void tc_release(hb_embed_model* m);
int64_t tc_workspace_bytes(int B, int F);
int tc_embed_clips(const hb_embed_model* m, const float* mel_dev, int B, int F, const int32_t* slot_offsets_host,
                   int n_slots, float* out_dev, void* workspace_dev, int64_t workspace_bytes, cudaStream_t stream);
int64_t tc_activation(const hb_embed_model* m, const float* mel_dev, int B, int F, int layer, float* out_dev,
                      int64_t out_capacity, void* workspace_dev, int64_t workspace_bytes, cudaStream_t stream);
// embed_tcg.cu (blocks 1 and 2, several positions per accumulator column)
int tcg_prepare(hb_embed_model* m, const float* weights_host);
void tcg_release(hb_embed_model* m);
int tcg_block1(const hb_embed_model* m, const float* mel_dev, __half* out_dev, int B, int in_T, float* dbg, int dbg_layer,
               cudaStream_t st);
int tcg_block2(const hb_embed_model* m, const __half* in_dev, __half* out_dev, int B, int in_T, float* dbg, int dbg_layer,
               cudaStream_t st);
int tcg_block3(const hb_embed_model* m, const __half* in_dev, __half* out_dev, int B, int in_T, float* dbg, int dbg_layer,
               cudaStream_t st);
int tcg_block4(const hb_embed_model* m, const __half* in_dev, __half* out_dev, int B, int in_T, float* dbg, int dbg_layer,
               cudaStream_t st, __half* pool2_out = nullptr);
int tcg_block4_rows_per_tile();
int tcg_block4_max_rows();
int tcg_check_timeout();
// embed_tail.cu (conv2d_15's pool + conv2d_16 .. 19, one launch per layer over both pool phases)
int tail_prepare(hb_embed_model* m, const float* weights_host);
void tail_release(hb_embed_model* m);
int64_t tail_scratch_bytes(int B, int T15);
int tail_run(const hb_embed_model* m, const __half* block4_out, bool pooled, int B, int T15, float* out0, float* out1, void* scratch,
             int64_t scratch_bytes, int upto, float* dbg_out, cudaStream_t st, const int32_t* slot_map_dev = nullptr,
             float* out_slots = nullptr, int n_slots = 0);
int tail_max_dup();
int tail_check_timeout();
int tail_debug_times(long long* out_host);
int tcg_debug_times(long long* out_host);
}  // namespace hb
