// Shared declarations of the wake-word classifier (K8): parameter layout, the model object, and the entry points of the fused step
// (mlp_fused.cu) and of the tensor-core products (gemm_tf32.cu).  Reference: wakeword.py:171-348, modules/multi_layer_perceptron.py:115-124.
#pragma once
#include "hb_common.cuh"

namespace hb {

constexpr int kIn = 1536, kDim = 96, kHid = 64, kStages = 4;  // stages: mlp_in, layers.0, layers.1, mlp_out
constexpr float kLnEps = 1e-5f;
constexpr int kMlpMaxPack = 64;                       // models per launch of the multi-model kernels (their pointers travel as kernel arguments)

struct StageOff {     // float offsets into the packed parameter vector
    int ln_w, ln_b, in_dim, out_dim;
    int hw, hb, ow, ob, gw, gb;
};

struct MlpLayout {
    StageOff s[kStages];
    int total;
};

inline MlpLayout make_mlp_layout() {
    MlpLayout L;
    int o = 0;
    auto stage = [&](int i, int in_dim, int out_dim) {
        StageOff& s = L.s[i];
        s.in_dim = in_dim; s.out_dim = out_dim;
        s.ln_w = o; o += in_dim;
        s.ln_b = o; o += in_dim;
        s.hw = o; o += kHid * in_dim;
        s.hb = o; o += kHid;
        s.ow = o; o += out_dim * kHid;
        s.ob = o; o += out_dim;
        s.gw = o; o += kHid * in_dim;
        s.gb = o; o += kHid;
    };
    stage(0, kIn, kDim);
    stage(1, kDim, kDim);
    stage(2, kDim, kDim);
    stage(3, kDim, 1);
    L.total = o;
    return L;
}

// ---- tensor-core products (gemm_tf32.cu): three TF32 passes = fp32-level accuracy --------------------------------------------------
struct TfArgs {
    const float* A; int lda;
    const float* B0; const float* B1; int bsplit; int ldb;      // B row n: n < bsplit ? B0 + n ldb : B1 + (n - bsplit) ldb   (K-major form)
    const float* bias0; const float* bias1; int biassplit;      // bias0 == nullptr: none
    float* C; int ldc;
    int M, N, K;
    // row statistics folded into the load: K-major form: A(m, k) <- (A(m, k) - mean[m]) rstd[m];  batch-major form: B(k, n) <- (B(k, n) - mean[k]) rstd[k]
    const float* mean; const float* rstd;
    int splits; int64_t split_stride;                           // K slices (blockIdx.z); slice z writes C + z split_stride (no bias when > 1)
    int ones_col;                                               // batch-major form: B(k, ones_col) = 1 (a column of A's column sums); -1: none
    int store_n;                                                // columns of C written (>= N to include the ones column; a multiple of 4)
};
constexpr int kTfGroupMax = 9;
struct TfGroup { TfArgs pr[kTfGroupMax]; int tile0[kTfGroupMax]; int nprob; };
// K-major form:      C[M, N] = A[M][K] B[N][K]^T (+ bias)
int gemm_tf32x3_launch(const TfArgs& a, cudaStream_t st);
// batch-major form:  C_p[M, N] = sum_k A_p[k][M] B_p[k][N]   (the weight-gradient shape: both operands stored with the reduction index
// outermost), a group of products (M <= 128 each, same K and K split) in one launch
int gemm_tf32x3_launch_group(TfGroup& g, cudaStream_t st);

}  // namespace hb

struct hb_mlp_model {
    float loss_scale = 1.0f;   // multiplies the loss and its gradients (the reference divides by its accumulation counter, trainer.py:441)
    float* p = nullptr;   // parameters
    float* g = nullptr;   // gradients of the last training step
    float* m = nullptr;   // Adam first moment
    float* v = nullptr;   // Adam second moment
    int* step = nullptr;
};

namespace hb {

// ---- the fused step (mlp_fused.cu) ---------------------------------------------------------------------------------------------------
int64_t mlp_fused_ws_floats(int B, int training);
// logits[B] of the batch; with `training` the activations the backward pass needs stay in `ws`
int mlp_fused_forward(const hb_mlp_model* m, const float* x, int B, float* ws, int training, const float** logits, cudaStream_t st);
// the multi-model forward's remainder (config 5): see mlp_fused.cu
int64_t mlp_fused_multi_wt_floats(int n_models);
int mlp_fused_multi_tail(const float* const* params, int n, const float* hg, int64_t ld, int B, float* wt_all, float* logits, cudaStream_t st);
// dz[B] = d loss / d logit  ->  every parameter's gradient, into g_out (packed like the parameters) or, when NULL, the model's own buffer
int mlp_fused_backward(hb_mlp_model* m, const float* x, int B, float* ws, const float* dz, cudaStream_t st, float* g_out = nullptr);

}  // namespace hb
