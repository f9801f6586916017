// K8: wake-word gated-MLP classifier -- batched forward, and one fused training step (forward,
// high-loss selection, weighted BCE, backward, Adam) -- as hand-written fp32 CUDA kernels.
//
// Replaces WakeWordMLPModel.forward (reference src/python/heybuddy/wakeword.py:334-348, with
// GatedMultiLayerPerceptron.forward, modules/multi_layer_perceptron.py:115-124) and the
// loss / backward / optimizer lines of WakeWordTrainer.train_epoch (trainer.py:405-462).
//
//   x[B,1536] -> LN(1536) -> gated(1536->64->96) -> 2 x [LN(96) -> gated(96->64->96)] -> LN(96)
//             -> gated(96->64->1) -> sigmoid            gated(u) = W_o (silu(W_h u + b_h) * (W_g u + b_g)) + b_o
//
// Parameters are packed in state_dict order (heybuddy_b200/spec.py classifier_param_shapes) so .pt
// checkpoints and the in-repo ONNX initializers interchange.  The step is launch-latency / HBM
// bound (25 MB of input per 4096-row batch, 1 MB of parameters); fp32 FMA keeps logits within 1e-3
// of the reference.
#include "hb_common.cuh"

#include <math.h>

namespace hb {

constexpr int kIn = 1536, kDim = 96, kHid = 64, kStages = 4;  // stages: mlp_in, layers.0, layers.1, mlp_out
constexpr float kLnEps = 1e-5f;

struct StageOff {     // float offsets into the packed parameter vector
    int ln_w, ln_b, in_dim, out_dim;
    int hw, hb, ow, ob, gw, gb;
};

struct MlpLayout {
    StageOff s[kStages];
    int total;
};

static MlpLayout make_layout() {
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
static const MlpLayout kLayout = make_layout();

// ---- kernels -------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row.
__global__ void ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                              float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int B, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* xr = x + (int64_t)row * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i];
    const float mu = warp_sum(s) / D;
    float v = 0.f;
    for (int i = lane; i < D; i += 32) { const float d = xr[i] - mu; v += d * d; }
    const float rs = rsqrtf(warp_sum(v) / D + kLnEps);
    float* yr = y + (int64_t)row * D;
    for (int i = lane; i < D; i += 32) yr[i] = (xr[i] - mu) * rs * w[i] + b[i];
    if (lane == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
}

// Generic tiled fp32 GEMM: C[m,n] (+)= sum_k A(m,k) * Bm(k,n) (+ bias[n]); element strides make NT / NN / TN.
constexpr int kTile = 64, kTk = 16;
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                   const float* __restrict__ Bm, int64_t sbk, int64_t sbn,
                                                   float* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                                                   int M, int N, int K, int accumulate, int k_per_split) {
    // split-K (gridDim.z > 1): slice z of K goes to the partial buffer C + z * M * ldc (ldc = N), summed in a fixed order by
    // splitk_reduce_kernel -- deterministic, unlike atomics
    __shared__ float As[kTk][kTile + 1];
    __shared__ float Bs[kTk][kTile + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
    const int k_begin = blockIdx.z * k_per_split;
    if (gridDim.z > 1) {
        C += (int64_t)blockIdx.z * M * ldc;
        K = min(K, k_begin + k_per_split);
    }
    float acc[4][4] = {};
    for (int k0 = k_begin; k0 < K; k0 += kTk) {
        for (int i = threadIdx.x; i < kTile * kTk; i += 256) {
            int m, k;
            if (sak == 1) { k = i % kTk; m = i / kTk; } else { m = i % kTile; k = i / kTile; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < K) ? A[gm * sam + gk * sak] : 0.f;
        }
        for (int i = threadIdx.x; i < kTile * kTk; i += 256) {
            int n, k;
            if (sbk == 1) { k = i % kTk; n = i / kTk; } else { n = i % kTile; k = i / kTile; }
            const int gn = n0 + n, gk = k0 + k;
            Bs[k][n] = (gn < N && gk < K) ? Bm[gk * sbk + gn * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTk; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty + 16 * i]; b[i] = Bs[k][tx + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty + 16 * i, gn = n0 + tx + 16 * j;
            if (gm < M && gn < N) {
                float v = acc[i][j] + (bias ? bias[gn] : 0.f);
                float* c = C + gm * ldc + gn;
                *c = accumulate ? *c + v : v;
            }
        }
}

__global__ void gate_fwd_kernel(const float* __restrict__ h, const float* __restrict__ g, float* __restrict__ a, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float hv = h[i];
        a[i] = hv / (1.f + expf(-hv)) * g[i];
    }
}

__global__ void gate_bwd_kernel(const float* __restrict__ h, const float* __restrict__ g, const float* __restrict__ da,
                                float* __restrict__ dh, float* __restrict__ dg, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float hv = h[i], sg = 1.f / (1.f + expf(-hv));
        const float silu = hv * sg;
        const float d = da[i];
        dg[i] = d * silu;
        dh[i] = d * g[i] * (sg * (1.f + hv * (1.f - sg)));
    }
}

__global__ void sigmoid_kernel(const float* __restrict__ z, float* __restrict__ p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1.f / (1.f + expf(-z[i]));
}

// stats = {loss_sum -> loss, n_selected, stepped, high_loss_rate}
__global__ void head_select_kernel(const float* __restrict__ z, const int64_t* __restrict__ y, float* __restrict__ p,
                                   float* __restrict__ stats, int B, float thr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int sel = 0;
    if (i < B) {
        const float pr = 1.f / (1.f + expf(-z[i]));
        p[i] = pr;
        sel = (y[i] == 0 && pr >= thr) || (y[i] == 1 && pr < 1.f - thr);
    }
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&stats[1], (float)__popc(m));
}

// n_total: rows selected over the whole (possibly multi-rank) batch = the normaliser of the mean loss
__global__ void head_grad_kernel(const float* __restrict__ p, const int64_t* __restrict__ y, float* __restrict__ dz,
                                 float* __restrict__ stats, const float* __restrict__ n_total, int B, float thr, float neg_w, float loss_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float n_sel = *n_total / loss_scale;     // loss_scale = 1 / accumulation_steps of the reference's loop
    float l = 0.f;
    if (i < B) {
        const float pr = p[i];
        const bool pos = y[i] == 1;
        const bool sel = (!pos && pr >= thr) || (pos && pr < 1.f - thr);
        float g = 0.f;
        if (sel && n_sel > 0.f) {
            const float w = pos ? 1.f : neg_w;
            // F.binary_cross_entropy clamps log() at -100
            const float lp = fmaxf(logf(pr), -100.f), l1p = fmaxf(logf(1.f - pr), -100.f);
            l = -w * (pos ? lp : l1p) / n_sel;
            g = w * (pr - (pos ? 1.f : 0.f)) / n_sel;   // d loss / d logit
        }
        dz[i] = g;
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(&stats[0], l);
}

__global__ void head_finish_kernel(float* __restrict__ stats, const float* __restrict__ n_total, int B, int min_selected) {
    stats[2] = *n_total >= (float)min_selected ? 1.f : 0.f;
    stats[3] = stats[1] / (float)B;
}

// column sums: out[n] (+)= sum_m X[m, n]
__global__ void colsum_kernel(const float* __restrict__ X, float* __restrict__ out, int M, int N, int accumulate) {
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int part = threadIdx.x >> 5;  // 8 row partitions
    __shared__ float red[8][33];
    float s = 0.f;
    if (n < N)
        for (int m = part; m < M; m += 8) s += X[(int64_t)m * N + n];
    red[part][threadIdx.x & 31] = s;
    __syncthreads();
    if (part == 0 && n < N) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
        out[n] = accumulate ? out[n] + t : t;
    }
}

// LayerNorm backward: dgamma[n] = sum_m dy*xhat, dbeta[n] = sum_m dy; optionally dx.
// xhat is recomputed from the LN output: xhat = (y - beta) / gamma is unsafe for gamma = 0, so the caller passes x, mean, rstd.
__global__ void ln_bwd_params_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                     int M, int N) {
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int part = threadIdx.x >> 5;
    __shared__ float r1[8][33], r2[8][33];
    float a = 0.f, b = 0.f;
    if (n < N)
        for (int m = part; m < M; m += 8) {
            const float d = dy[(int64_t)m * N + n];
            a += d * (x[(int64_t)m * N + n] - mean[m]) * rstd[m];
            b += d;
        }
    r1[part][threadIdx.x & 31] = a;
    r2[part][threadIdx.x & 31] = b;
    __syncthreads();
    if (part == 0 && n < N) {
        float ta = 0.f, tb = 0.f;
        for (int i = 0; i < 8; ++i) { ta += r1[i][threadIdx.x & 31]; tb += r2[i][threadIdx.x & 31]; }
        dgamma[n] = ta;
        dbeta[n] = tb;
    }
}

// dx = rstd * (dyg - mean(dyg) - xhat * mean(dyg * xhat)), dyg = dy * gamma; one warp per row
__global__ void ln_bwd_input_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
                                    int B, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float mu = mean[row], rs = rstd[row];
    const float* dyr = dy + (int64_t)row * D;
    const float* xr = x + (int64_t)row * D;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float g = dyr[i] * gamma[i], xh = (xr[i] - mu) * rs;
        s1 += g;
        s2 += g * xh;
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    for (int i = lane; i < D; i += 32) {
        const float g = dyr[i] * gamma[i], xh = (xr[i] - mu) * rs;
        dx[(int64_t)row * D + i] = rs * (g - s1 - xh * s2);
    }
}

// torch.optim.Adam defaults (betas 0.9/0.999, eps 1e-8, no weight decay); a no-op when stats[2] == 0
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int* __restrict__ step, const float* __restrict__ stats, float lr, int n) {
    if (stats[2] == 0.f) return;
    const int t = *step + 1;
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float bc1 = 1.f - powf(b1, (float)t), bc2 = 1.f - powf(b2, (float)t);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
    }
}
__global__ void adam_step_inc_kernel(int* step, const float* stats) {
    if (stats[2] != 0.f) *step += 1;
}

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

// C[m, n] (+)= bias[n] + sum over the S partial results, in slice order
__global__ void splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                                     int M, int N, int S, int accumulate) {
    const int64_t total = (int64_t)M * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / N), n = (int)(i - (int64_t)m * N);
        float v = bias ? bias[n] : 0.f;
        for (int z = 0; z < S; ++z) v += part[(int64_t)z * total + i];
        float* c = C + (int64_t)m * ldc + n;
        *c = accumulate ? *c + v : v;
    }
}

constexpr int64_t kPartFloats = 16ll * 64 * 1536;     // split-K partial buffer: up to 16 slices of the largest weight gradient

// part != nullptr: GEMMs with few output tiles and a long K (the weight gradients: K = batch) are split over K so they fill
// the GPU; partial sums go through `part` and are reduced in a fixed order.
static int gemm(const float* A, int64_t sam, int64_t sak, const float* Bm, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
                const float* bias, int M, int N, int K, int accumulate, cudaStream_t st, float* part = nullptr) {
    dim3 grid(ceil_div(N, kTile), ceil_div(M, kTile));
    const int tiles = (int)(grid.x * grid.y);
    int S = 1;
    if (part != nullptr && tiles < 148 && K >= 512) {
        S = std::min(std::min(K / 128, 32), ceil_div(2 * 148, tiles));
        while (S > 1 && (int64_t)S * M * N > kPartFloats) --S;
    }
    if (S <= 1) {
        gemm_kernel<<<grid, 256, 0, st>>>(A, sam, sak, Bm, sbk, sbn, C, ldc, bias, M, N, K, accumulate, K);
        HB_LAUNCHED();
        return HB_OK;
    }
    const int k_per = ceil_div(ceil_div(K, S), kTk) * kTk;
    grid.z = ceil_div(K, k_per);
    gemm_kernel<<<grid, 256, 0, st>>>(A, sam, sak, Bm, sbk, sbn, part, N, nullptr, M, N, K, 0, k_per);
    HB_LAUNCHED();
    const int64_t total = (int64_t)M * N;
    splitk_reduce_kernel<<<(int)std::min<int64_t>(ceil_div64(total, 256), 1184), 256, 0, st>>>(part, C, ldc, bias, M, N, (int)grid.z, accumulate);
    HB_LAUNCHED();
    return HB_OK;
}
// y[B,N] = x[B,K] W[N,K]^T + b
static int linear_fwd(const float* x, const float* W, const float* b, float* y, int B, int N, int K, cudaStream_t st, float* part = nullptr) {
    return gemm(x, K, 1, W, 1, K, y, N, b, B, N, K, 0, st, part);
}

// workspace carve-up (floats)
struct Ws {
    float *xn, *mean[kStages], *rstd[kStages];
    float *u[kStages];       // LN output = stage input (u[0] = xn)
    float *h[kStages], *g[kStages], *a[kStages], *o[kStages];
    float *dz, *d_o, *d_a, *d_h, *d_g, *d_u, *d_x;
    float* stats_tmp;
    float* part;             // split-K partial sums
};
static int64_t carve(Ws* w, float* base, int B, int training) {
    int64_t off = 0;
    auto take = [&](int64_t n) { float* p = base ? base + off : nullptr; off += (n + 63) & ~63ll; return p; };
    for (int s = 0; s < kStages; ++s) {
        const int in_dim = kLayout.s[s].in_dim, out_dim = kLayout.s[s].out_dim;
        w->u[s] = take((int64_t)B * in_dim);
        w->mean[s] = take(B);
        w->rstd[s] = take(B);
        w->h[s] = take((int64_t)B * kHid);
        w->g[s] = take((int64_t)B * kHid);
        w->a[s] = take((int64_t)B * kHid);
        w->o[s] = take((int64_t)B * out_dim);
    }
    w->xn = w->u[0];
    if (training) {
        w->dz = take(B);
        w->d_o = take((int64_t)B * kDim);
        w->d_a = take((int64_t)B * kHid);
        w->d_h = take((int64_t)B * kHid);
        w->d_g = take((int64_t)B * kHid);
        w->d_u = take((int64_t)B * kIn);
        w->d_x = take((int64_t)B * kDim);
    }
    w->part = take(kPartFloats);
    return off;
}

static int forward_impl(const hb_mlp_model* m, const float* x, int B, const Ws& w, cudaStream_t st) {
    const float* cur = x;
    for (int s = 0; s < kStages; ++s) {
        const StageOff& L = kLayout.s[s];
        ln_fwd_kernel<<<ceil_div(B, 8), 256, 0, st>>>(cur, m->p + L.ln_w, m->p + L.ln_b, w.u[s], w.mean[s], w.rstd[s], B, L.in_dim);
        HB_LAUNCHED();
        int rc;
        if ((rc = linear_fwd(w.u[s], m->p + L.hw, m->p + L.hb, w.h[s], B, kHid, L.in_dim, st, w.part))) return rc;
        if ((rc = linear_fwd(w.u[s], m->p + L.gw, m->p + L.gb, w.g[s], B, kHid, L.in_dim, st, w.part))) return rc;
        const int64_t n = (int64_t)B * kHid;
        gate_fwd_kernel<<<(int)std::min<int64_t>(ceil_div64(n, 256), 1184), 256, 0, st>>>(w.h[s], w.g[s], w.a[s], n);
        HB_LAUNCHED();
        if ((rc = linear_fwd(w.a[s], m->p + L.ow, m->p + L.ob, w.o[s], B, L.out_dim, kHid, st))) return rc;
        cur = w.o[s];
    }
    return HB_OK;
}

}  // namespace hb

using namespace hb;

extern "C" int64_t hb_mlp_num_params(void) { return kLayout.total; }

extern "C" int hb_mlp_create(hb_mlp_model** out, const float* params_host, int64_t n_floats) {
    HB_REQUIRE(out && params_host, "hb_mlp_create: null pointer");
    HB_REQUIRE(n_floats == kLayout.total, "hb_mlp_create: expected %d floats (default architecture), got %lld", kLayout.total,
               (long long)n_floats);
    hb_mlp_model* m = new hb_mlp_model();
    const size_t bytes = (size_t)n_floats * sizeof(float);
    HB_CUDA_OK(cudaMalloc(&m->p, bytes));
    HB_CUDA_OK(cudaMalloc(&m->g, bytes));
    HB_CUDA_OK(cudaMalloc(&m->m, bytes));
    HB_CUDA_OK(cudaMalloc(&m->v, bytes));
    HB_CUDA_OK(cudaMalloc(&m->step, sizeof(int)));
    HB_CUDA_OK(cudaMemcpy(m->p, params_host, bytes, cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemset(m->g, 0, bytes));
    HB_CUDA_OK(cudaMemset(m->m, 0, bytes));
    HB_CUDA_OK(cudaMemset(m->v, 0, bytes));
    HB_CUDA_OK(cudaMemset(m->step, 0, sizeof(int)));
    *out = m;
    return HB_OK;
}

extern "C" int hb_mlp_destroy(hb_mlp_model* m) {
    if (!m) return HB_OK;
    cudaFree(m->p); cudaFree(m->g); cudaFree(m->m); cudaFree(m->v); cudaFree(m->step);
    delete m;
    return HB_OK;
}

extern "C" int hb_mlp_get_params(const hb_mlp_model* m, float* params_host, int64_t n_floats) {
    HB_REQUIRE(m && params_host && n_floats == kLayout.total, "hb_mlp_get_params: bad argument");
    HB_CUDA_OK(cudaMemcpy(params_host, m->p, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return HB_OK;
}

extern "C" int hb_mlp_set_params(hb_mlp_model* m, const float* params_host, int64_t n_floats) {
    HB_REQUIRE(m && params_host && n_floats == kLayout.total, "hb_mlp_set_params: bad argument");
    HB_CUDA_OK(cudaMemcpy(m->p, params_host, (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice));
    return HB_OK;
}

extern "C" int hb_mlp_get_grads(const hb_mlp_model* m, float* grads_host, int64_t n_floats) {
    HB_REQUIRE(m && grads_host && n_floats == kLayout.total, "hb_mlp_get_grads: bad argument");
    HB_CUDA_OK(cudaMemcpy(grads_host, m->g, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return HB_OK;
}

extern "C" int hb_mlp_set_loss_scale(hb_mlp_model* m, float scale) {
    HB_REQUIRE(m && scale > 0.f, "hb_mlp_set_loss_scale: bad argument");
    m->loss_scale = scale;
    return HB_OK;
}

// Adam state (torch.optim.Adam's exp_avg / exp_avg_sq / step) <-> host, packed like the parameters: checkpoint + resume
extern "C" int hb_mlp_get_adam(const hb_mlp_model* m, float* exp_avg_host, float* exp_avg_sq_host, int* step_host, int64_t n_floats) {
    HB_REQUIRE(m && exp_avg_host && exp_avg_sq_host && step_host && n_floats == kLayout.total, "hb_mlp_get_adam: bad argument");
    HB_CUDA_OK(cudaMemcpy(exp_avg_host, m->m, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    HB_CUDA_OK(cudaMemcpy(exp_avg_sq_host, m->v, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    HB_CUDA_OK(cudaMemcpy(step_host, m->step, sizeof(int), cudaMemcpyDeviceToHost));
    return HB_OK;
}

extern "C" int hb_mlp_set_adam(hb_mlp_model* m, const float* exp_avg_host, const float* exp_avg_sq_host, int step, int64_t n_floats) {
    HB_REQUIRE(m && exp_avg_host && exp_avg_sq_host && step >= 0 && n_floats == kLayout.total, "hb_mlp_set_adam: bad argument");
    HB_CUDA_OK(cudaMemcpy(m->m, exp_avg_host, (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(m->v, exp_avg_sq_host, (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(m->step, &step, sizeof(int), cudaMemcpyHostToDevice));
    return HB_OK;
}

namespace hb {
// one Philox4x32-10 block per four elements: counter = (index / 4, 7, call lo, call hi), key = seed (same generator as the draw table)
__device__ __forceinline__ uint4 philox_cls(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// nn.Dropout(p) on the classifier input (wakeword.py:197,338): y = x * keep / (1 - p), keep ~ Bernoulli(1 - p) per element
__global__ void dropout_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t n4, float p, uint2 key, uint64_t call) {
    const float scale = 1.0f / (1.0f - p);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = philox_cls(make_uint4((uint32_t)i, 7u ^ ((uint32_t)(i >> 32) << 8), (uint32_t)call, (uint32_t)(call >> 32)), key);
        const float4 v = x[i];
        const float t = p * 16777216.0f;     // keep when the 24-bit uniform >= p
        y[i] = make_float4((float)(r.x >> 8) >= t ? v.x * scale : 0.f, (float)(r.y >> 8) >= t ? v.y * scale : 0.f,
                           (float)(r.z >> 8) >= t ? v.z * scale : 0.f, (float)(r.w >> 8) >= t ? v.w * scale : 0.f);
    }
}
}  // namespace hb

extern "C" int hb_mlp_dropout(const float* x_dev, float* y_dev, int64_t n, float p, uint64_t seed, uint64_t call, void* stream) {
    HB_REQUIRE(x_dev && y_dev && n >= 0 && n % 4 == 0 && p >= 0.f && p < 1.f, "hb_mlp_dropout: bad argument (n must be a multiple of 4)");
    HB_REQUIRE(((reinterpret_cast<uintptr_t>(x_dev) | reinterpret_cast<uintptr_t>(y_dev)) & 15) == 0, "hb_mlp_dropout: 16-byte aligned buffers");
    if (n == 0) return HB_OK;
    const int64_t n4 = n / 4;
    dropout_kernel<<<(int)std::min<int64_t>(ceil_div64(n4, 256), 2368), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x_dev), reinterpret_cast<float4*>(y_dev), n4, p, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), call);
    HB_LAUNCHED();
    return HB_OK;
}

extern "C" int64_t hb_mlp_workspace_bytes(int B, int training) {
    if (B < 0) return HB_ERR_INVALID;
    Ws w;
    return carve(&w, nullptr, B, training) * (int64_t)sizeof(float) + 256;
}

extern "C" int hb_mlp_forward(const hb_mlp_model* m, const float* x_dev, float* prob_dev, int B, void* workspace_dev,
                              int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && x_dev && prob_dev && workspace_dev, "hb_mlp_forward: null pointer");
    HB_REQUIRE(workspace_bytes >= hb_mlp_workspace_bytes(B, 0), "hb_mlp_forward: workspace too small");
    if (B == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Ws w;
    carve(&w, reinterpret_cast<float*>(workspace_dev), B, 0);
    int rc = forward_impl(m, x_dev, B, w, st);
    if (rc) return rc;
    sigmoid_kernel<<<ceil_div(B, 256), 256, 0, st>>>(w.o[kStages - 1], prob_dev, B);
    HB_LAUNCHED();
    return HB_OK;
}

// ---- config 5: M models on the same inputs, one launch chain --------------------------------------------------------------
// The models differ in every parameter, including norm_in's affine -- but LN(x) = xhat * gamma + beta with xhat = (x - mu) * rstd
// shared by all of them, so gamma folds into the first-layer weights and beta into their bias:
//     W'_m[n][k] = W_m[n][k] * gamma_m[k],   b'_m[n] = b_m[n] + sum_k W_m[n][k] * beta_m[k]
// and the 2 M first-layer products (hidden and gate of every model; 97 % of a forward pass's FLOPs) become ONE stacked GEMM
// xhat[B,1536] x W'[M*128,1536]^T that reads the input once.  The 96-wide remainder of every model runs as GEMMs batched over the
// models (blockIdx.z = model).  The fold is redone on every call (it reads the models' live parameters; 1 MB per model).
namespace hb {

constexpr int kMaxPack = 64;
struct PtrPack { const float* p[kMaxPack]; };

// xhat = (x - mean) * rstd, one warp per row (LayerNorm without the affine)
__global__ void xhat_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* xr = x + (int64_t)row * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i];
    const float mu = warp_sum(s) / D;
    float v = 0.f;
    for (int i = lane; i < D; i += 32) { const float d = xr[i] - mu; v += d * d; }
    const float rs = rsqrtf(warp_sum(v) / D + kLnEps);
    float* yr = y + (int64_t)row * D;
    for (int i = lane; i < D; i += 32) yr[i] = (xr[i] - mu) * rs;
}

// one warp per stacked row (model m, n in [0,128): n < 64 hidden, else gate): W' row and b'
__global__ void fold_first_layer_kernel(PtrPack pk, int m0, int M, StageOff L, float* __restrict__ Wst, float* __restrict__ bst) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M * 2 * kHid) return;
    const int m = row / (2 * kHid), n = row - m * 2 * kHid;
    const float* p = pk.p[m];
    const float* W = p + (n < kHid ? L.hw + n * kIn : L.gw + (n - kHid) * kIn);
    const float* gamma = p + L.ln_w;
    const float* beta = p + L.ln_b;
    float* out = Wst + ((int64_t)(m0 + m) * 2 * kHid + n) * kIn;
    float acc = 0.f;
    for (int k = lane; k < kIn; k += 32) {
        const float w = W[k];
        out[k] = w * gamma[k];
        acc = fmaf(w, beta[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) bst[(int64_t)(m0 + m) * 2 * kHid + n] = acc + p[(n < kHid ? L.hb + n : L.gb + n - kHid)];
}

// HG [B][Mtot*128] -> A [Mtot][B][64] = silu(h) * g
__global__ void gate_stacked_kernel(const float* __restrict__ hg, float* __restrict__ a, int B, int M) {
    const int64_t n = (int64_t)B * M * kHid;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % kHid);
        const int64_t r = i / kHid;
        const int m = (int)(r % M);
        const int64_t b = r / M;
        const float hv = hg[(b * M + m) * 2 * kHid + j], gv = hg[(b * M + m) * 2 * kHid + kHid + j];
        a[((int64_t)m * B + b) * kHid + j] = hv / (1.f + expf(-hv)) * gv;
    }
}

// LayerNorm with per-model affine: x, y [M][B][D]; blockIdx.y = model
__global__ void ln_fwd_batched_kernel(const float* __restrict__ x, PtrPack pk, int w_off, int b_off, float* __restrict__ y, int B, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, m = blockIdx.y;
    if (row >= B) return;
    const float* w = pk.p[m] + w_off;
    const float* b = pk.p[m] + b_off;
    const float* xr = x + ((int64_t)m * B + row) * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i];
    const float mu = warp_sum(s) / D;
    float v = 0.f;
    for (int i = lane; i < D; i += 32) { const float d = xr[i] - mu; v += d * d; }
    const float rs = rsqrtf(warp_sum(v) / D + kLnEps);
    float* yr = y + ((int64_t)m * B + row) * D;
    for (int i = lane; i < D; i += 32) yr[i] = (xr[i] - mu) * rs * w[i] + b[i];
}

// y[m] [B][N] = x[m] [B][K] W_m[N][K]^T + b_m, batched over the models (blockIdx.z = model); same tile loop as gemm_kernel
__global__ void __launch_bounds__(256) linear_batched_kernel(const float* __restrict__ X, PtrPack pk, int w_off, int b_off,
                                                             float* __restrict__ Y, int B, int N, int K) {
    __shared__ float As[kTk][kTile + 1];
    __shared__ float Bs[kTk][kTile + 1];
    const int m = blockIdx.z;
    const float* A = X + (int64_t)m * B * K;
    const float* W = pk.p[m] + w_off;
    const float* bias = pk.p[m] + b_off;
    float* C = Y + (int64_t)m * B * N;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += kTk) {
        for (int i = threadIdx.x; i < kTile * kTk; i += 256) {
            const int k = i % kTk, r = i / kTk;
            As[k][r] = (m0 + r < B && k0 + k < K) ? A[(int64_t)(m0 + r) * K + k0 + k] : 0.f;
            Bs[k][r] = (n0 + r < N && k0 + k < K) ? W[(int64_t)(n0 + r) * K + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTk; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty + 16 * i]; b[i] = Bs[k][tx + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty + 16 * i, gn = n0 + tx + 16 * j;
            if (gm < B && gn < N) C[(int64_t)gm * N + gn] = acc[i][j] + bias[gn];
        }
}

struct MultiWs {
    float *xhat, *Wst, *bst, *hg, *a, *o, *u, *h, *g;
    float* part;
};
static int64_t carve_multi(MultiWs* w, float* base, int M, int B) {
    int64_t off = 0;
    auto take = [&](int64_t n) { float* p = base ? base + off : nullptr; off += (n + 63) & ~63ll; return p; };
    w->xhat = take((int64_t)B * kIn);
    w->Wst = take((int64_t)M * 2 * kHid * kIn);
    w->bst = take((int64_t)M * 2 * kHid);
    w->hg = take((int64_t)B * M * 2 * kHid);
    w->a = take((int64_t)M * B * kHid);
    w->o = take((int64_t)M * B * kDim);
    w->u = take((int64_t)M * B * kDim);
    w->h = take((int64_t)M * B * kHid);
    w->g = take((int64_t)M * B * kHid);
    w->part = take(kPartFloats);
    return off;
}

}  // namespace hb

extern "C" int64_t hb_mlp_multi_workspace_bytes(int M, int B) {
    if (M < 0 || B < 0) return HB_ERR_INVALID;
    MultiWs w;
    return carve_multi(&w, nullptr, M, B) * (int64_t)sizeof(float) + 256;
}

extern "C" int hb_mlp_forward_multi(hb_mlp_model* const* models, int M, const float* x_dev, float* prob_dev, int B,
                                    void* workspace_dev, int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(models && M >= 0 && B >= 0, "hb_mlp_forward_multi: bad argument");
    if (M == 0 || B == 0) return HB_OK;
    HB_REQUIRE(x_dev && prob_dev && workspace_dev, "hb_mlp_forward_multi: null pointer");
    HB_REQUIRE(workspace_bytes >= hb_mlp_multi_workspace_bytes(M, B), "hb_mlp_forward_multi: workspace too small (hb_mlp_multi_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    MultiWs w;
    carve_multi(&w, reinterpret_cast<float*>(workspace_dev), M, B);
    int rc;
    // 1. shared normalised input
    xhat_kernel<<<ceil_div(B, 8), 256, 0, st>>>(x_dev, w.xhat, B, kIn);
    HB_LAUNCHED();
    // 2. fold norm_in into the stacked first layer (packs of up to 64 models per launch: the pointers travel as kernel arguments)
    for (int m0 = 0; m0 < M; m0 += kMaxPack) {
        PtrPack pk;
        const int n = std::min(kMaxPack, M - m0);
        for (int i = 0; i < n; ++i) {
            HB_REQUIRE(models[m0 + i] != nullptr, "hb_mlp_forward_multi: null model %d", m0 + i);
            pk.p[i] = models[m0 + i]->p;
        }
        fold_first_layer_kernel<<<ceil_div(n * 2 * kHid, 8), 256, 0, st>>>(pk, m0, n, kLayout.s[0], w.Wst, w.bst);
        HB_LAUNCHED();
    }
    // 3. ONE stacked GEMM: hidden and gate pre-activations of every model
    if ((rc = linear_fwd(w.xhat, w.Wst, w.bst, w.hg, B, M * 2 * kHid, kIn, st, w.part))) return rc;
    const int64_t n_gate = (int64_t)B * M * kHid;
    const int gate_grid = (int)std::min<int64_t>(ceil_div64(n_gate, 256), 2368);
    gate_stacked_kernel<<<gate_grid, 256, 0, st>>>(w.hg, w.a, B, M);
    HB_LAUNCHED();
    // 4. the 96-wide remainder, batched over the models
    for (int m0 = 0; m0 < M; m0 += kMaxPack) {
        PtrPack pk;
        const int n = std::min(kMaxPack, M - m0);
        for (int i = 0; i < n; ++i) pk.p[i] = models[m0 + i]->p;
        float* a = w.a + (int64_t)m0 * B * kHid;
        float* o = w.o + (int64_t)m0 * B * kDim;
        float* u = w.u + (int64_t)m0 * B * kDim;
        float* h = w.h + (int64_t)m0 * B * kHid;
        float* g = w.g + (int64_t)m0 * B * kHid;
        const dim3 grid_o(ceil_div(kDim, kTile), ceil_div(B, kTile), n), grid_h(1, ceil_div(B, kTile), n);
        linear_batched_kernel<<<grid_o, 256, 0, st>>>(a, pk, kLayout.s[0].ow, kLayout.s[0].ob, o, B, kDim, kHid);
        HB_LAUNCHED();
        for (int s = 1; s < kStages; ++s) {
            const StageOff& L = kLayout.s[s];
            ln_fwd_batched_kernel<<<dim3(ceil_div(B, 8), n), 256, 0, st>>>(o, pk, L.ln_w, L.ln_b, u, B, kDim);
            HB_LAUNCHED();
            linear_batched_kernel<<<grid_h, 256, 0, st>>>(u, pk, L.hw, L.hb, h, B, kHid, kDim);
            HB_LAUNCHED();
            linear_batched_kernel<<<grid_h, 256, 0, st>>>(u, pk, L.gw, L.gb, g, B, kHid, kDim);
            HB_LAUNCHED();
            const int64_t ng = (int64_t)n * B * kHid;
            gate_fwd_kernel<<<(int)std::min<int64_t>(ceil_div64(ng, 256), 2368), 256, 0, st>>>(h, g, a, ng);
            HB_LAUNCHED();
            // stage 3's output is one logit per row: o is reused as [n][B][1]
            const dim3 grid_out(ceil_div(L.out_dim, kTile), ceil_div(B, kTile), n);
            linear_batched_kernel<<<grid_out, 256, 0, st>>>(a, pk, L.ow, L.ob, o, B, L.out_dim, kHid);
            HB_LAUNCHED();
        }
        const int total = n * B;   // logits [n][B] contiguous
        sigmoid_kernel<<<ceil_div(total, 256), 256, 0, st>>>(o, prob_dev + (int64_t)m0 * B, total);
        HB_LAUNCHED();
    }
    return HB_OK;
}

namespace hb {

// forward + sigmoid + high-loss selection: prob, stats[1] = rows selected in THIS batch (stats[0,2,3] zeroed)
static int select_impl(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float thr, float* prob_dev, float* stats_dev,
                       const Ws& w, cudaStream_t st) {
    int rc = forward_impl(m, x_dev, B, w, st);
    if (rc) return rc;
    HB_CUDA_OK(cudaMemsetAsync(stats_dev, 0, 4 * sizeof(float), st));
    head_select_kernel<<<ceil_div(B, 256), 256, 0, st>>>(w.o[3], y_dev, prob_dev, stats_dev, B, thr);
    HB_LAUNCHED();
    return HB_OK;
}

// weighted BCE over the selected rows divided by *n_total, and its gradient with respect to every parameter -> m->g
static int backward_impl(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float negative_weight, float thr,
                         const float* n_total_dev, int min_selected, const float* prob_dev, float* stats_dev, const Ws& w,
                         cudaStream_t st) {
    int rc;
    head_grad_kernel<<<ceil_div(B, 256), 256, 0, st>>>(prob_dev, y_dev, w.dz, stats_dev, n_total_dev, B, thr, negative_weight, m->loss_scale);
    HB_LAUNCHED();
    head_finish_kernel<<<1, 1, 0, st>>>(stats_dev, n_total_dev, B, min_selected);
    HB_LAUNCHED();
    const float* d_out = w.dz;   // gradient wrt the stage's output o[s]
    for (int s = kStages - 1; s >= 0; --s) {
        const StageOff& L = kLayout.s[s];
        const int in_dim = L.in_dim, out_dim = L.out_dim;
        // output linear: dW_o = d_out^T a, db_o = colsum(d_out), d_a = d_out W_o
        if ((rc = gemm(d_out, 1, out_dim, w.a[s], kHid, 1, m->g + L.ow, kHid, nullptr, out_dim, kHid, B, 0, st, w.part))) return rc;
        colsum_kernel<<<ceil_div(out_dim, 32), 256, 0, st>>>(d_out, m->g + L.ob, B, out_dim, 0);
        HB_LAUNCHED();
        if ((rc = gemm(d_out, out_dim, 1, m->p + L.ow, kHid, 1, w.d_a, kHid, nullptr, B, kHid, out_dim, 0, st))) return rc;
        const int64_t n = (int64_t)B * kHid;
        gate_bwd_kernel<<<(int)std::min<int64_t>(ceil_div64(n, 256), 1184), 256, 0, st>>>(w.h[s], w.g[s], w.d_a, w.d_h, w.d_g, n);
        HB_LAUNCHED();
        // hidden / gate linears: dW = d^T u, db = colsum(d), d_u = d_h W_h + d_g W_g
        if ((rc = gemm(w.d_h, 1, kHid, w.u[s], in_dim, 1, m->g + L.hw, in_dim, nullptr, kHid, in_dim, B, 0, st, w.part))) return rc;
        if ((rc = gemm(w.d_g, 1, kHid, w.u[s], in_dim, 1, m->g + L.gw, in_dim, nullptr, kHid, in_dim, B, 0, st, w.part))) return rc;
        colsum_kernel<<<ceil_div(kHid, 32), 256, 0, st>>>(w.d_h, m->g + L.hb, B, kHid, 0);
        HB_LAUNCHED();
        colsum_kernel<<<ceil_div(kHid, 32), 256, 0, st>>>(w.d_g, m->g + L.gb, B, kHid, 0);
        HB_LAUNCHED();
        if ((rc = gemm(w.d_h, kHid, 1, m->p + L.hw, in_dim, 1, w.d_u, in_dim, nullptr, B, in_dim, kHid, 0, st))) return rc;
        if ((rc = gemm(w.d_g, kHid, 1, m->p + L.gw, in_dim, 1, w.d_u, in_dim, nullptr, B, in_dim, kHid, 1, st))) return rc;
        // LayerNorm: parameter grads always, input grad unless the input is the data
        const float* ln_in = (s == 0) ? x_dev : w.o[s - 1];
        ln_bwd_params_kernel<<<ceil_div(in_dim, 32), 256, 0, st>>>(w.d_u, ln_in, w.mean[s], w.rstd[s], m->g + L.ln_w, m->g + L.ln_b, B, in_dim);
        HB_LAUNCHED();
        if (s > 0) {
            ln_bwd_input_kernel<<<ceil_div(B, 8), 256, 0, st>>>(w.d_u, ln_in, m->p + L.ln_w, w.mean[s], w.rstd[s], w.d_x, B, in_dim);
            HB_LAUNCHED();
            // d_x becomes the next d_out; keep it in d_o so d_x can be rewritten
            HB_CUDA_OK(cudaMemcpyAsync(w.d_o, w.d_x, (size_t)B * in_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
            d_out = w.d_o;
        }
    }
    return HB_OK;
}

static int adam_impl(hb_mlp_model* m, float lr, const float* stats_dev, cudaStream_t st) {
    adam_kernel<<<148, 256, 0, st>>>(m->p, m->g, m->m, m->v, m->step, stats_dev, lr, kLayout.total);
    HB_LAUNCHED();
    adam_step_inc_kernel<<<1, 1, 0, st>>>(m->step, stats_dev);
    HB_LAUNCHED();
    return HB_OK;
}

}  // namespace hb

extern "C" int hb_mlp_train_step(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float lr,
                                 float negative_weight, float high_loss_threshold, int min_selected, float* prob_dev,
                                 float* stats_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && x_dev && y_dev && prob_dev && stats_dev && workspace_dev, "hb_mlp_train_step: null pointer");
    HB_REQUIRE(B > 0, "hb_mlp_train_step: empty batch");
    HB_REQUIRE(workspace_bytes >= hb_mlp_workspace_bytes(B, 1), "hb_mlp_train_step: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    Ws w;
    carve(&w, reinterpret_cast<float*>(workspace_dev), B, 1);
    int rc;
    if ((rc = select_impl(m, x_dev, y_dev, B, high_loss_threshold, prob_dev, stats_dev, w, st))) return rc;
    // single device: the batch's own selection count is the normaliser
    if ((rc = backward_impl(m, x_dev, y_dev, B, negative_weight, high_loss_threshold, stats_dev + 1, min_selected, prob_dev, stats_dev, w, st)))
        return rc;
    return adam_impl(m, lr, stats_dev, st);
}

// ---- data-parallel form of the same step: select -> (all-reduce n) -> backward -> (all-reduce grads) -> adam ----
extern "C" int hb_mlp_select(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float high_loss_threshold,
                             float* prob_dev, float* stats_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && x_dev && y_dev && prob_dev && stats_dev && workspace_dev && B > 0, "hb_mlp_select: bad argument");
    HB_REQUIRE(workspace_bytes >= hb_mlp_workspace_bytes(B, 1), "hb_mlp_select: workspace too small");
    Ws w;
    carve(&w, reinterpret_cast<float*>(workspace_dev), B, 1);
    return select_impl(m, x_dev, y_dev, B, high_loss_threshold, prob_dev, stats_dev, w, (cudaStream_t)stream);
}

extern "C" int hb_mlp_backward(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float negative_weight,
                               float high_loss_threshold, const float* n_selected_total_dev, int min_selected,
                               const float* prob_dev, float* stats_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && x_dev && y_dev && prob_dev && stats_dev && workspace_dev && n_selected_total_dev && B > 0, "hb_mlp_backward: bad argument");
    HB_REQUIRE(workspace_bytes >= hb_mlp_workspace_bytes(B, 1), "hb_mlp_backward: workspace too small");
    Ws w;
    carve(&w, reinterpret_cast<float*>(workspace_dev), B, 1);
    return backward_impl(m, x_dev, y_dev, B, negative_weight, high_loss_threshold, n_selected_total_dev, min_selected, prob_dev,
                         stats_dev, w, (cudaStream_t)stream);
}

extern "C" int hb_mlp_grads_copy(hb_mlp_model* m, float* buf_dev, int64_t n_floats, int to_model, void* stream) {
    HB_REQUIRE(m && buf_dev && n_floats == kLayout.total, "hb_mlp_grads_copy: expected %d floats", kLayout.total);
    HB_CUDA_OK(cudaMemcpyAsync(to_model ? m->g : buf_dev, to_model ? buf_dev : m->g, (size_t)n_floats * sizeof(float),
                               cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return HB_OK;
}

extern "C" int hb_mlp_adam(hb_mlp_model* m, float lr, const float* stats_dev, void* stream) {
    HB_REQUIRE(m && stats_dev, "hb_mlp_adam: null pointer");
    return adam_impl(m, lr, stats_dev, (cudaStream_t)stream);
}
