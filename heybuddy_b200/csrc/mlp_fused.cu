// K8, fused: the wake-word classifier's forward and backward pass in a dozen launches instead of a hundred.
//
//   x[B,1536] -> LN -> gated(1536->64->96) -> 2 x [LN(96) -> gated(96->64->96)] -> LN(96) -> gated(96->64->1)
//   gated(u) = W_o (silu(W_h u + b_h) * (W_g u + b_g)) + b_o            (reference wakeword.py:334-348, multi_layer_perceptron.py:115-124)
//
// Everything behind the first-layer product is row-local with 55 k weights, so one CTA carries 32 rows through all of it
// (`fwd_tail_kernel`, `bwd_tail_kernel`: lane = row, warp = a group of output columns, activations as [feature][row] planes in shared
// memory, each weight matrix staged once per CTA and read as warp-uniform 16-byte broadcasts).
//
// The first layer is the only large product.  LN(x) = xhat gamma + beta, so gamma folds into the weights and beta into the bias
// (`prep_kernel`), the row statistics fold into the product's operand load, and the product runs on tcgen05 as three TF32 passes
// (gemm_tf32.cu) over K slices whose partial sums the tail kernel adds up.  The hidden and the gate rows are interleaved (row 2j = hidden
// j, 2j + 1 = gate j) so that one thread holds both halves of a gate.
//
// Backward: for every stage, with G = d_hg^T xhat  [128, in] and db = column sums of d_hg,
//     dW = gamma G + beta db      d gamma = sum_n W[n] G[n]      d beta = sum_n W[n] db[n]
// so the weight gradient, both LayerNorm parameter gradients and the bias gradient come from ONE product over the batch, and the
// 1536-wide input gradient of the first layer (a 1.6 GFLOP product plus a 25 MB round trip in the staged path) is never formed.
// G of the first layer runs on tcgen05 (batch-major operands, register transposes in the producers); the 96-wide ones and the
// output-linear gradients share one grouped FMA launch (`grads_kernel`); `finish_kernel` adds the K slices up in a fixed order and
// applies the formulas above.  Every sum has a fixed order: the step is deterministic.
#include "mlp_common.cuh"

#include <algorithm>
#include <math.h>

namespace hb {

namespace {

const MlpLayout kL = make_mlp_layout();

constexpr int kRows = 32;                 // rows per CTA of the tail kernels
constexpr int kHg = 2 * kHid;             // 128 interleaved hidden / gate columns
constexpr int kTailThreads = 256;

// transposed copies of the small weights for the forward tail (written by prep_kernel)
constexpr int kWoT0 = 0;                                  // stage s: [64][out_dim]  (stage 3: [64])
constexpr int kWoTStride = kHid * kDim;
constexpr int kWoT3 = 3 * kWoTStride;
constexpr int kWhgT1 = kWoT3 + kHid;                      // stages 1..3: [96][128] interleaved columns, then the interleaved bias [128]
constexpr int kWhgTStride = kDim * kHg + kHg;
constexpr int kWI1 = kWhgT1 + 3 * kWhgTStride;          // stages 1..3: [128][96] interleaved rows, as the backward tail reads them
constexpr int kWIStride = kHg * kDim;
constexpr int kWtFloats = kWI1 + 3 * kWIStride;

__host__ __device__ constexpr int wot_off(int s) { return s < 3 ? kWoT0 + s * kWoTStride : kWoT3; }
__host__ __device__ constexpr int whgt_off(int s) { return kWhgT1 + (s - 1) * kWhgTStride; }
__host__ __device__ constexpr int wi_off(int s) { return kWI1 + (s - 1) * kWIStride; }

// gradient partial sums of the grouped launch: per K slice, [G_1 | G_2 | G_3 | O_0 | O_1 | O_2 | O_3 | row sums]
constexpr int kGOff1 = 0;                                 // G_s (s = 1..3): [128][96]
constexpr int kGStride = kHg * kDim;
constexpr int kOOff0 = 3 * kGStride;                      // O_s (s = 0..2): [96][64];  O_3: [1][64]
constexpr int kOStride = kDim * kHid;
constexpr int kOOff3 = kOOff0 + 3 * kOStride;
constexpr int kRsHg0 = kOOff3 + kHid;                     // column sums of d_hg, stages 0..3: [128] each
constexpr int kRsO0 = kRsHg0 + 4 * kHg;                   // column sums of d_o, stages 0..2: [96] each; stage 3 (dz): [1]
constexpr int kPartFloatsPerSlice = ((kRsO0 + 3 * kDim + 1 + 63) / 64) * 64;

constexpr int kGradSlicesMax = 16;
constexpr int kG0SlicesMax = 16;
constexpr int kFwdSlicesMax = 8;

int fwd_slices(int B) {
    const int tiles = ceil_div(B, 128);
    return std::max(1, std::min(kFwdSlicesMax, 148 / tiles));
}
int grad_slices(int B) { return std::max(1, std::min(kGradSlicesMax, B / 64)); }
int g0_slices(int B) { return std::max(1, std::min(std::min(kG0SlicesMax, 12), ceil_div(B, 32) / 4)); }

struct FusedWs {
    float *w0f, *b0f, *wt, *mean0, *rstd0, *hgpart;
    float *hg[kStages], *a[kStages], *xh[kStages], *rstd[kStages], *logit;
    float *dhg[kStages], *dout[kStages - 1], *part, *g0part;
};

int64_t carve(FusedWs* w, float* base, int B, int training) {
    int64_t off = 0;
    auto take = [&](int64_t n) { float* p = base ? base + off : nullptr; off += (n + 63) & ~63ll; return p; };
    w->w0f = take((int64_t)kHg * kIn);
    w->b0f = take(kHg);
    w->wt = take(kWtFloats);
    w->mean0 = take(B);
    w->rstd0 = take(B);
    w->hgpart = take((int64_t)fwd_slices(B) * B * kHg);
    w->logit = take(B);
    for (int s = 0; s < kStages; ++s) {
        // forward-only calls keep nothing: the buffers stay unallocated
        w->hg[s] = training ? take((int64_t)B * kHg) : nullptr;
        w->a[s] = training ? take((int64_t)B * kHid) : nullptr;
        w->xh[s] = (training && s > 0) ? take((int64_t)B * kDim) : nullptr;
        w->rstd[s] = (training && s > 0) ? take(B) : nullptr;
    }
    if (training) {
        for (int s = 0; s < kStages; ++s) w->dhg[s] = take((int64_t)B * kHg);
        for (int s = 0; s < kStages - 1; ++s) w->dout[s] = take((int64_t)B * kDim);
        w->part = take((int64_t)grad_slices(B) * kPartFloatsPerSlice);
        w->g0part = take((int64_t)g0_slices(B) * kHg * kIn);
    }
    return off;
}

struct Stages { StageOff s[kStages]; };
Stages stages() { Stages t; for (int i = 0; i < kStages; ++i) t.s[i] = kL.s[i]; return t; }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- prep: first-layer fold (interleaved rows) and the transposed small weights ---------------------------------------------------------
constexpr int kPrepFoldBlocks = kHg / 8;      // 8 warps per block, one stacked row per warp
__global__ void __launch_bounds__(256) prep_kernel(const float* __restrict__ p, const Stages L, float* __restrict__ w0f, float* __restrict__ b0f,
                                                   float* __restrict__ wt) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (blockIdx.x < kPrepFoldBlocks) {
        const int n = blockIdx.x * 8 + (tid >> 5), j = n >> 1, gate = n & 1;
        const StageOff& S = L.s[0];
        const float* W = p + (gate ? S.gw : S.hw) + j * kIn;
        const float* gamma = p + S.ln_w;
        const float* beta = p + S.ln_b;
        float acc = 0.f;
        for (int k = lane; k < kIn; k += 32) {
            const float w = W[k];
            w0f[n * kIn + k] = w * gamma[k];
            acc = fmaf(w, beta[k], acc);
        }
        acc = warp_sum_f(acc);
        if (lane == 0) b0f[n] = acc + p[(gate ? S.gb : S.hb) + j];
        return;
    }
    const int t0 = (blockIdx.x - kPrepFoldBlocks) * 256 + tid, stride = (gridDim.x - kPrepFoldBlocks) * 256;
    for (int s = 0; s < kStages; ++s) {
        const StageOff& S = L.s[s];
        // WoT[j][n] = Wo[n][j]
        for (int i = t0; i < kHid * S.out_dim; i += stride) {
            const int j = i / S.out_dim, n = i - j * S.out_dim;
            wt[wot_off(s) + i] = p[S.ow + n * kHid + j];
        }
        if (s == 0) continue;
        // WhgT[k][2 j + gate] = (gate ? Wg : Wh)[j][k], bias likewise
        float* dst = wt + whgt_off(s);
        for (int i = t0; i < kDim * kHg; i += stride) {
            const int k = i / kHg, n = i - k * kHg, j = n >> 1;
            dst[i] = p[((n & 1) ? S.gw : S.hw) + j * kDim + k];
        }
        for (int n = t0; n < kHg; n += stride) dst[kDim * kHg + n] = p[((n & 1) ? S.gb : S.hb) + (n >> 1)];
        for (int i = t0; i < kHg * kDim; i += stride) {
            const int n = i / kDim, k = i - n * kDim;
            wt[wi_off(s) + i] = p[((n & 1) ? S.gw : S.hw) + (n >> 1) * kDim + k];
        }
    }
}

// mean and 1/std of every input row (LayerNorm(1536) without the affine), one warp per row, the row read once
__global__ void __launch_bounds__(256) rowstats_kernel(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ rstd, int B) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)row * kIn);
    float4 v[kIn / 128];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kIn / 128; ++i) { v[i] = __ldg(xr + lane + 32 * i); s += (v[i].x + v[i].y) + (v[i].z + v[i].w); }
    const float mu = warp_sum_f(s) * (1.f / kIn);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kIn / 128; ++i) {
        const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rs = rsqrtf(warp_sum_f(q) * (1.f / kIn) + kLnEps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

// ---- tail kernels: shared helpers ------------------------------------------------------------------------------------------------------
// acc[j] += sum_k act[k][lane] W[k][col0 + j]      (act: [K][32] plane, W: [K][pitch] in shared memory, 16-byte broadcast reads)
template <int K, int NJ>
__device__ __forceinline__ void plane_fma(const float* __restrict__ act, const float* __restrict__ W, int pitch, int col0, int lane, float (&acc)[NJ]) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float a = act[k * kRows + lane];
        const float4* w4 = reinterpret_cast<const float4*>(W + k * pitch + col0);
#pragma unroll
        for (int q = 0; q < NJ / 4; ++q) {
            const float4 w = w4[q];
            acc[4 * q] = fmaf(a, w.x, acc[4 * q]);
            acc[4 * q + 1] = fmaf(a, w.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(a, w.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(a, w.w, acc[4 * q + 3]);
        }
    }
}

__device__ __forceinline__ void stage_copy(float* __restrict__ dst, const float* __restrict__ src, int n_floats, int tid) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < n_floats / 4; i += kTailThreads) d4[i] = __ldg(s4 + i);
}

template <int N>
__device__ __forceinline__ void load_row(float (&v)[N], const float* __restrict__ src, bool ok) {
#pragma unroll
    for (int q = 0; q < N / 4; ++q) {
        const float4 t = ok ? *reinterpret_cast<const float4*>(src + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}
template <int N>
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float (&v)[N], bool ok) {
    if (!ok) return;
#pragma unroll
    for (int q = 0; q < N / 4; ++q) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// sum over the 8 warps of a per-(warp, lane) value: red is [8][32]; the caller separates uses of the same array by a barrier
__device__ __forceinline__ float cross_warp_sum(float* red, float v, int warp, int lane) {
    red[warp * kRows + lane] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w * kRows + lane];
    return s;
}

struct FwdArgs {
    const float* p; const float* wt; const float* b0f; const float* hgpart; int splits; int64_t split_stride;
    float* hg[kStages]; float* a[kStages]; float* xh[kStages]; float* rstd[kStages]; float* logit;
    int B; int save;
    Stages L;
};

constexpr int kTailSmemFloats = kHg * kDim + kHg * kRows + kDim * kRows + 4 * 8 * kRows;     // weights | wide plane | narrow plane | reductions

// first-layer pre-activations (K slices) -> logits
__global__ void __launch_bounds__(kTailThreads) fwd_tail_kernel(const FwdArgs f) {
    extern __shared__ __align__(16) float sm[];
    float* wbuf = sm;                              // one weight matrix at a time (up to [96][128])
    float* abuf = wbuf + kHg * kDim;               // gated activations [64][32]
    float* ubuf = abuf + kHg * kRows;              // stage input [96][32]
    float* red = ubuf + kDim * kRows;              // [4][8][32]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = blockIdx.x * kRows + lane;
    const bool ok = row < f.B;
    const bool save = ok && f.save;

    for (int s = 0; s < kStages; ++s) {
        const StageOff& S = f.L.s[s];
        float hgv[16];
        if (s == 0) {
            load_row(hgv, f.b0f + warp * 16, true);
            for (int z = 0; z < f.splits; ++z) {
                float t[16];
                load_row(t, f.hgpart + z * f.split_stride + (int64_t)row * kHg + warp * 16, ok);
#pragma unroll
                for (int e = 0; e < 16; ++e) hgv[e] += t[e];
            }
        } else {
            __syncthreads();                       // every warp is done with the previous matrix; ubuf is complete
            stage_copy(wbuf, f.wt + whgt_off(s), kDim * kHg, tid);
            __syncthreads();
            load_row(hgv, f.wt + whgt_off(s) + kDim * kHg + warp * 16, true);
            plane_fma<kDim, 16>(ubuf, wbuf, kHg, warp * 16, lane, hgv);
        }
        if (save) store_row(f.hg[s] + (int64_t)row * kHg + warp * 16, hgv, true);
        float av[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float h = hgv[2 * j], g = hgv[2 * j + 1];
            av[j] = h / (1.f + expf(-h)) * g;
            abuf[(warp * 8 + j) * kRows + lane] = av[j];
        }
        if (save) store_row(f.a[s] + (int64_t)row * kHid + warp * 8, av, true);
        __syncthreads();                           // abuf complete; wbuf free
        if (s < kStages - 1) {
            stage_copy(wbuf, f.wt + wot_off(s), kHid * kDim, tid);
            __syncthreads();
            float ov[12];
            load_row(ov, f.p + S.ob + warp * 12, true);
            plane_fma<kHid, 12>(abuf, wbuf, kDim, warp * 12, lane, ov);
            // LayerNorm of the next stage over the row's 96 features (12 per warp)
            const StageOff& N = f.L.s[s + 1];
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 12; ++j) ps += ov[j];
            const float mu = cross_warp_sum(red + (s & 1) * 2 * 8 * kRows, ps, warp, lane) * (1.f / kDim);
            float pq = 0.f;
#pragma unroll
            for (int j = 0; j < 12; ++j) { const float d = ov[j] - mu; pq = fmaf(d, d, pq); }
            const float rs = rsqrtf(cross_warp_sum(red + (s & 1) * 2 * 8 * kRows + 8 * kRows, pq, warp, lane) * (1.f / kDim) + kLnEps);
            float xh[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                xh[j] = (ov[j] - mu) * rs;
                ubuf[(warp * 12 + j) * kRows + lane] = fmaf(xh[j], __ldg(f.p + N.ln_w + warp * 12 + j), __ldg(f.p + N.ln_b + warp * 12 + j));
            }
            if (save) {
                store_row(f.xh[s + 1] + (int64_t)row * kDim + warp * 12, xh, true);
                if (warp == 0) f.rstd[s + 1][row] = rs;
            }
        } else {
            float ps = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) ps = fmaf(av[j], __ldg(f.p + S.ow + warp * 8 + j), ps);
            const float z = cross_warp_sum(red + (s & 1) * 2 * 8 * kRows, ps, warp, lane) + __ldg(f.p + S.ob);
            if (ok && warp == 0) f.logit[row] = z;
        }
    }
}

struct BwdArgs {
    const float* p; const float* wt; const float* dz;
    const float* hg[kStages]; const float* xh[kStages]; const float* rstd[kStages];
    float* dhg[kStages]; float* dout[kStages - 1];
    int B;
    Stages L;
};

// d loss / d logit -> d_hg of every stage and d_o of stages 0..2 (what the weight-gradient products read)
__global__ void __launch_bounds__(kTailThreads) bwd_tail_kernel(const BwdArgs f) {
    extern __shared__ __align__(16) float sm[];
    float* wbuf = sm;                              // [128][96] interleaved hidden / gate rows, or W_o [96][64]
    float* dbuf = wbuf + kHg * kDim;               // d_hg plane [128][32]
    float* dobuf = dbuf + kHg * kRows;             // d_o plane [96][32]
    float* red = dobuf + kDim * kRows;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = blockIdx.x * kRows + lane;
    const bool ok = row < f.B;

    for (int s = kStages - 1; s >= 0; --s) {
        const StageOff& S = f.L.s[s];
        float dav[8];
        if (s == kStages - 1) {
            const float dzr = ok ? f.dz[row] : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) dav[j] = dzr * __ldg(f.p + S.ow + warp * 8 + j);
        } else {
            __syncthreads();                       // dobuf complete; every warp is done with the previous matrix
            stage_copy(wbuf, f.p + S.ow, kDim * kHid, tid);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; ++j) dav[j] = 0.f;
            plane_fma<kDim, 8>(dobuf, wbuf, kHid, warp * 8, lane, dav);
        }
        float hgv[16], dhv[16];
        load_row(hgv, f.hg[s] + (int64_t)row * kHg + warp * 16, ok);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float h = hgv[2 * j], g = hgv[2 * j + 1];
            const float sg = 1.f / (1.f + expf(-h));
            dhv[2 * j] = dav[j] * g * (sg * (1.f + h * (1.f - sg)));      // d hidden
            dhv[2 * j + 1] = dav[j] * h * sg;                             // d gate
        }
        store_row(f.dhg[s] + (int64_t)row * kHg + warp * 16, dhv, ok);
        if (s == 0) break;
#pragma unroll
        for (int e = 0; e < 16; ++e) dbuf[(warp * 16 + e) * kRows + lane] = dhv[e];
        __syncthreads();                           // dbuf complete; wbuf free
        stage_copy(wbuf, f.wt + wi_off(s), kHg * kDim, tid);
        __syncthreads();
        float duv[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) duv[j] = 0.f;
        plane_fma<kHg, 12>(dbuf, wbuf, kDim, warp * 12, lane, duv);
        // LayerNorm backward (input of stage s = output of stage s - 1)
        float xh[12];
        load_row(xh, f.xh[s] + (int64_t)row * kDim + warp * 12, ok);
        float p1 = 0.f, p2 = 0.f;
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            duv[j] *= __ldg(f.p + S.ln_w + warp * 12 + j);
            p1 += duv[j];
            p2 = fmaf(duv[j], xh[j], p2);
        }
        const float m1 = cross_warp_sum(red + (s & 1) * 2 * 8 * kRows, p1, warp, lane) * (1.f / kDim);
        const float m2 = cross_warp_sum(red + (s & 1) * 2 * 8 * kRows + 8 * kRows, p2, warp, lane) * (1.f / kDim);
        const float rs = ok ? f.rstd[s][row] : 0.f;
        float dov[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            dov[j] = rs * (duv[j] - m1 - xh[j] * m2);
            dobuf[(warp * 12 + j) * kRows + lane] = dov[j];
        }
        store_row(f.dout[s - 1] + (int64_t)row * kDim + warp * 12, dov, ok);
    }
}

// ---- grouped weight-gradient products: C_p[M, N] = sum_k A_p[k][m] B_p[k][n], plus the column sums of A_p -----------------------------
constexpr int kGradProbs = 8;
struct GradProb {
    const float* A; const float* B;
    int M, N, lda, ldb;
    int part_off, rs_off;       // into a slice of the partial buffer
    int tile0, tiles_n;
};
struct GradArgs {
    GradProb pr[kGradProbs];
    int nprob, K, k_per_slice;
    float* part;
};

__global__ void __launch_bounds__(256) grads_kernel(const GradArgs g) {
    __shared__ __align__(16) float As[2][16][64];
    __shared__ __align__(16) float Bs[2][16][64];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int pi = 0;
    while (pi + 1 < g.nprob && (int)blockIdx.x >= g.pr[pi + 1].tile0) ++pi;
    const GradProb& P = g.pr[pi];
    const int t = blockIdx.x - P.tile0, tm = t / P.tiles_n, tn = t - tm * P.tiles_n;
    const int m0 = tm * 64, n0 = tn * 64;
    const int k_begin = blockIdx.y * g.k_per_slice, k_end = min(g.K, k_begin + g.k_per_slice);
    const int lk = tid >> 4, lc = (tid & 15) * 4;          // this thread's load slot: row lk of the 16-row chunk, 4 columns from lc
    auto fetch = [&](const float* X, int ld, int dim, int c0, int k) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (X == nullptr || k >= k_end) return v;
        const float* src = X + (int64_t)k * ld + c0 + lc;
        const int left = dim - (c0 + lc);
        if (left >= 4 && (ld & 3) == 0) return __ldg(reinterpret_cast<const float4*>(src));
        if (left > 0) v.x = __ldg(src);
        if (left > 1) v.y = __ldg(src + 1);
        if (left > 2) v.z = __ldg(src + 2);
        if (left > 3) v.w = __ldg(src + 3);
        return v;
    };
    float acc[4][4] = {};
    float rsum[4] = {0.f, 0.f, 0.f, 0.f};
    const bool do_rs = tn == 0 && tx == 0;
    float4 va = fetch(P.A, P.lda, P.M, m0, k_begin + lk), vb = fetch(P.B, P.ldb, P.N, n0, k_begin + lk);
    int buf = 0;
    for (int k0 = k_begin; k0 < k_end; k0 += 16) {
        *reinterpret_cast<float4*>(&As[buf][lk][lc]) = va;
        *reinterpret_cast<float4*>(&Bs[buf][lk][lc]) = vb;
        __syncthreads();
        if (k0 + 16 < k_end) { va = fetch(P.A, P.lda, P.M, m0, k0 + 16 + lk); vb = fetch(P.B, P.ldb, P.N, n0, k0 + 16 + lk); }
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                if (do_rs) rsum[i] += av[i];
            }
        }
        buf ^= 1;       // the next chunk goes to the other buffer: one barrier per chunk is enough
    }
    float* out = g.part + (int64_t)blockIdx.y * kPartFloatsPerSlice;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= P.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < P.N) out[P.part_off + m * P.N + n] = acc[i][j];
        }
        if (do_rs) out[P.rs_off + m] = rsum[i];
    }
}

// ---- finish: add the K slices up and turn G, db into the parameter gradients ---------------------------------------------------------------
struct FinArgs {
    const float* p; float* g;
    const float* part; int slices;
    const float* g0part; int g0_slices;
    Stages L;
};
constexpr int kFinColBlocks = kIn / 32 + 3 * (kDim / 32);    // 32 input columns of one stage per block
constexpr int kFinOutBlocks = 8;

__global__ void __launch_bounds__(256) finish_kernel(const FinArgs f) {
    __shared__ float db[kHg];
    __shared__ float red[2][8][32];
    const int tid = threadIdx.x;
    if (blockIdx.x >= kFinColBlocks) {
        // output linears: dW_o = O, db_o = column sums of d_o
        const int t0 = (blockIdx.x - kFinColBlocks) * 256 + tid, stride = kFinOutBlocks * 256;
        for (int s = 0; s < kStages; ++s) {
            const StageOff& S = f.L.s[s];
            const int off = s < 3 ? kOOff0 + s * kOStride : kOOff3, rs_off = kRsO0 + s * kDim;
            for (int i = t0; i < S.out_dim * kHid; i += stride) {
                float v = 0.f;
                for (int z = 0; z < f.slices; ++z) v += f.part[(int64_t)z * kPartFloatsPerSlice + off + i];
                f.g[S.ow + i] = v;
            }
            for (int n = t0; n < S.out_dim; n += stride) {
                float v = 0.f;
                for (int z = 0; z < f.slices; ++z) v += f.part[(int64_t)z * kPartFloatsPerSlice + rs_off + n];
                f.g[S.ob + n] = v;
            }
        }
        return;
    }
    int s, kb;
    if ((int)blockIdx.x < kIn / 32) { s = 0; kb = blockIdx.x; }
    else { s = 1 + (blockIdx.x - kIn / 32) / (kDim / 32); kb = (blockIdx.x - kIn / 32) % (kDim / 32); }
    const StageOff& S = f.L.s[s];
    const int in = S.in_dim;
    if (tid < kHg) {
        float v = 0.f;
        for (int z = 0; z < f.slices; ++z) v += f.part[(int64_t)z * kPartFloatsPerSlice + kRsHg0 + s * kHg + tid];
        db[tid] = v;
        if (kb == 0) f.g[((tid & 1) ? S.gb : S.hb) + (tid >> 1)] = v;
    }
    __syncthreads();
    const int kx = tid & 31, ny = tid >> 5, k = kb * 32 + kx;
    const float gamma = f.p[S.ln_w + k], beta = f.p[S.ln_b + k];
    const float* G = s == 0 ? f.g0part : f.part + kGOff1 + (s - 1) * kGStride;
    const int nz = s == 0 ? f.g0_slices : f.slices;
    const int64_t zstride = s == 0 ? (int64_t)kHg * kIn : kPartFloatsPerSlice;
    float dgam = 0.f, dbet = 0.f;
    for (int n = ny; n < kHg; n += 8) {
        float Gv = 0.f;
        for (int z = 0; z < nz; ++z) Gv += G[z * zstride + (int64_t)n * in + k];
        const int widx = ((n & 1) ? S.gw : S.hw) + (n >> 1) * in + k;
        f.g[widx] = fmaf(gamma, Gv, beta * db[n]);
        const float w = f.p[widx];
        dgam = fmaf(w, Gv, dgam);
        dbet = fmaf(w, db[n], dbet);
    }
    red[0][ny][kx] = dgam;
    red[1][ny][kx] = dbet;
    __syncthreads();
    if (ny == 0) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a += red[0][w][kx]; b += red[1][w][kx]; }
        f.g[S.ln_w + k] = a;
        f.g[S.ln_b + k] = b;
    }
}

int configure_tail_smem() {
    static bool done = false;
    if (!done) {
        const int bytes = kTailSmemFloats * (int)sizeof(float);
        HB_CUDA_OK(cudaFuncSetAttribute(fwd_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        HB_CUDA_OK(cudaFuncSetAttribute(bwd_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done = true;
    }
    return HB_OK;
}

}  // namespace

int64_t mlp_fused_ws_floats(int B, int training) {
    FusedWs w;
    return carve(&w, nullptr, B, training);
}

int mlp_fused_forward(const hb_mlp_model* m, const float* x, int B, float* ws, int training, const float** logits, cudaStream_t st) {
    HB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0, "classifier: 16-byte aligned buffers");
    int rc;
    if ((rc = configure_tail_smem())) return rc;
    FusedWs w;
    carve(&w, ws, B, training);
    const Stages L = stages();
    prep_kernel<<<kPrepFoldBlocks + 32, 256, 0, st>>>(m->p, L, w.w0f, w.b0f, w.wt);
    HB_LAUNCHED();
    rowstats_kernel<<<ceil_div(B, 8), 256, 0, st>>>(x, w.mean0, w.rstd0, B);
    HB_LAUNCHED();
    TfArgs t;
    t.A = x; t.lda = kIn;
    t.B0 = w.w0f; t.B1 = w.w0f; t.bsplit = 1 << 30; t.ldb = kIn;
    t.bias0 = nullptr; t.bias1 = nullptr; t.biassplit = 1 << 30;
    t.C = w.hgpart; t.ldc = kHg;
    t.M = B; t.N = kHg; t.K = kIn;
    t.mean = w.mean0; t.rstd = w.rstd0;
    t.splits = fwd_slices(B); t.split_stride = (int64_t)B * kHg;
    if ((rc = gemm_tf32x3_launch(t, st))) return rc;
    FwdArgs f;
    f.p = m->p; f.wt = w.wt; f.b0f = w.b0f; f.hgpart = w.hgpart; f.splits = t.splits; f.split_stride = t.split_stride;
    for (int s = 0; s < kStages; ++s) { f.hg[s] = w.hg[s]; f.a[s] = w.a[s]; f.xh[s] = w.xh[s]; f.rstd[s] = w.rstd[s]; }
    f.logit = w.logit;
    f.B = B; f.save = training;
    f.L = L;
    fwd_tail_kernel<<<ceil_div(B, kRows), kTailThreads, kTailSmemFloats * sizeof(float), st>>>(f);
    HB_LAUNCHED();
    *logits = w.logit;
    return HB_OK;
}

int mlp_fused_backward(hb_mlp_model* m, const float* x, int B, float* ws, const float* dz, cudaStream_t st) {
    int rc;
    if ((rc = configure_tail_smem())) return rc;
    FusedWs w;
    carve(&w, ws, B, 1);
    const Stages L = stages();
    BwdArgs b;
    b.p = m->p; b.wt = w.wt; b.dz = dz;
    for (int s = 0; s < kStages; ++s) { b.hg[s] = w.hg[s]; b.xh[s] = w.xh[s]; b.rstd[s] = w.rstd[s]; b.dhg[s] = w.dhg[s]; }
    for (int s = 0; s < kStages - 1; ++s) b.dout[s] = w.dout[s];
    b.B = B;
    b.L = L;
    bwd_tail_kernel<<<ceil_div(B, kRows), kTailThreads, kTailSmemFloats * sizeof(float), st>>>(b);
    HB_LAUNCHED();

    // first layer: G_0 = d_hg0^T xhat_0 on the tensor cores, xhat formed from x and the row statistics in the operand load
    TfArgs t;
    t.A = w.dhg[0]; t.lda = kHg;
    t.B0 = x; t.B1 = nullptr; t.bsplit = 1 << 30; t.ldb = kIn;
    t.bias0 = nullptr; t.bias1 = nullptr; t.biassplit = 1 << 30;
    t.C = w.g0part; t.ldc = kIn;
    t.M = kHg; t.N = kIn; t.K = B;
    t.mean = w.mean0; t.rstd = w.rstd0;
    t.splits = g0_slices(B); t.split_stride = (int64_t)kHg * kIn;
    if ((rc = gemm_tf32x3_launch_batch_major(t, st))) return rc;

    GradArgs g;
    int np = 0, tile = 0;
    auto add = [&](const float* A, int lda, int M, const float* Bm, int ldb, int N, int part_off, int rs_off) {
        GradProb& P = g.pr[np++];
        P.A = A; P.B = Bm; P.M = M; P.N = N; P.lda = lda; P.ldb = ldb; P.part_off = part_off; P.rs_off = rs_off;
        P.tile0 = tile; P.tiles_n = std::max(1, ceil_div(N, 64));
        tile += ceil_div(M, 64) * P.tiles_n;
    };
    add(w.dhg[0], kHg, kHg, nullptr, 0, 0, 0, kRsHg0);                                                  // column sums only
    for (int s = 1; s < kStages; ++s) add(w.dhg[s], kHg, kHg, w.xh[s], kDim, kDim, kGOff1 + (s - 1) * kGStride, kRsHg0 + s * kHg);
    for (int s = 0; s < kStages - 1; ++s) add(w.dout[s], kDim, kDim, w.a[s], kHid, kHid, kOOff0 + s * kOStride, kRsO0 + s * kDim);
    add(dz, 1, 1, w.a[3], kHid, kHid, kOOff3, kRsO0 + 3 * kDim);
    g.nprob = np;
    g.K = B;
    const int slices = grad_slices(B);
    g.k_per_slice = ceil_div(ceil_div(B, slices), 16) * 16;
    g.part = w.part;
    const int used = ceil_div(B, g.k_per_slice);
    grads_kernel<<<dim3(tile, used), 256, 0, st>>>(g);
    HB_LAUNCHED();

    FinArgs fin;
    fin.p = m->p; fin.g = m->g;
    fin.part = w.part; fin.slices = used;
    fin.g0part = w.g0part; fin.g0_slices = t.splits;
    fin.L = L;
    finish_kernel<<<kFinColBlocks + kFinOutBlocks, 256, 0, st>>>(fin);
    HB_LAUNCHED();
    return HB_OK;
}

}  // namespace hb
