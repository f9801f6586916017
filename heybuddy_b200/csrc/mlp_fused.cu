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
#include "tc_ptx.cuh"

#include <algorithm>
#include <math.h>

namespace hb {

namespace {

const MlpLayout kL = make_mlp_layout();

constexpr int kRows = 32;                 // rows per CTA of the tail kernels
constexpr int kHg = 2 * kHid;             // 128 interleaved hidden / gate columns
constexpr int kTailThreads = 128;
constexpr int kNjHg = kHg / 16, kNjA = kHid / 16, kNjO = kDim / 16;     // columns per thread: 8 pre-activations (4 gates), 4, 6

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

// gradient partial sums of the grouped launch's small products: per K slice, 8 slots of [128][128]
//   slot 0: column 0 = column sums of d_hg0;  slots 1..3: G_s [128][96] and, in column 96, the column sums of d_hg_s;
//   slots 4..6: O_s [96][64] (stages 0..2) and, in column 64, the column sums of d_o_s;  slot 7: O_3 [1][64], column 64 = sum of dz
constexpr int kSlot = 128 * 128;
constexpr int kSmallFloatsPerSlice = 8 * kSlot;
constexpr int kGradSlicesMax = 7;
constexpr int kFwdSlicesMax = 8;

int fwd_slices(int B) {
    const int tiles = ceil_div(B, 128);
    return std::max(1, std::min(kFwdSlicesMax, 148 / tiles));
}
int grad_slices(int B) { return std::max(1, std::min(kGradSlicesMax, ceil_div(B, 32) / 4)); }     // 20 tiles x 7 slices = 140 CTAs

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
        w->part = take((int64_t)grad_slices(B) * kSmallFloatsPerSlice);
        w->g0part = take((int64_t)grad_slices(B) * kHg * kIn);
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
constexpr int kPrepFoldBlocks = kHg / 2;      // two stacked rows per block, four warps (a quarter of the 1536 columns each) per row
constexpr int kPrepBlocks = kPrepFoldBlocks + 96;
__device__ __forceinline__ void prep_body(const int block, const float* __restrict__ p, const Stages& L, float* __restrict__ w0f,
                                          float* __restrict__ b0f, float* __restrict__ wt) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (block < kPrepFoldBlocks) {
        __shared__ float part[8];
        const int warp = tid >> 5, n = block * 2 + (warp >> 2), j = n >> 1, gate = n & 1, k0 = (warp & 3) * (kIn / 4);
        const StageOff& S = L.s[0];
        const float* W = p + (gate ? S.gw : S.hw) + j * kIn;
        const float* gamma = p + S.ln_w;
        const float* beta = p + S.ln_b;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < kIn / 4 / 32; ++i) {
            const int k = k0 + lane + 32 * i;
            const float w = W[k];
            w0f[n * kIn + k] = w * gamma[k];
            acc = fmaf(w, beta[k], acc);
        }
        acc = warp_sum_f(acc);
        if (lane == 0) part[warp] = acc;
        __syncthreads();
        if ((warp & 3) == 0 && lane == 0) b0f[n] = ((part[warp] + part[warp + 1]) + (part[warp + 2] + part[warp + 3])) + p[(gate ? S.gb : S.hb) + j];
        return;
    }
    const int t0 = (block - kPrepFoldBlocks) * 256 + tid, stride = (kPrepBlocks - kPrepFoldBlocks) * 256;
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
__device__ __forceinline__ void rowstats_body(const int block, const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ rstd, int B) {
    const int row = block * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
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

// one launch for both: the parameter-only blocks first (they are the long ones), then the row statistics
__global__ void __launch_bounds__(256) pre_kernel(const float* __restrict__ p, const Stages L, float* __restrict__ w0f, float* __restrict__ b0f,
                                                  float* __restrict__ wt, const float* __restrict__ x, float* __restrict__ mean,
                                                  float* __restrict__ rstd, int B) {
    if (blockIdx.x < kPrepBlocks) prep_body(blockIdx.x, p, L, w0f, b0f, wt);
    else rowstats_body(blockIdx.x - kPrepBlocks, x, mean, rstd, B);
}

// the transposed small weights of n models (blockIdx.y = model) for the multi-model forward tail
struct ModelList { const float* p[kMlpMaxPack]; };
__global__ void __launch_bounds__(256) prep_multi_kernel(const ModelList ml, const Stages L, float* __restrict__ wt_all) {
    prep_body(kPrepFoldBlocks + blockIdx.x, ml.p[blockIdx.y], L, nullptr, nullptr, wt_all + (int64_t)blockIdx.y * kWtFloats);
}

// ---- tail kernels ----------------------------------------------------------------------------------------------------------------------
// One CTA = 32 rows; 128 threads, thread (rq, cg) = rows 4 rq .. 4 rq + 3 x one of 16 column groups: a 4 x NJ register tile per product
// (NJ = 8 pre-activations = 4 gates, 6 outputs, 4 gated activations), so each 16-byte shared-memory read feeds 16 - 32 FMAs (a 1 x NJ tile
// per thread ran at a quarter of this: one weight word per FMA is all the shared-memory return path delivers).  Activations are
// [feature][row] planes; the weight matrices arrive by cp.async.bulk into two buffers, matrix i + 1 in flight while matrix i is used.
constexpr int kMatFloats = kDim * kHg + kHg;              // the largest staged matrix: [96][128] + its bias
constexpr int kRed = 16 * kRows;
constexpr int kSmallMatFloats = kDim * kHid;              // the other matrices: [64][96] / [96][64]; large and small ones alternate in both kernels
constexpr int kTailSmemFloats = kMatFloats + kSmallMatFloats + kHg * kRows + kDim * kRows + 2 * kRed;      // 107 KB: two CTAs per SM
constexpr int kTailSmemBytes = 128 + kTailSmemFloats * 4;

struct MatList { const float* src[6]; int floats[6]; };

struct WeightPipe {
    uint64_t* bar;      // [2]
    float* buf[2];      // even / odd matrices of the list (one holds the large ones, the other the small ones)
    __device__ __forceinline__ void issue(const MatList& ml, int i) {
        mbar_expect_tx(&bar[i & 1], (uint32_t)ml.floats[i] * 4u);
        bulk_g2s(buf[i & 1], ml.src[i], (uint32_t)ml.floats[i] * 4u, &bar[i & 1]);
    }
    // matrix i is in its buffer and everything every thread wrote to shared memory before this call is visible; matrix i - 1 is no longer
    // needed by anyone, so its buffer takes matrix i + 1
    __device__ __forceinline__ const float* acquire(const MatList& ml, int i, int n) {
        if (threadIdx.x < 32) mbar_wait(&bar[i & 1], (uint32_t)(i >> 1) & 1u, 16u + i);
        __syncthreads();
        if (threadIdx.x == 0 && i >= 1 && i + 1 < n) issue(ml, i + 1);
        return buf[i & 1];
    }
};

// acc[i][j] += sum_k act[k][r0 + i] W[k][col0 + j]
template <int K, int NJ>
__device__ __forceinline__ void plane_fma4(const float* __restrict__ act, const float* __restrict__ W, int pitch, int col0, int r0, float (&acc)[4][NJ]) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(act + k * kRows + r0);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        float w[NJ];
        if (NJ % 4 == 0) {
#pragma unroll
            for (int q = 0; q < NJ / 4; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(W + k * pitch + col0 + 4 * q);
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < NJ / 2; ++q) {
                const float2 t = *reinterpret_cast<const float2*>(W + k * pitch + col0 + 2 * q);
                w[2 * q] = t.x; w[2 * q + 1] = t.y;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
}

// sigmoid / silu on the exp2 unit (ex2.approx, rcp.approx: ~2 ulp each; the step's tolerances are 1e-3 on logits, 2e-3 on gradients)
__device__ __forceinline__ float sigmoid_fast(float h) { return __fdividef(1.f, 1.f + __expf(-h)); }
__device__ __forceinline__ float silu_fast(float h) { return __fdividef(h, 1.f + __expf(-h)); }

template <int N>
__device__ __forceinline__ void load_row(float (&v)[N], const float* __restrict__ src, bool ok) {
    if (N % 4 == 0) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const float4 t = ok ? *reinterpret_cast<const float4*>(src + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < N / 2; ++q) {
            const float2 t = ok ? *reinterpret_cast<const float2*>(src + 2 * q) : make_float2(0.f, 0.f);
            v[2 * q] = t.x; v[2 * q + 1] = t.y;
        }
    }
}
template <int N>
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float (&v)[N], bool ok) {
    if (!ok) return;
    if (N % 4 == 0) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < N / 2; ++q) *reinterpret_cast<float2*>(dst + 2 * q) = make_float2(v[2 * q], v[2 * q + 1]);
    }
}
// a thread's 4 x NJ tile -> the plane [col0 + j][r0 .. r0 + 3]
template <int NJ>
__device__ __forceinline__ void store_plane(float* __restrict__ plane, int col0, int r0, const float (&v)[4][NJ]) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) *reinterpret_cast<float4*>(plane + (col0 + j) * kRows + r0) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
}
// per-row sums over the 16 column groups: every thread deposits its four rows' partial sums, and after the caller's barrier reads the totals
__device__ __forceinline__ void red_put(float* red, int cg, int r0, const float (&v)[4]) {
    *reinterpret_cast<float4*>(red + cg * kRows + r0) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void red_get(const float* red, int r0, float (&v)[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) {
        const float4 t = *reinterpret_cast<const float4*>(red + w * kRows + r0);
        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
    }
}

struct FwdArgs {
    const float* p; const float* wt; const float* b0f; const float* hgpart; int splits; int64_t split_stride;
    float* hg[kStages]; float* a[kStages]; float* xh[kStages]; float* rstd[kStages]; float* logit;
    int B; int save;
    long long* stamps;      // profiling aid: clock64 of CTA 0 / thread 0 at the phase boundaries (nullptr: off)
    Stages L;
    // multi-model form (blockIdx.y = model): parameters of every model, transposed weights at wt + model * kWtFloats, the stacked
    // first-layer output hgpart [B][ld] with model m's [hidden 64 | gate 64] at column m * 128 (bias included), logits [model][B]
    const float* plist[kMlpMaxPack];
    int64_t ld;
};

// first-layer pre-activations (K slices) -> logits
template <bool kMulti>
__global__ void __launch_bounds__(kTailThreads, 2) fwd_tail_kernel(const FwdArgs f) {
    extern __shared__ __align__(128) unsigned char smraw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw);
    float* sm = reinterpret_cast<float*>(smraw + 128);
    float* abuf = sm + kMatFloats + kSmallMatFloats;   // gated activations [64][32]
    float* ubuf = abuf + kHg * kRows;              // stage input [96][32]
    float* red = ubuf + kDim * kRows;              // [2][16][32]
    const int tid = threadIdx.x, rq = tid & 7, cg = tid >> 3, r0 = 4 * rq;
    const int row0 = blockIdx.x * kRows + r0;
    bool ok[4], save[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ok[i] = row0 + i < f.B; save[i] = ok[i] && f.save && !kMulti; }
    const int model = kMulti ? blockIdx.y : 0;
    const float* __restrict__ P = kMulti ? f.plist[model] : f.p;
    const float* __restrict__ WT = f.wt + (int64_t)model * kWtFloats;

    MatList ml;
    for (int s = 0; s < 3; ++s) {
        ml.src[2 * s] = WT + wot_off(s); ml.floats[2 * s] = kHid * kDim;
        ml.src[2 * s + 1] = WT + whgt_off(s + 1); ml.floats[2 * s + 1] = kMatFloats;
    }
    WeightPipe wp{bars, {sm + kMatFloats, sm}};          // even matrices (W_o) are the small ones
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
        wp.issue(ml, 0);
        wp.issue(ml, 1);
    }
    __syncthreads();
    int n_stamp = 0;
    auto stamp = [&]() { if (f.stamps != nullptr && blockIdx.x == 0 && tid == 0) f.stamps[n_stamp++] = clock64(); };
    stamp();

#pragma unroll
    for (int s = 0; s < kStages; ++s) {        // unrolled: every per-stage pointer and offset is a fixed kernel-parameter slot
        const StageOff& S = f.L.s[s];
        float hgv[4][kNjHg];
        if (s == 0 && kMulti) {
            // the stacked product's row: this model's hidden columns, then its gate columns; bias included
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float h[kNjA], g[kNjA];
                const float* src = f.hgpart + (int64_t)(row0 + i) * f.ld + model * kHg + cg * kNjA;
                load_row(h, src, ok[i]);
                load_row(g, src + kHid, ok[i]);
#pragma unroll
                for (int j = 0; j < kNjA; ++j) { hgv[i][2 * j] = h[j]; hgv[i][2 * j + 1] = g[j]; }
            }
        } else if (s == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) load_row(hgv[i], f.b0f + cg * kNjHg, true);
            // all four rows of a K slice in flight together; the slices add up in a fixed order
#pragma unroll 4
            for (int z = 0; z < f.splits; ++z) {
                float t[4][kNjHg];
#pragma unroll
                for (int i = 0; i < 4; ++i) load_row(t[i], f.hgpart + z * f.split_stride + (int64_t)(row0 + i) * kHg + cg * kNjHg, ok[i]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int e = 0; e < kNjHg; ++e) hgv[i][e] += t[i][e];
            }
        } else {
            const float* W = wp.acquire(ml, 2 * s - 1, 6);        // ubuf is complete
            stamp();
#pragma unroll
            for (int i = 0; i < 4; ++i) load_row(hgv[i], W + kDim * kHg + cg * kNjHg, true);
            plane_fma4<kDim, kNjHg>(ubuf, W, kHg, cg * kNjHg, r0, hgv);
        }
        stamp();
        float av[4][kNjA];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < kNjA; ++j) {
                const float h = hgv[i][2 * j], g = hgv[i][2 * j + 1];
                av[i][j] = silu_fast(h) * g;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (save[i]) store_row(f.hg[s] + (int64_t)(row0 + i) * kHg + cg * kNjHg, hgv[i], true);
            if (save[i]) store_row(f.a[s] + (int64_t)(row0 + i) * kHid + cg * kNjA, av[i], true);
        }
        if (s < kStages - 1) {
            store_plane(abuf, cg * kNjA, r0, av);
            stamp();
            const float* W = wp.acquire(ml, 2 * s, 6);            // abuf is complete
            stamp();
            const StageOff& N = f.L.s[s + 1];
            float ov[4][kNjO], gam[kNjO], bet[kNjO];
            load_row(gam, P + N.ln_w + cg * kNjO, true);
            load_row(bet, P + N.ln_b + cg * kNjO, true);
#pragma unroll
            for (int i = 0; i < 4; ++i) load_row(ov[i], P + S.ob + cg * kNjO, true);
            plane_fma4<kHid, kNjO>(abuf, W, kDim, cg * kNjO, r0, ov);
            stamp();
            // LayerNorm of the next stage over each row's 96 features (6 per column group)
            float ps[4], mu[4], rs[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ps[i] = 0.f;
#pragma unroll
                for (int j = 0; j < kNjO; ++j) ps[i] += ov[i][j];
            }
            red_put(red, cg, r0, ps);
            __syncthreads();
            red_get(red, r0, mu);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                mu[i] *= (1.f / kDim);
                ps[i] = 0.f;
#pragma unroll
                for (int j = 0; j < kNjO; ++j) { const float d = ov[i][j] - mu[i]; ps[i] = fmaf(d, d, ps[i]); }
            }
            red_put(red + kRed, cg, r0, ps);
            __syncthreads();
            red_get(red + kRed, r0, rs);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                rs[i] = rsqrtf(rs[i] * (1.f / kDim) + kLnEps);
                float xh[kNjO];
#pragma unroll
                for (int j = 0; j < kNjO; ++j) {
                    xh[j] = (ov[i][j] - mu[i]) * rs[i];
                    ov[i][j] = fmaf(xh[j], gam[j], bet[j]);
                }
                if (save[i]) {
                    store_row(f.xh[s + 1] + (int64_t)(row0 + i) * kDim + cg * kNjO, xh, true);
                    if (cg == 0) f.rstd[s + 1][row0 + i] = rs[i];
                }
            }
            store_plane(ubuf, cg * kNjO, r0, ov);
            stamp();
        } else {
            float ps[4], z[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ps[i] = 0.f;
#pragma unroll
                for (int j = 0; j < kNjA; ++j) ps[i] = fmaf(av[i][j], __ldg(P + S.ow + cg * kNjA + j), ps[i]);
            }
            __syncthreads();                       // the previous stage's reads of `red` are over
            red_put(red, cg, r0, ps);
            __syncthreads();
            red_get(red, r0, z);
            if (cg == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (ok[i]) f.logit[(int64_t)model * f.B + row0 + i] = z[i] + __ldg(P + S.ob);
            }
        }
    }
    stamp();
}

struct BwdArgs {
    const float* p; const float* wt; const float* dz;
    const float* hg[kStages]; const float* xh[kStages]; const float* rstd[kStages];
    float* dhg[kStages]; float* dout[kStages - 1];
    int B;
    Stages L;
};

// d loss / d logit -> d_hg of every stage and d_o of stages 0..2 (what the weight-gradient products read)
__global__ void __launch_bounds__(kTailThreads, 2) bwd_tail_kernel(const BwdArgs f) {
    extern __shared__ __align__(128) unsigned char smraw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw);
    float* sm = reinterpret_cast<float*>(smraw + 128);
    float* dbuf = sm + kMatFloats + kSmallMatFloats;   // d_hg plane [128][32]
    float* dobuf = dbuf + kHg * kRows;             // d_o plane [96][32]
    float* red = dobuf + kDim * kRows;
    const int tid = threadIdx.x, rq = tid & 7, cg = tid >> 3, r0 = 4 * rq;
    const int row0 = blockIdx.x * kRows + r0;
    bool ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ok[i] = row0 + i < f.B;

    // in the order of use: W_I of stage 3, then (W_o, W_I) of stages 2 and 1, then W_o of stage 0
    MatList ml;
    ml.src[0] = f.wt + wi_off(3); ml.floats[0] = kHg * kDim;
    ml.src[1] = f.p + f.L.s[2].ow; ml.floats[1] = kDim * kHid;
    ml.src[2] = f.wt + wi_off(2); ml.floats[2] = kHg * kDim;
    ml.src[3] = f.p + f.L.s[1].ow; ml.floats[3] = kDim * kHid;
    ml.src[4] = f.wt + wi_off(1); ml.floats[4] = kHg * kDim;
    ml.src[5] = f.p + f.L.s[0].ow; ml.floats[5] = kDim * kHid;
    WeightPipe wp{bars, {sm, sm + kMatFloats}};          // even matrices (W_I) are the large ones
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
        wp.issue(ml, 0);
        wp.issue(ml, 1);
    }
    __syncthreads();

    int mat = 0;
#pragma unroll
    for (int s = kStages - 1; s >= 0; --s) {
        const StageOff& S = f.L.s[s];
        // this stage's saved pre-activations: requested before the product that precedes their use
        float hgv[4][kNjHg];
#pragma unroll
        for (int i = 0; i < 4; ++i) load_row(hgv[i], f.hg[s] + (int64_t)(row0 + i) * kHg + cg * kNjHg, ok[i]);
        float dav[4][kNjA];
        if (s == kStages - 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float dzr = ok[i] ? f.dz[row0 + i] : 0.f;
#pragma unroll
                for (int j = 0; j < kNjA; ++j) dav[i][j] = dzr * __ldg(f.p + S.ow + cg * kNjA + j);
            }
        } else {
            const float* W = wp.acquire(ml, mat++, 6);            // dobuf is complete
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < kNjA; ++j) dav[i][j] = 0.f;
            plane_fma4<kDim, kNjA>(dobuf, W, kHid, cg * kNjA, r0, dav);
        }
        float dhv[4][kNjHg];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < kNjA; ++j) {
                const float h = hgv[i][2 * j], g = hgv[i][2 * j + 1];
                const float sg = sigmoid_fast(h);
                dhv[i][2 * j] = dav[i][j] * g * (sg * (1.f + h * (1.f - sg)));      // d hidden
                dhv[i][2 * j + 1] = dav[i][j] * h * sg;                             // d gate
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) store_row(f.dhg[s] + (int64_t)(row0 + i) * kHg + cg * kNjHg, dhv[i], ok[i]);
        if (s == 0) break;
        store_plane(dbuf, cg * kNjHg, r0, dhv);
        float gam[kNjO], xh[4][kNjO], rsv[4], p1[4], p2[4], m1[4], m2[4];
        load_row(gam, f.p + S.ln_w + cg * kNjO, true);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            load_row(xh[i], f.xh[s] + (int64_t)(row0 + i) * kDim + cg * kNjO, ok[i]);
            rsv[i] = ok[i] ? f.rstd[s][row0 + i] : 0.f;
        }
        const float* W = wp.acquire(ml, mat++, 6);                // dbuf is complete
        float duv[4][kNjO];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < kNjO; ++j) duv[i][j] = 0.f;
        plane_fma4<kHg, kNjO>(dbuf, W, kDim, cg * kNjO, r0, duv);
        // LayerNorm backward (input of stage s = output of stage s - 1)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            p1[i] = 0.f; p2[i] = 0.f;
#pragma unroll
            for (int j = 0; j < kNjO; ++j) {
                duv[i][j] *= gam[j];
                p1[i] += duv[i][j];
                p2[i] = fmaf(duv[i][j], xh[i][j], p2[i]);
            }
        }
        red_put(red, cg, r0, p1);
        red_put(red + kRed, cg, r0, p2);
        __syncthreads();
        red_get(red, r0, m1);
        red_get(red + kRed, r0, m2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float rs = rsv[i];
            const float a1 = m1[i] * (1.f / kDim), a2 = m2[i] * (1.f / kDim);
#pragma unroll
            for (int j = 0; j < kNjO; ++j) duv[i][j] = rs * (duv[i][j] - a1 - xh[i][j] * a2);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) store_row(f.dout[s - 1] + (int64_t)(row0 + i) * kDim + cg * kNjO, duv[i], ok[i]);
        store_plane(dobuf, cg * kNjO, r0, duv);
    }
}

// ---- finish: add the K slices up (in slice order) and turn G and db into the parameter gradients --------------------------------------------
struct FinArgs {
    const float* p; float* g;
    const float* part;          // the grouped launch's small products: [slices][8 slots][128][128]
    const float* g0;            // the first-layer G: [slices][128][1536]
    int slices;
    Stages L;
};
constexpr int kFinColBlocks = kIn / 32 + 3 * (kDim / 32);    // 32 input columns of one stage per block
constexpr int kFinOutBlocks = 8;
constexpr int kFinThreads = 1024, kFinGroups = kFinThreads / 32;

__global__ void __launch_bounds__(kFinThreads) finish_kernel(const FinArgs f) {
    __shared__ float db[kHg];
    __shared__ float red[2][kFinGroups][32];
    const int tid = threadIdx.x;
    const float* __restrict__ part = f.part;
    const float* __restrict__ p = f.p;
    float* __restrict__ g = f.g;
    const int nz = f.slices;
    // sum over the K slices of one word, slice order, the loads of all slices in flight together
    auto zsum = [&](const float* __restrict__ base, int64_t stride) {
        float v[kGradSlicesMax];
#pragma unroll
        for (int z = 0; z < kGradSlicesMax; ++z) v[z] = z < nz ? base[z * stride] : 0.f;
        float t = v[0];
#pragma unroll
        for (int z = 1; z < kGradSlicesMax; ++z) t += v[z];
        return t;
    };
    if (blockIdx.x >= kFinColBlocks) {
        // output linears: dW_o = O, db_o = column sums of d_o
        const int t0 = (blockIdx.x - kFinColBlocks) * kFinThreads + tid, stride = kFinOutBlocks * kFinThreads;
        for (int s = 0; s < kStages; ++s) {
            const StageOff& S = f.L.s[s];
            const float* __restrict__ O = part + (4 + s) * kSlot;
            for (int i = t0; i < S.out_dim * kHid; i += stride) g[S.ow + i] = zsum(O + (i / kHid) * 128 + (i % kHid), kSmallFloatsPerSlice);
            for (int n = t0; n < S.out_dim; n += stride) g[S.ob + n] = zsum(O + n * 128 + kHid, kSmallFloatsPerSlice);
        }
        return;
    }
    int s, kb;
    if ((int)blockIdx.x < kIn / 32) { s = 0; kb = blockIdx.x; }
    else { s = 1 + (blockIdx.x - kIn / 32) / (kDim / 32); kb = (blockIdx.x - kIn / 32) % (kDim / 32); }
    const StageOff& S = f.L.s[s];
    const int in = S.in_dim;
    if (tid < kHg) {
        const float v = zsum(s == 0 ? part + tid * 128 : part + s * kSlot + tid * 128 + kDim, kSmallFloatsPerSlice);
        db[tid] = v;
        if (kb == 0) g[((tid & 1) ? S.gb : S.hb) + (tid >> 1)] = v;
    }
    __syncthreads();
    const int kx = tid & 31, ny = tid >> 5, k = kb * 32 + kx;
    const float gamma = p[S.ln_w + k], beta = p[S.ln_b + k];
    const float* __restrict__ G = s == 0 ? f.g0 : part + s * kSlot;
    const int ldg = s == 0 ? kIn : 128;
    const int64_t zstride = s == 0 ? (int64_t)kHg * kIn : kSmallFloatsPerSlice;
    float Gv[kHg / kFinGroups], Wv[kHg / kFinGroups];
#pragma unroll
    for (int i = 0; i < kHg / kFinGroups; ++i) {
        const int n = ny + kFinGroups * i;
        Gv[i] = zsum(G + (int64_t)n * ldg + k, zstride);
        Wv[i] = p[((n & 1) ? S.gw : S.hw) + (n >> 1) * in + k];
    }
    float dgam = 0.f, dbet = 0.f;
#pragma unroll
    for (int i = 0; i < kHg / kFinGroups; ++i) {
        const int n = ny + kFinGroups * i;
        g[((n & 1) ? S.gw : S.hw) + (n >> 1) * in + k] = fmaf(gamma, Gv[i], beta * db[n]);
        dgam = fmaf(Wv[i], Gv[i], dgam);
        dbet = fmaf(Wv[i], db[n], dbet);
    }
    red[0][ny][kx] = dgam;
    red[1][ny][kx] = dbet;
    __syncthreads();
    if (ny == 0) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < kFinGroups; ++w) { a += red[0][w][kx]; b += red[1][w][kx]; }
        g[S.ln_w + k] = a;
        g[S.ln_b + k] = b;
    }
}

int configure_tail_smem() {
    static bool done = false;
    if (!done) {
        HB_CUDA_OK(cudaFuncSetAttribute(fwd_tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemBytes));
        HB_CUDA_OK(cudaFuncSetAttribute(fwd_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemBytes));
        HB_CUDA_OK(cudaFuncSetAttribute(bwd_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmemBytes));
        done = true;
    }
    return HB_OK;
}

}  // namespace

static long long* g_fwd_stamps = nullptr;
}  // namespace hb
// profiling aid (not in the public header): clock64 stamps of CTA 0 of the next forward-tail launches go to `stamps_dev` (32 values; NULL: off)
extern "C" int hb_debug_mlp_stamps(long long* stamps_dev) { hb::g_fwd_stamps = stamps_dev; return 0; }
namespace hb {

int mlp_fused_check_timeout() {
    unsigned int flag = 0;
    HB_CUDA_OK(cudaMemcpyFromSymbol(&flag, g_tc_timeout, sizeof(flag)));
    HB_REQUIRE(flag == 0, "classifier tail kernels: an mbarrier wait timed out (pipeline bug; barrier code %u)", flag);
    return HB_OK;
}

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
    pre_kernel<<<kPrepBlocks + ceil_div(B, 8), 256, 0, st>>>(m->p, L, w.w0f, w.b0f, w.wt, x, w.mean0, w.rstd0, B);
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
    f.stamps = g_fwd_stamps;
    f.L = L;
    f.ld = 0;
    fwd_tail_kernel<false><<<ceil_div(B, kRows), kTailThreads, kTailSmemBytes, st>>>(f);
    HB_LAUNCHED();
    *logits = w.logit;
    return HB_OK;
}

int64_t mlp_fused_multi_wt_floats(int n_models) { return (int64_t)n_models * kWtFloats; }

// The 96-wide remainder of n <= 64 models behind ONE stacked first-layer product: hg [B][ld], model i's [hidden 64 | gate 64]
// pre-activations (bias included) at column i * 128  ->  logits [n][B].  One forward-tail CTA per (32 rows, model).
int mlp_fused_multi_tail(const float* const* params, int n, const float* hg, int64_t ld, int B, float* wt_all, float* logits, cudaStream_t st) {
    HB_REQUIRE(n >= 1 && n <= kMlpMaxPack && ld % 4 == 0, "classifier (multi): bad model count or row stride");
    int rc;
    if ((rc = configure_tail_smem())) return rc;
    const Stages L = stages();
    ModelList ml;
    FwdArgs f;
    for (int i = 0; i < n; ++i) { ml.p[i] = params[i]; f.plist[i] = params[i]; }
    prep_multi_kernel<<<dim3(kPrepBlocks - kPrepFoldBlocks, n), 256, 0, st>>>(ml, L, wt_all);
    HB_LAUNCHED();
    f.p = nullptr; f.wt = wt_all; f.b0f = nullptr; f.hgpart = hg; f.splits = 1; f.split_stride = 0;
    for (int s = 0; s < kStages; ++s) { f.hg[s] = nullptr; f.a[s] = nullptr; f.xh[s] = nullptr; f.rstd[s] = nullptr; }
    f.logit = logits;
    f.B = B; f.save = 0;
    f.stamps = nullptr;
    f.L = L;
    f.ld = ld;
    fwd_tail_kernel<true><<<dim3(ceil_div(B, kRows), n), kTailThreads, kTailSmemBytes, st>>>(f);
    HB_LAUNCHED();
    return HB_OK;
}

int mlp_fused_backward(hb_mlp_model* m, const float* x, int B, float* ws, const float* dz, cudaStream_t st, float* g_out) {
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
    bwd_tail_kernel<<<ceil_div(B, kRows), kTailThreads, kTailSmemBytes, st>>>(b);
    HB_LAUNCHED();

    // every weight-gradient product of the step in one grouped tcgen05 launch: G_0 = d_hg0^T xhat_0 (xhat formed from x and the row
    // statistics in the operand load), the 96-wide G_s, the output-linear gradients, and -- as a column of ones in the B operand -- the
    // column sums that are the bias gradients
    const int slices = grad_slices(B);
    TfGroup grp;
    int np = 0;
    auto add = [&](const float* A, int lda, int M, const float* Bm, int ldb, int N, int ones_col, float* C, int ldc, int store_n, int64_t stride) {
        TfArgs& t = grp.pr[np++];
        t.A = A; t.lda = lda;
        t.B0 = Bm; t.B1 = nullptr; t.bsplit = 1 << 30; t.ldb = ldb;
        t.bias0 = nullptr; t.bias1 = nullptr; t.biassplit = 1 << 30;
        t.C = C; t.ldc = ldc;
        t.M = M; t.N = N; t.K = B;
        t.mean = nullptr; t.rstd = nullptr;
        t.splits = slices; t.split_stride = stride;
        t.ones_col = ones_col; t.store_n = store_n;
    };
    add(w.dhg[0], kHg, kHg, x, kIn, kIn, -1, w.g0part, kIn, kIn, (int64_t)kHg * kIn);
    grp.pr[0].mean = w.mean0; grp.pr[0].rstd = w.rstd0;
    add(w.dhg[0], kHg, kHg, nullptr, 0, 0, 0, w.part, 128, 4, kSmallFloatsPerSlice);
    for (int s = 1; s < kStages; ++s) add(w.dhg[s], kHg, kHg, w.xh[s], kDim, kDim, kDim, w.part + s * kSlot, 128, kDim + 4, kSmallFloatsPerSlice);
    for (int s = 0; s < kStages - 1; ++s) add(w.dout[s], kDim, kDim, w.a[s], kHid, kHid, kHid, w.part + (4 + s) * kSlot, 128, kHid + 4, kSmallFloatsPerSlice);
    add(dz, 1, 1, w.a[3], kHid, kHid, kHid, w.part + 7 * kSlot, 128, kHid + 4, kSmallFloatsPerSlice);
    grp.nprob = np;
    if ((rc = gemm_tf32x3_launch_group(grp, st))) return rc;

    FinArgs fin;
    fin.p = m->p; fin.g = g_out != nullptr ? g_out : m->g;
    fin.part = w.part;
    fin.g0 = w.g0part;
    fin.slices = slices;
    fin.L = L;
    finish_kernel<<<kFinColBlocks + kFinOutBlocks, kFinThreads, 0, st>>>(fin);
    HB_LAUNCHED();
    return HB_OK;
}

}  // namespace hb
