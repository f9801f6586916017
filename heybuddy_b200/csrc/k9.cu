// K9: the reference's per-clip numpy augmentations on the device -- SevenBandParametricEQ and TanhDistortion
// (audiomentations.Compose at reference src/python/heybuddy/dataset/augmented.py:79-90, applied to every length-fixed clip at
// :325-328).  audiomentations is absent offline: both are restated from the library's published behaviour (parity unpinned; the
// arithmetic spec and every constant live in heybuddy_b200/dataset/k9.py, the CPU restatement in oracle/k9.py).
//
// Both kernels work in place on the f32 [n][T] length-fixed clips and touch only the clips whose coin came up (index lists from
// the draw table), so a chunk with the reference's default probabilities (0.25 each) pays for a quarter of its clips.
#include "hb_common.cuh"

#include <math.h>

namespace hb {

constexpr int kEqBands = 7;

// ---- SevenBandParametricEQ: seven cascaded biquads, causal, zero initial state (scipy.signal.sosfilt) ----------------------------
// scipy filters in float64; a float32 DIRECT form loses up to 1e-2 on the low-Q shelves (poles a few 1e-3 from z = 1: the
// coefficients cancel).  Round 2 first ran the cascade in FP64 (1.9 ms per 2029 clips: B200 issues non-tensor FP64 at one lane per
// clock per SM sub-partition).  The same transfer function in the trapezoidal state-variable form (Zavalishin's TPT structure) has no
// such cancellation: any stable biquad maps exactly onto it --
//     z^-1 = (1 - s) / (1 + s):  H = (n2 s^2 + n1 s + n0) / (d2 s^2 + d1 s + d0),  n2 = b0 - b1 + b2, n1 = 2 (b0 - b2), n0 = b0 + b1 + b2,
//                                                                                 d2 = 1 - a1 + a2,  d1 = 2 (1 - a2),  d0 = 1 + a1 + a2
//     g = sqrt(d0 / d2), k = d1 / (d2 g);  out = m0 v0 + m1 v1 + m2 v2,  m0 = n2 / d2, m1 = n1 / (d2 g) - m0 k, m2 = n0 / (d2 g^2) - m0
//     v3 = v0 - ic2;  v1 = a1 ic1 + a2 v3;  v2 = ic2 + a2 ic1 + a3 v3;  ic1 = 2 v1 - ic1;  ic2 = 2 v2 - ic2   (a1 = 1 / (1 + g (g + k)), a2 = g a1, a3 = g a2)
// -- and in FLOAT32 it stays within 4e-6 of float64 sosfilt over the whole parameter range (measured on the CPU, 64 random cascades;
// 8e-12 in float64).  The per-section constants are derived from the caller's float64 sos once per lane, in float64.
//
// The cascade is a SYSTOLIC ARRAY across lanes: eight lanes per clip (four clips per warp), lane s owns section s (lane 7 is a pure
// delay); at tick tau lane s works on sample tau - 3 s, its input arriving by shuffle from lane s - 1's output of three ticks ago,
// so the shuffle is off the loop-carried chain (ic1, ic2: four dependent float32 operations per tick).  Lane 7's output lags the
// input by 24 samples, so loads and stores are whole, aligned float4s (in place: the store trails the load by six groups); lane 0's
// loads run one 32-sample block ahead of the arithmetic.
constexpr int kEqLanes = 8, kEqDelay = 3, kEqLagGroups = kEqLanes * kEqDelay / 4;      // 24 samples = 6 float4 groups
constexpr int kEqBlock = 8;                                                             // float4 groups per prefetched block

__global__ void __launch_bounds__(32) k9_eq_kernel(float* __restrict__ clips, const int32_t* __restrict__ clip_index,
                                                   const double* __restrict__ sos, int k, int T) {
    static_assert(kEqDelay == 3 && (kEqLanes * kEqDelay) % 4 == 0, "the delay registers below are written out for a delay of 3");
    const int lane = threadIdx.x & 31, sec = lane & (kEqLanes - 1);
    const int i = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (32 / kEqLanes) + lane / kEqLanes;   // clip of this 8-lane group
    const bool live = i < k;
    float a1 = 0.f, a2 = 0.f, a3 = 0.f, m0 = 1.f, m1 = 0.f, m2 = 0.f;       // lane 7 (and idle groups): identity
    if (live && sec < kEqBands) {
        const double* c = sos + ((int64_t)i * kEqBands + sec) * 5;
        const double b0 = c[0], b1 = c[1], b2 = c[2], p1 = c[3], p2 = c[4];
        const double n2 = b0 - b1 + b2, n1 = 2.0 * (b0 - b2), n0 = b0 + b1 + b2;
        const double d2 = 1.0 - p1 + p2, d1 = 2.0 * (1.0 - p2), d0 = 1.0 + p1 + p2;
        const double g = sqrt(d0 / d2), kk = d1 / (d2 * g);
        const double hp = n2 / d2;
        const double e1 = 1.0 / (1.0 + g * (g + kk));
        a1 = (float)e1; a2 = (float)(g * e1); a3 = (float)(g * g * e1);
        m0 = (float)hp; m1 = (float)(n1 / (d2 * g) - hp * kk); m2 = (float)(n0 / (d2 * g * g) - hp);
    }
    float* x = clips + (int64_t)(live ? clip_index[i] : 0) * T;
    const bool vec = (T & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    const bool loader = live && sec == 0, storer = live && sec == kEqLanes - 1;
    float ic1 = 0.f, ic2 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;        // state; this lane's last three outputs
    float up = 0.f;                                                   // lane s - 1's output of three ticks ago (fetched one tick ahead)
    const int groups = (T + 3) / 4, total = groups + kEqLagGroups;
    auto load_group = [&](int g) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (loader && g < groups) {
            if (vec) v = *reinterpret_cast<const float4*>(x + 4 * g);
            else {
                if (4 * g + 0 < T) v.x = x[4 * g + 0];
                if (4 * g + 1 < T) v.y = x[4 * g + 1];
                if (4 * g + 2 < T) v.z = x[4 * g + 2];
                if (4 * g + 3 < T) v.w = x[4 * g + 3];
            }
        }
        return v;
    };
    float4 nxt[kEqBlock];
#pragma unroll
    for (int q = 0; q < kEqBlock; ++q) nxt[q] = load_group(q);
    for (int g0 = 0; g0 < total; g0 += kEqBlock) {
        float4 cur[kEqBlock];
#pragma unroll
        for (int q = 0; q < kEqBlock; ++q) cur[q] = nxt[q];
#pragma unroll
        for (int q = 0; q < kEqBlock; ++q) nxt[q] = load_group(g0 + kEqBlock + q);     // in flight during this block's 32 ticks
#pragma unroll
        for (int q = 0; q < kEqBlock; ++q) {
            const int g = g0 + q;
            if (g >= total) break;
            const float xin[4] = {cur[q].x, cur[q].y, cur[q].z, cur[q].w};
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float v0 = sec == 0 ? xin[j] : up;
                up = __shfl_up_sync(0xffffffffu, d2, 1, kEqLanes);     // the neighbour's d3 of the NEXT tick: the shuffle's latency is off the chain
                o[j] = d3;                                // lane 7: sample tau - 24 = 4 (g - 6) + j
                const float v3 = v0 - ic2;
                const float v1 = fmaf(a2, v3, a1 * ic1);
                const float v2 = fmaf(a3, v3, fmaf(a2, ic1, ic2));
                ic1 = fmaf(2.0f, v1, -ic1);
                ic2 = fmaf(2.0f, v2, -ic2);
                const float y = fmaf(m0, v0, fmaf(m1, v1, m2 * v2));
                d3 = d2; d2 = d1; d1 = y;
            }
            if (storer && g >= kEqLagGroups) {
                float* dst = x + 4 * (g - kEqLagGroups);
                if (vec) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                else
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (4 * (g - kEqLagGroups) + j < T) dst[j] = o[j];
            }
        }
    }
}

// ---- TanhDistortion ------------------------------------------------------------------------------------------------------------------
// threshold = np.percentile(|x|, 100 - 99 * amount) (linear interpolation between the two neighbouring order statistics); the order
// statistics come from an exact radix select over the bit patterns of |x| (non-negative floats order like their bits): four 8-bit
// passes per rank over the clip's magnitudes in shared memory.
constexpr int kTanhThreads = 512;

__device__ __forceinline__ uint32_t radix_select(const uint32_t* __restrict__ bits, int n, int rank, uint32_t* hist, uint32_t* bcast) {
    uint32_t prefix = 0, mask = 0;
    int remaining = rank;                 // rank among the elements that match the prefix
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += kTanhThreads) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kTanhThreads) {
            const uint32_t v = bits[i];
            if ((v & mask) == prefix) atomicAdd(&hist[(v >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, b = 0;
            for (; b < 256; ++b) {
                const int c = (int)hist[b];
                if (remaining < acc + c) break;
                acc += c;
            }
            bcast[0] = (uint32_t)b;
            bcast[1] = (uint32_t)(remaining - acc);
        }
        __syncthreads();
        prefix |= bcast[0] << shift;
        mask |= 255u << shift;
        remaining = (int)bcast[1];
        __syncthreads();
    }
    return prefix;
}

__device__ __forceinline__ float block_sum_512(float v, float* scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < kTanhThreads / 32 ? scratch[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

__global__ void __launch_bounds__(kTanhThreads) k9_tanh_kernel(float* __restrict__ clips, const int32_t* __restrict__ clip_index,
                                                               const float* __restrict__ amount, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t bcast[2];
    __shared__ float scratch[40];
    uint32_t* bits = reinterpret_cast<uint32_t*>(smem_raw);
    float* x = clips + (int64_t)clip_index[blockIdx.x] * T;
    for (int i = threadIdx.x; i < T; i += kTanhThreads) bits[i] = __float_as_uint(fabsf(x[i]));
    __syncthreads();
    // np.percentile(..., q) with q = 100 - 99 * amount: virtual index q / 100 * (n - 1), linear interpolation
    const double q = 100.0 - 99.0 * (double)amount[blockIdx.x];
    const double pos = q / 100.0 * (double)(T - 1);
    int lo = (int)floor(pos);
    lo = max(0, min(lo, T - 1));
    const int hi = min(lo + 1, T - 1);
    const float frac = (float)(pos - (double)lo);
    const float v_lo = __uint_as_float(radix_select(bits, T, lo, hist, bcast));
    const float v_hi = hi == lo ? v_lo : __uint_as_float(radix_select(bits, T, hi, hist, bcast));
    const float threshold = v_lo + frac * (v_hi - v_lo);
    const float gain = 0.5f / (threshold + 1e-6f);
    float* y = reinterpret_cast<float*>(smem_raw);      // the magnitudes are not needed any more
    float sx = 0.f, sy = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < T; i += kTanhThreads) {
        const float v = x[i];
        const float d = tanhf(gain * v);
        y[i] = d;
        sx = fmaf(v, v, sx);
        sy = fmaf(d, d, sy);
    }
    const float rms_before = sqrtf(block_sum_512(sx, scratch) / (float)T);
    const float rms_after = sqrtf(block_sum_512(sy, scratch) / (float)T);
    const float post = rms_before > 1e-9f ? rms_before / rms_after : 1.0f;
    for (int i = threadIdx.x; i < T; i += kTanhThreads) x[i] = y[i] * post;
}

}  // namespace hb

using namespace hb;

extern "C" int hb_k9_eq_f32(float* clips_dev, const int32_t* clip_index_dev, const double* sos_dev, int k, int T, void* stream) {
    HB_REQUIRE(k >= 0 && T > 0 && (k == 0 || (clips_dev && clip_index_dev && sos_dev)), "hb_k9_eq_f32: bad argument");
    if (k == 0) return HB_OK;
    // one warp (four clips) per block: a warp is one latency chain, so spread the warps over every SM sub-partition
    k9_eq_kernel<<<ceil_div(k, 32 / kEqLanes), 32, 0, (cudaStream_t)stream>>>(clips_dev, clip_index_dev, sos_dev, k, T);
    HB_LAUNCHED();
    return HB_OK;
}

extern "C" int hb_k9_tanh_f32(float* clips_dev, const int32_t* clip_index_dev, const float* amount_dev, int k, int T, void* stream) {
    HB_REQUIRE(k >= 0 && T > 0 && T <= 24576 && (k == 0 || (clips_dev && clip_index_dev && amount_dev)), "hb_k9_tanh_f32: bad argument (T <= 24576)");
    if (k == 0) return HB_OK;
    static bool configured = false;
    if (!configured) {
        HB_CUDA_OK(cudaFuncSetAttribute(k9_tanh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576 * 4));
        configured = true;
    }
    k9_tanh_kernel<<<k, kTanhThreads, (size_t)T * 4, (cudaStream_t)stream>>>(clips_dev, clip_index_dev, amount_dev, T);
    HB_LAUNCHED();
    return HB_OK;
}
