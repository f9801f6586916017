// K1-K4: fused augmentation -- coloured/white noise at a batch SNR, gain, background noise at a
// per-clip SNR, and room-impulse-response reverb -- one kernel, one CTA per clip, the clip
// resident in shared memory from the first load to the final store.
//
// Replaces the library calls of execute_augment_batch (reference
// src/python/heybuddy/dataset/augmented.py:363-392): torch_audiomentations AddColoredNoise +
// Gain (mode="per_batch"), torchaudio.functional.add_noise, speechbrain reverberate.
// Arithmetic spec: SURVEY.md A.3 / oracle/augment.py.
//
// Reverb = circular convolution of exactly T samples (speechbrain convolve1d(use_fft=True) on
// the rotated RIR), computed with an exact-length real FFT: the T real samples are packed into
// M = T/2 complex points, transformed by a mixed-radix (4,2,3,5) Stockham FFT that ping-pongs
// between two M-point shared-memory buffers, untangled / multiplied by the precomputed RIR
// spectrum / re-tangled in one pointwise pass, and transformed back by the same forward FFT
// under conjugation.  No zero padding: T = 23040 = 2 * (4^4 * 3^2 * 5).
//
// HBM traffic per clip: read clip (4T) + read noise row (4T, when drawn) + read coloured base
// (64 KB, L2-resident per batch) + RIR spectrum (4T+8, L2-resident per batch) + write 4T.
#include "mel_core.cuh"

#include "cplx.cuh"

#include <math.h>

#include <map>
#include <mutex>
#include <vector>

namespace hb {

constexpr int kAugThreads = 1024;
constexpr int kMaxPasses = 12;
constexpr int kColoredBase = 16000;
constexpr int kMaxT = 23040;

struct FftPlan {
    int M;                 // complex points (T / 2)
    int n_passes;
    int radix[kMaxPasses];
    const float2* tw_m;    // exp(-2 pi i k / M), k in [0, M)
    const float2* tw_t;    // exp(-2 pi i k / T), k in [0, M]
    int tw_lo;             // two-level twiddle split: exp(-2 pi i m / M) = hi[m / tw_lo] * lo[m % tw_lo]
    int tw_hi;             // number of hi entries (tw_hi * tw_lo >= M)
};

constexpr int kTwMax = 128;   // entries per twiddle level kept in shared memory

struct TwTables {
    float2 hi[kTwMax];
    float2 lo[kTwMax];
};

// exp(-2 pi i m / M) from the two-level table: one shared-memory pair + one complex multiply instead of a
// scattered global load per butterfly input.
__device__ __forceinline__ float2 twiddle(const TwTables& t, int m, int lo_n) {
    const int h = m / lo_n;
    return cmulf(t.hi[h], t.lo[m - h * lo_n]);
}

// One Stockham autosort pass of radix R over M points (forward transform).
template <int R>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ in, float2* __restrict__ out, int M, int Ns,
                                         const TwTables& tw, int lo_n) {
    const int nb = M / R;
    const int tw_step = M / (Ns * R);
    for (int j = threadIdx.x; j < nb; j += kAugThreads) {
        const int k = j % Ns;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[j + r * nb];
        if (k != 0) {
            const float2 w1 = twiddle(tw, k * tw_step, lo_n);
            float2 w = w1;
            v[1] = cmulf(v[1], w);
#pragma unroll
            for (int r = 2; r < R; ++r) {
                w = cmulf(w, w1);   // w1^r
                v[r] = cmulf(v[r], w);
            }
        }
        dft<R>(v);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out[base + r * Ns] = v[r];
    }
    __syncthreads();
}

// Runs the whole plan from `a` (ping-ponging with `b`); returns the buffer holding the result.
__device__ __forceinline__ float2* fft_forward(float2* a, float2* b, const FftPlan& plan, const TwTables& tw) {
    int Ns = 1;
    for (int p = 0; p < plan.n_passes; ++p) {
        const int R = plan.radix[p];
        if (R == 4) fft_pass<4>(a, b, plan.M, Ns, tw, plan.tw_lo);
        else if (R == 2) fft_pass<2>(a, b, plan.M, Ns, tw, plan.tw_lo);
        else if (R == 3) fft_pass<3>(a, b, plan.M, Ns, tw, plan.tw_lo);
        else fft_pass<5>(a, b, plan.M, Ns, tw, plan.tw_lo);
        Ns *= R;
        float2* t = a; a = b; b = t;
    }
    return a;
}

__device__ __forceinline__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // protect scratch reuse
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < kAugThreads / 32) ? scratch[threadIdx.x] : 0.f;
    if (warp == 0) {
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// X[k], X[M-k] of the real signal from Z = FFT_M(x_even + i x_odd); w = exp(-2 pi i k / T).
__device__ __forceinline__ void untangle(float2 A, float2 B, float2 w, float2* Xk, float2* Xmk) {
    const float2 E = cscale(cadd(A, cconj(B)), 0.5f);
    const float2 O = mul_neg_i(cscale(csub(A, cconj(B)), 0.5f));
    const float2 wO = cmulf(w, O);
    *Xk = cadd(E, wO);
    // X[M-k] = conj(E) + W^{M-k} conj(O),  W^{M-k} = -conj(w)
    *Xmk = csub(cconj(E), cconj(wO));
}

// ------------------------------------------------------------------------------------------------
// RIR spectrum: kernels f32 [n][T] -> H f32 [n][M+1][2]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAugThreads, 1)
rir_spectrum_kernel(const float* __restrict__ kernels, float2* __restrict__ spec, int T, FftPlan plan) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = plan.M;
    float2* buf0 = reinterpret_cast<float2*>(smem_raw);
    float2* buf1 = buf0 + M;
    __shared__ TwTables tw;
    for (int i = threadIdx.x; i < plan.tw_hi; i += kAugThreads) tw.hi[i] = plan.tw_m[i * plan.tw_lo];
    for (int i = threadIdx.x; i < plan.tw_lo; i += kAugThreads) tw.lo[i] = plan.tw_m[i];
    const float* src = kernels + (int64_t)blockIdx.x * T;
    for (int i = threadIdx.x; i < M; i += kAugThreads) buf0[i] = make_float2(src[2 * i], src[2 * i + 1]);
    __syncthreads();
    const float2* Z = fft_forward(buf0, buf1, plan, tw);
    float2* H = spec + (int64_t)blockIdx.x * (M + 1);
    for (int k = threadIdx.x; k <= M / 2; k += kAugThreads) {
        if (k == 0) {
            H[0] = make_float2(Z[0].x + Z[0].y, 0.f);
            H[M] = make_float2(Z[0].x - Z[0].y, 0.f);
        } else {
            float2 Xk, Xmk;
            untangle(Z[k], Z[M - k], __ldg(plan.tw_t + k), &Xk, &Xmk);
            H[k] = Xk;
            H[M - k] = Xmk;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// coloured-noise patterns on the device: N(0,1)[16000] from the draw table's counters -> 1/f^decay shaping -> unit RMS
// ------------------------------------------------------------------------------------------------
// Philox4x32-10, the same function as heybuddy_b200/dataset/draws.py:philox4x32 (counter = (index, stream, batch lo, batch hi),
// key = seed lo / hi).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
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

// Box-Muller pair from two 32-bit words: u1 = ((a >> 9) + 0.5) / 2^23 in (0, 1), u2 = (b >> 8) / 2^24 in [0, 1), both exact in fp32
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)(a >> 9) + 0.5f) * (1.0f / 8388608.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif((float)(b >> 8) * (2.0f / 16777216.0f), &s, &c);
    return make_float2(r * c, r * s);
}

// One CTA per coloured batch: torch_audiomentations `_gen_noise` -- rfft of the 1 s N(0,1) pattern,
// mask 1 / linspace(1, sqrt(sr / 2), bins)^f_decay, irfft, unit RMS -- with the pattern drawn from stream 3 of the batch's
// Philox counters.  f_decay == 0 (white noise) skips the transform pair: the mask is 1.
__global__ void __launch_bounds__(kAugThreads, 1)
colored_bases_kernel(uint2 key, const int64_t* __restrict__ batch_ids, const float* __restrict__ f_decay, float* __restrict__ out,
                     FftPlan plan) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float scratch[40];
    __shared__ TwTables tw;
    constexpr int N = kColoredBase;
    const int M = plan.M;                         // N / 2
    float2* buf0 = reinterpret_cast<float2*>(smem_raw);
    float2* buf1 = buf0 + M;
    const int tid = threadIdx.x;
    const int64_t g = batch_ids[blockIdx.x];
    const float fd = f_decay[blockIdx.x];
    for (int i = tid; i < plan.tw_hi; i += kAugThreads) tw.hi[i] = __ldg(plan.tw_m + i * plan.tw_lo);
    for (int i = tid; i < plan.tw_lo; i += kAugThreads) tw.lo[i] = __ldg(plan.tw_m + i);
    float sumsq = 0.f;
    for (int j = tid; j < N / 4; j += kAugThreads) {
        const uint4 x = philox4x32_10(make_uint4((uint32_t)j, 3u, (uint32_t)g, (uint32_t)((uint64_t)g >> 32)), key);
        const float2 a = box_muller(x.x, x.y), b = box_muller(x.z, x.w);
        buf0[2 * j] = a;
        buf0[2 * j + 1] = b;
        sumsq += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y;
    }
    __syncthreads();
    float2* Y = buf0;
    float scale = 1.0f;
    if (fd != 0.0f) {
        float2* Z = fft_forward(buf0, buf1, plan, tw);
        float2* other = (Z == buf0) ? buf1 : buf0;
        const float slope = (sqrtf((float)(N / 2)) - 1.0f) / (float)(N / 2);      // linspace(1, sqrt(sr / 2), N / 2 + 1)
        auto mask = [&](int k) { return exp2f(-fd * log2f(1.0f + slope * (float)k)); };
        for (int k = tid; k <= M / 2; k += kAugThreads) {
            if (k == 0) {
                const float x0 = Z[0].x + Z[0].y, xm = Z[0].x - Z[0].y;
                const float y0 = x0 * mask(0), ym = xm * mask(M);
                Z[0] = make_float2(0.5f * (y0 + ym), -0.5f * (y0 - ym));
            } else {
                const float2 w = __ldg(plan.tw_t + k);
                float2 Xk, Xmk;
                untangle(Z[k], Z[M - k], w, &Xk, &Xmk);
                const float2 Yk = cscale(Xk, mask(k));
                const float2 Ymk = cscale(Xmk, mask(M - k));
                const float2 Ye = cscale(cadd(Yk, cconj(Ymk)), 0.5f);
                const float2 Yo = cmulf(cscale(csub(Yk, cconj(Ymk)), 0.5f), cconj(w));
                const float2 zk = cadd(Ye, mul_pos_i(Yo));
                const float2 zmk = cadd(cconj(Ye), mul_pos_i(cconj(Yo)));
                Z[k] = cconj(zk);
                if (k != M - k) Z[M - k] = cconj(zmk);
            }
        }
        __syncthreads();
        Y = fft_forward(Z, other, plan, tw);      // conj(M * (y_even + i y_odd))
        sumsq = 0.f;
        const float inv_m = 1.0f / (float)M;
        for (int i = tid; i < M; i += kAugThreads) {
            const float2 v = Y[i];
            const float2 y = make_float2(v.x * inv_m, -v.y * inv_m);
            Y[i] = y;
            sumsq += y.x * y.x + y.y * y.y;
        }
    }
    scale = rsqrtf(block_sum(sumsq, scratch) / (float)N);
    float2* dst = reinterpret_cast<float2*>(out + (int64_t)blockIdx.x * N);
    for (int i = tid; i < M; i += kAugThreads) dst[i] = cscale(Y[i], scale);
}

// ------------------------------------------------------------------------------------------------
// fused per-clip augmentation
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAugThreads, 1)
augment_kernel(const float* __restrict__ clips, const float* __restrict__ noise_bank,
               const float* __restrict__ colored_bases, const float2* __restrict__ rir_specs,
               const hb_clip_aug* __restrict__ params, float* __restrict__ out, int T, FftPlan plan) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float scratch[40];
    __shared__ TwTables tw;
    const int M = plan.M;
    float2* buf0 = reinterpret_cast<float2*>(smem_raw);
    float2* buf1 = buf0 + M;
    float* x = reinterpret_cast<float*>(buf0);
    float* nz = reinterpret_cast<float*>(buf1);
    const int tid = threadIdx.x;
    const hb_clip_aug p = params[blockIdx.x];
    const float* src = clips + (int64_t)blockIdx.x * T;
    float* dst = out + (int64_t)blockIdx.x * T;
    const bool has_colored = p.colored_index >= 0 && colored_bases != nullptr;
    const bool has_noise = p.noise_offset >= 0 && noise_bank != nullptr;
    const bool has_rir = p.rir_index >= 0 && rir_specs != nullptr;
    if (has_rir) {
        for (int i = tid; i < plan.tw_hi; i += kAugThreads) tw.hi[i] = __ldg(plan.tw_m + i * plan.tw_lo);
        for (int i = tid; i < plan.tw_lo; i += kAugThreads) tw.lo[i] = __ldg(plan.tw_m + i);
    }

    // ---- load clip (128-bit loads; T is even, rows are 16-byte aligned when T % 4 == 0) --------
    float sumsq = 0.f;
    if ((T & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* x4 = reinterpret_cast<float4*>(x);
        for (int i = tid; i < T / 4; i += kAugThreads) {
            const float4 v = __ldg(s4 + i);
            x4[i] = v;
            sumsq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
    } else {
        for (int i = tid; i < T; i += kAugThreads) {
            const float v = __ldg(src + i);
            x[i] = v;
            sumsq += v * v;
        }
    }
    // prefetch the noise row into the second buffer while the clip is being processed
    float nsq = 0.f;
    if (has_noise) {
        const float* nsrc = noise_bank + p.noise_offset;
        if ((T & 3) == 0 && ((reinterpret_cast<uintptr_t>(nsrc) & 15) == 0)) {
            const float4* s4 = reinterpret_cast<const float4*>(nsrc);
            float4* n4 = reinterpret_cast<float4*>(nz);
            for (int i = tid; i < T / 4; i += kAugThreads) {
                const float4 v = __ldg(s4 + i);
                n4[i] = v;
                nsq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
        } else {
            for (int i = tid; i < T; i += kAugThreads) {
                const float v = __ldg(nsrc + i);
                nz[i] = v;
                nsq += v * v;
            }
        }
    }
    __syncthreads();

    // ---- K1 coloured noise + K2 gain ---------------------------------------------------------------
    if (has_colored) {
        const float rms = sqrtf(block_sum(sumsq, scratch) / (float)T);
        const float a = rms / exp10f(p.colored_snr_db * 0.05f);
        const float* base = colored_bases + (int64_t)p.colored_index * kColoredBase;
        for (int i = tid; i < T; i += kAugThreads) x[i] = (x[i] + a * __ldg(base + (i % kColoredBase))) * p.gain;
        __syncthreads();
    } else if (p.gain != 1.0f) {
        for (int i = tid; i < T; i += kAugThreads) x[i] *= p.gain;
        __syncthreads();
    }

    // ---- K3 background noise at the per-clip SNR -------------------------------------------------------
    if (has_noise) {
        float e = 0.f;
        for (int i = tid; i < T; i += kAugThreads) e += x[i] * x[i];
        const float e_s = block_sum(e, scratch);
        const float e_n = block_sum(nsq, scratch);
        const float orig = 10.0f * (log10f(e_s) - log10f(e_n));
        const float scale = exp10f((orig - p.noise_snr_db) * 0.05f);
        for (int i = tid; i < T; i += kAugThreads) x[i] = fmaf(scale, nz[i], x[i]);
        __syncthreads();
    }

    if (!has_rir) {
        for (int i = tid; i < T; i += kAugThreads) dst[i] = x[i];
        return;
    }

    // ---- K4 reverb -------------------------------------------------------------------------------------
    float a = 0.f;
    for (int i = tid; i < T; i += kAugThreads) a += fabsf(x[i]);
    const float amp_x = block_sum(a, scratch) / (float)T;

    float2* Z = fft_forward(buf0, buf1, plan, tw);
    float2* other = (Z == buf0) ? buf1 : buf0;
    const float2* H = rir_specs + (int64_t)p.rir_index * (M + 1);
    // untangle -> multiply -> re-tangle, conjugated so the forward FFT inverts it
    for (int k = tid; k <= M / 2; k += kAugThreads) {
        if (k == 0) {
            const float x0 = Z[0].x + Z[0].y, xm = Z[0].x - Z[0].y;
            const float y0 = x0 * __ldg(&H[0]).x, ym = xm * __ldg(&H[M]).x;
            Z[0] = make_float2(0.5f * (y0 + ym), -0.5f * (y0 - ym));
        } else {
            const float2 w = __ldg(plan.tw_t + k);
            float2 Xk, Xmk;
            untangle(Z[k], Z[M - k], w, &Xk, &Xmk);
            const float2 Yk = cmulf(Xk, __ldg(&H[k]));
            const float2 Ymk = cmulf(Xmk, __ldg(&H[M - k]));
            const float2 Ye = cscale(cadd(Yk, cconj(Ymk)), 0.5f);
            const float2 Yo = cmulf(cscale(csub(Yk, cconj(Ymk)), 0.5f), cconj(w));
            // Z'[k] = Ye + i Yo ; Z'[M-k] = conj(Ye) + i Yo2, Yo2 = (Ymk - conj(Yk))/2 * conj(W^{M-k}) = conj(Yo)
            const float2 zk = cadd(Ye, mul_pos_i(Yo));
            const float2 zmk = cadd(cconj(Ye), mul_pos_i(cconj(Yo)));
            Z[k] = cconj(zk);
            if (k != M - k) Z[M - k] = cconj(zmk);
        }
    }
    __syncthreads();
    float2* Yc = fft_forward(Z, other, plan, tw);   // = conj(M * z'), z' = y_even + i y_odd
    float* y = reinterpret_cast<float*>(Yc);
    const float inv_m = 1.0f / (float)M;
    float ay = 0.f;
    for (int i = tid; i < M; i += kAugThreads) {
        const float2 v = Yc[i];
        const float re = v.x * inv_m, im = -v.y * inv_m;
        Yc[i] = make_float2(re, im);
        ay += fabsf(re) + fabsf(im);
    }
    const float amp_y = block_sum(ay, scratch) / (float)T;
    const float g = amp_x / (amp_y + 1e-14f);
    for (int i = tid; i < T; i += kAugThreads) dst[i] = y[i] * g;
}

// ------------------------------------------------------------------------------------------------
// fused per-clip augmentation, T = 23040 (the 1.44 s clip of the featurization path)
// ------------------------------------------------------------------------------------------------
// Same arithmetic as augment_kernel, restructured around the FFT: M = 11520 = 16 * 16 * 9 * 5 with the
// butterflies held in registers, so the reverb costs 7 shared-memory passes instead of 15:
//     forward  R16 (Ns 1) -> R16 (Ns 16) -> R9 (Ns 256)
//     middle   R5 (Ns 2304) + untangle * H + re-tangle + transposed R5, in registers: the last forward pass leaves
//              X[k + 2304 r] of butterfly k in one thread and X[M - k - 2304 r] in butterfly 2304 - k, so a thread
//              that owns both butterflies owns every (k, M - k) pair it needs
//     inverse  transposed R9 -> R16 -> R16 (the FFT matrix is symmetric: running the transposed passes in
//              reverse order on conjugated data is the inverse transform, natural order in and out)
// 768 threads: 720 radix-16 butterflies per pass (94 % of the threads busy), 1280 radix-9, 1153 middle tasks.
// Shared-memory index i lives at i + i / 16 (float2 units): a 64-bit access is served one half-warp at a time, and with this
// skew the 16 lanes of a half-warp hit 16 different bank pairs in the stride-16 stores of the first pass and the stride-16
// loads of the last (i + i / 32 left them two-way conflicted: 1440 instead of 720 wavefronts per pass).
// profiling aid: cycle stamps of the fused kernel's phases for the CTAs that run on the same SMs as CTAs 0..7 one wave later
static __device__ long long g_aug_times[8][8];
#define AUG_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x >= 1184 && blockIdx.x < 1192) g_aug_times[blockIdx.x - 1184][i] = clock64(); } while (0)

constexpr int kFastT = 23040;
constexpr int kFastM = kFastT / 2;
constexpr int kFastThreads = 768;
constexpr int kFastBuf = kFastM + kFastM / 16;     // skewed float2 slots per buffer

__device__ __forceinline__ int sk(int i) { return i + (i >> 4); }

template <int R> __host__ __device__ constexpr int out_idx(int k) { return k; }                 // where dftr<R> leaves X[k]
template <> __host__ __device__ constexpr int out_idx<16>(int k) { return 4 * (k & 3) + (k >> 2); }
template <> __host__ __device__ constexpr int out_idx<9>(int k) { return 3 * (k % 3) + k / 3; }

template <int R>
__device__ __forceinline__ void dftr(float2 (&v)[R]);

template <>
__device__ __forceinline__ void dftr<5>(float2 (&v)[5]) { dft<5>(v); }

__device__ __forceinline__ void dft3r(float2& v0, float2& v1, float2& v2) {
    const float2 t1 = cadd(v1, v2);
    const float2 m1 = make_float2(v0.x - 0.5f * t1.x, v0.y - 0.5f * t1.y);
    const float2 t2 = cscale(csub(v1, v2), 0.86602540378443864676f);
    v0 = cadd(v0, t1);
    v1 = cadd(m1, mul_neg_i(t2));
    v2 = cadd(m1, mul_pos_i(t2));
}
__device__ __forceinline__ void dft4r(float2& v0, float2& v1, float2& v2, float2& v3) {
    const float2 a0 = cadd(v0, v2), a1 = csub(v0, v2);
    const float2 a2 = cadd(v1, v3), a3 = csub(v1, v3);
    v0 = cadd(a0, a2);
    v2 = csub(a0, a2);
    v1 = cadd(a1, mul_neg_i(a3));
    v3 = cadd(a1, mul_pos_i(a3));
}

template <>
__device__ __forceinline__ void dftr<9>(float2 (&v)[9]) {
    // 9 = 3 x 3: v[a + 3 kb] = Y[a][kb], twiddles W9^(a kb), then v[3 kb + ka] = X[kb + 3 ka]
#pragma unroll
    for (int a = 0; a < 3; ++a) dft3r(v[a], v[a + 3], v[a + 6]);
    v[4] = cmulf(v[4], make_float2(0.76604444311897803520f, -0.64278760968653932632f));   // W9^1
    v[5] = cmulf(v[5], make_float2(0.17364817766693034885f, -0.98480775301220805937f));   // W9^2
    v[7] = cmulf(v[7], make_float2(0.17364817766693034885f, -0.98480775301220805937f));   // W9^2
    v[8] = cmulf(v[8], make_float2(-0.93969262078590838405f, -0.34202014332566873304f));  // W9^4
#pragma unroll
    for (int kb = 0; kb < 3; ++kb) dft3r(v[3 * kb], v[3 * kb + 1], v[3 * kb + 2]);
}

template <>
__device__ __forceinline__ void dftr<16>(float2 (&v)[16]) {
    constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
#pragma unroll
    for (int a = 0; a < 4; ++a) dft4r(v[a], v[a + 4], v[a + 8], v[a + 12]);
    v[5] = cmulf(v[5], make_float2(c1, -s1));
    v[6] = make_float2(h * (v[6].x + v[6].y), h * (v[6].y - v[6].x));
    v[7] = cmulf(v[7], make_float2(s1, -c1));
    v[9] = make_float2(h * (v[9].x + v[9].y), h * (v[9].y - v[9].x));
    v[10] = mul_neg_i(v[10]);
    v[11] = make_float2(h * (v[11].y - v[11].x), -h * (v[11].x + v[11].y));
    v[13] = cmulf(v[13], make_float2(s1, -c1));
    v[14] = make_float2(h * (v[14].y - v[14].x), -h * (v[14].x + v[14].y));
    v[15] = cmulf(v[15], make_float2(-c1, s1));
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) dft4r(v[4 * kb], v[4 * kb + 1], v[4 * kb + 2], v[4 * kb + 3]);
}

// w[r] = w1^r, r = 1..R-1, multiplication depth <= log2(R) + 2
template <int R>
__device__ __forceinline__ void powers(float2 w1, float2 (&w)[R]) {
    w[0] = make_float2(1.f, 0.f);
    w[1] = w1;
#pragma unroll
    for (int r = 2; r < R; ++r) w[r] = (r & 1) ? cmulf(w[r - 1], w1) : cmulf(w[r / 2], w[r / 2]);
}

// forward Stockham pass: gather in[j + r nb], input-side twiddles, DFT, scatter out[(j - k) R + k + r Ns]
template <int R, int NS>
__device__ __forceinline__ void fast_pass_fwd(const float2* __restrict__ in, float2* __restrict__ out, const TwTables& tw,
                                              int lo_n) {
    constexpr int nb = kFastM / R;
    constexpr int step = kFastM / (NS * R);
    for (int j = threadIdx.x; j < nb; j += kFastThreads) {
        const int k = j % NS;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[sk(j + r * nb)];
        if (NS > 1) {
            float2 w[R];
            powers<R>(twiddle(tw, k * step, lo_n), w);
#pragma unroll
            for (int r = 1; r < R; ++r) v[r] = cmulf(v[r], w[r]);
        }
        dftr<R>(v);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) out[sk(base + r * NS)] = v[out_idx<R>(r)];
    }
    __syncthreads();
}

// the transpose of fast_pass_fwd<R, NS>: gather in[(j - k) R + k + r Ns], DFT, output-side twiddles, scatter out[j + r nb]
template <int R, int NS>
__device__ __forceinline__ void fast_pass_bwd(const float2* __restrict__ in, float2* __restrict__ out, const TwTables& tw,
                                              int lo_n) {
    constexpr int nb = kFastM / R;
    constexpr int step = kFastM / (NS * R);
    for (int j = threadIdx.x; j < nb; j += kFastThreads) {
        const int k = j % NS;
        const int base = (j - k) * R + k;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = in[sk(base + r * NS)];
        dftr<R>(v);
        if (NS > 1) {
            float2 w[R];
            powers<R>(twiddle(tw, k * step, lo_n), w);
            out[sk(j)] = v[out_idx<R>(0)];
#pragma unroll
            for (int r = 1; r < R; ++r) out[sk(j + r * nb)] = cmulf(v[out_idx<R>(r)], w[r]);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) out[sk(j + r * nb)] = v[out_idx<R>(r)];
        }
    }
    __syncthreads();
}

// (Z[k], Z[M-k]) -> conj of the re-tangled (Z'[k], Z'[M-k]) after multiplying the real spectrum by H
__device__ __forceinline__ void pair_update(float2& zk, float2& zmk, int k, const float2* __restrict__ H,
                                            const float2* __restrict__ tw_t) {
    const float2 w = __ldg(tw_t + k);
    float2 Xk, Xmk;
    untangle(zk, zmk, w, &Xk, &Xmk);
    const float2 Yk = cmulf(Xk, __ldg(H + k));
    const float2 Ymk = cmulf(Xmk, __ldg(H + (kFastM - k)));
    const float2 Ye = cscale(cadd(Yk, cconj(Ymk)), 0.5f);
    const float2 Yo = cmulf(cscale(csub(Yk, cconj(Ymk)), 0.5f), cconj(w));
    zk = cconj(cadd(Ye, mul_pos_i(Yo)));
    zmk = cconj(cadd(cconj(Ye), mul_pos_i(cconj(Yo))));
}

// last forward pass (R5, Ns = 2304) + spectrum product + first inverse pass, in place
__device__ __forceinline__ void fast_middle(float2* __restrict__ buf, const float2* __restrict__ H,
                                            const float2* __restrict__ tw_t, const TwTables& tw, int lo_n) {
    constexpr int NS = kFastM / 5;   // 2304 = nb
    for (int t = threadIdx.x; t <= NS / 2; t += kFastThreads) {
        float2 a[5], b[5], wa[5], wb[5];
        const int ja = t, jb = (t == 0) ? 0 : NS - t;
        powers<5>(twiddle(tw, ja, lo_n), wa);
#pragma unroll
        for (int r = 0; r < 5; ++r) a[r] = buf[sk(ja + r * NS)];
#pragma unroll
        for (int r = 1; r < 5; ++r) a[r] = cmulf(a[r], wa[r]);
        dftr<5>(a);
        if (t == 0) {
            // butterfly 0 pairs with itself: (r, 5 - r); bin 0 carries DC and Nyquist
            const float x0 = a[0].x + a[0].y, xm = a[0].x - a[0].y;
            const float y0 = x0 * __ldg(&H[0]).x, ym = xm * __ldg(&H[kFastM]).x;
            a[0] = make_float2(0.5f * (y0 + ym), -0.5f * (y0 - ym));
            pair_update(a[1], a[4], NS, H, tw_t);
            pair_update(a[2], a[3], 2 * NS, H, tw_t);
        } else if (t == NS / 2) {
            // butterfly 1152 pairs with itself: (r, 4 - r); r = 2 is bin M / 2
            pair_update(a[0], a[4], t, H, tw_t);
            pair_update(a[1], a[3], t + NS, H, tw_t);
            float2 dup = a[2];
            pair_update(a[2], dup, t + 2 * NS, H, tw_t);
        } else {
            // W_M^(2304 - t) = W5 * conj(W_M^t)
            powers<5>(cmulf(make_float2(0.30901699437494742410f, -0.95105651629515357212f), cconj(wa[1])), wb);
#pragma unroll
            for (int r = 0; r < 5; ++r) b[r] = buf[sk(jb + r * NS)];
#pragma unroll
            for (int r = 1; r < 5; ++r) b[r] = cmulf(b[r], wb[r]);
            dftr<5>(b);
#pragma unroll
            for (int r = 0; r < 5; ++r) pair_update(a[r], b[4 - r], t + r * NS, H, tw_t);
            dftr<5>(b);
            buf[sk(jb)] = b[0];
#pragma unroll
            for (int r = 1; r < 5; ++r) buf[sk(jb + r * NS)] = cmulf(b[r], wb[r]);
        }
        dftr<5>(a);
        buf[sk(ja)] = a[0];
#pragma unroll
        for (int r = 1; r < 5; ++r) buf[sk(ja + r * NS)] = cmulf(a[r], wa[r]);
    }
    __syncthreads();
}

__device__ __forceinline__ float block_sum_n(float v, float* scratch, int n_warps) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();  // protect scratch reuse
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = (lane < n_warps) ? scratch[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// bulk L2 prefetch of the byte range [p, p + bytes) widened to 16-byte boundaries
__device__ __forceinline__ void prefetch_range_l2(const void* p, int64_t bytes) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)15;
    const uint32_t n = (uint32_t)(((reinterpret_cast<uintptr_t>(p) + (uintptr_t)bytes + 15) & ~(uintptr_t)15) - a);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(n) : "memory");
}

// kI16: the length fix (a1: /32768, front-truncate or left-pad by pad_before) is fused into the load -- the clip comes
// straight from the ragged int16 samples and the f32 [n][T] intermediate never exists in HBM.
// kMel (production mode): the augmented clip never leaves the SM either -- it is left in shared memory
// (natural order, in the buffer the FFT no longer needs) and 16 of the CTA's 24 warps turn it into the clip's 141 log-mel frames
// with the same frame-pair routine as the stand-alone mel kernel (mel_core.cuh: bit-identical frames); `out` receives
// [n][141][32] mel instead of [n][T] audio.  HBM traffic per clip: source + noise row + 18 KB of mel.
constexpr int kFusedFrames = 1 + (kFastT - kNFFT) / kHop;   // 141 frames = 71 pairs = 3 rounds of the CTA's 24 warps
constexpr int kFusedTr = 2 * 16 * kTrStride;          // float2 per warp: two transpose tiles (the power rows alias them)
constexpr int kFusedWarpsInBuf = (kFastBuf / kFusedTr) < (kFastThreads / 32) ? (kFastBuf / kFusedTr) : (kFastThreads / 32);   // 22 in the freed FFT buffer
static_assert(2 * kPowerRow * sizeof(float) <= kFusedTr * sizeof(float2), "the power rows alias the transpose tiles");

template <bool kI16, bool kMel>
__global__ void __launch_bounds__(kFastThreads, 1)
augment_fast_kernel(const float* __restrict__ clips, const int16_t* __restrict__ samples, const int64_t* __restrict__ offsets,
                    const int32_t* __restrict__ pad_before, const float* __restrict__ noise_bank,
                    const float* __restrict__ colored_bases, const float2* __restrict__ rir_specs,
                    const hb_clip_aug* __restrict__ params, float* __restrict__ out, FftPlan plan, int prefetch_ahead,
                    const MelTables* __restrict__ mel_tables, float mel_scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float scratch[40];
    __shared__ TwTables tw;
    __shared__ MelShared mel_s;                 // kMel only (4.9 KB)
    __shared__ float2 mel_w256[kMel ? 256 : 1];  // W256^k: the fused kernel has no registers to spare for a lane's 16 twiddles
    __shared__ float2 mel_extra[kMel ? (kFastThreads / 32 - kFusedWarpsInBuf) * kFusedTr : 1];   // scratch of the warps that do not fit the FFT buffer
    if (kMel) {
        mel_load_shared(mel_s, *mel_tables, threadIdx.x, kFastThreads);
        for (int i = threadIdx.x; i < 256; i += kFastThreads) mel_w256[i] = mel_tables->w256[i];
    }
    constexpr int T = kFastT, M = kFastM, NT = kFastThreads, NW = kFastThreads / 32;
    AUG_STAMP(0);
    float2* buf0 = reinterpret_cast<float2*>(smem_raw);   // skewed: point n at sk(n)
    float2* buf1 = buf0 + kFastBuf;
    const int tid = threadIdx.x;
    const hb_clip_aug p = params[blockIdx.x];
    const float2* src = kI16 ? nullptr : reinterpret_cast<const float2*>(clips + (int64_t)blockIdx.x * T);
    float2* dst = reinterpret_cast<float2*>(out + (int64_t)blockIdx.x * T);
    const bool has_colored = p.colored_index >= 0 && colored_bases != nullptr;
    const bool has_noise = p.noise_offset >= 0 && noise_bank != nullptr;
    const bool has_rir = p.rir_index >= 0 && rir_specs != nullptr;
    if (has_rir) {
        for (int i = tid; i < plan.tw_hi; i += NT) tw.hi[i] = __ldg(plan.tw_m + i * plan.tw_lo);
        for (int i = tid; i < plan.tw_lo; i += NT) tw.lo[i] = __ldg(plan.tw_m + i);
    }

    // ---- load the clip and (prefetch) its noise row; both stay in shared memory ------------------------
    float sumsq = 0.f, nsq = 0.f;
    if (kI16) {
        const int64_t s0 = offsets[blockIdx.x];
        const int len = (int)(offsets[blockIdx.x + 1] - s0);
        const int pad = len >= T ? 0 : pad_before[blockIdx.x];
        const int16_t* s = samples + s0 - pad;
        // all 15 loads of a thread in flight at once; sample pairs as one 32-bit load when the clip starts on an even sample
        if ((reinterpret_cast<uintptr_t>(s) & 3) == 0) {
            const uint32_t* s2 = reinterpret_cast<const uint32_t*>(s);
            uint32_t raw[M / NT];
#pragma unroll
            for (int i = 0; i < M / NT; ++i) {
                const int n = tid + i * NT, j = 2 * n - pad;
                const bool lo_ok = j >= 0 && j < len, hi_ok = j + 1 >= 0 && j + 1 < len;
                raw[i] = (lo_ok && hi_ok) ? __ldg(s2 + n)
                                          : ((lo_ok ? (uint32_t)(uint16_t)__ldg(s + 2 * n) : 0u) | (hi_ok ? ((uint32_t)(uint16_t)__ldg(s + 2 * n + 1) << 16) : 0u));
            }
#pragma unroll
            for (int i = 0; i < M / NT; ++i) {
                const int n = tid + i * NT;
                const float2 v = make_float2((float)(int16_t)(raw[i] & 0xffffu) * (1.0f / 32768.0f), (float)(int16_t)(raw[i] >> 16) * (1.0f / 32768.0f));
                buf0[sk(n)] = v;
                sumsq += v.x * v.x + v.y * v.y;
            }
        } else {
            int16_t r0[M / NT], r1[M / NT];
#pragma unroll
            for (int i = 0; i < M / NT; ++i) {
                const int n = tid + i * NT, j = 2 * n - pad;
                r0[i] = (j >= 0 && j < len) ? __ldg(s + 2 * n) : (int16_t)0;
                r1[i] = (j + 1 >= 0 && j + 1 < len) ? __ldg(s + 2 * n + 1) : (int16_t)0;
            }
#pragma unroll
            for (int i = 0; i < M / NT; ++i) {
                const int n = tid + i * NT;
                const float2 v = make_float2((float)r0[i] * (1.0f / 32768.0f), (float)r1[i] * (1.0f / 32768.0f));
                buf0[sk(n)] = v;
                sumsq += v.x * v.x + v.y * v.y;
            }
        }
    } else {
#pragma unroll 5
        for (int n = tid; n < M; n += NT) {
            const float2 v = __ldg(src + n);
            buf0[sk(n)] = v;
            sumsq += v.x * v.x + v.y * v.y;
        }
    }
    if (has_noise) {
        const float* nsrc = noise_bank + p.noise_offset;
        if ((reinterpret_cast<uintptr_t>(nsrc) & 7) == 0) {
            const float2* n2 = reinterpret_cast<const float2*>(nsrc);
            float2 nv[M / NT];
#pragma unroll
            for (int i = 0; i < M / NT; ++i) nv[i] = __ldg(n2 + tid + i * NT);     // 15 loads in flight
#pragma unroll
            for (int i = 0; i < M / NT; ++i) {
                buf1[tid + i * NT] = nv[i];
                nsq += nv[i].x * nv[i].x + nv[i].y * nv[i].y;
            }
        } else {
            for (int n = tid; n < M; n += NT) {
                const float2 v = make_float2(__ldg(nsrc + 2 * n), __ldg(nsrc + 2 * n + 1));
                buf1[n] = v;
                nsq += v.x * v.x + v.y * v.y;
            }
        }
    }

#ifndef HB_AUG_NO_PREFETCH
    // One CTA per SM and one clip per CTA: the load phase above cannot overlap this CTA's own FFT.  CTAs are dispatched in
    // order, so the clip that will follow on this SM is about one wave (gridDim-independent: prefetch_ahead CTAs) ahead: start
    // its source samples and noise row on their way to L2 now.
    if (tid == 0 && prefetch_ahead > 0 && (int)blockIdx.x + prefetch_ahead < (int)gridDim.x) {
        const int nb = (int)blockIdx.x + prefetch_ahead;
        if (kI16) {
            const int64_t q0 = offsets[nb], q1 = offsets[nb + 1];
            if (q1 > q0) prefetch_range_l2(samples + q0, (q1 - q0) * 2);
        } else {
            prefetch_range_l2(clips + (int64_t)nb * T, (int64_t)T * 4);
        }
        const int64_t noff = params[nb].noise_offset;
        if (noff >= 0 && noise_bank != nullptr) prefetch_range_l2(noise_bank + noff, (int64_t)T * 4);
    }
#endif
    AUG_STAMP(1);
    // every thread only ever touches its own points n = tid + i NT until the FFT: no barriers needed in between
    // ---- K1 coloured noise + K2 gain + K3 background noise: at most two passes over the clip --------------------
    // A gain that is not followed by coloured noise is not applied on its own: the energy K3 needs is gain^2 * sum(x^2) and
    // the factor rides the mixing pass.  The mean |x| the reverb rescaling needs is accumulated by whichever pass is last.
    float e = 0.f, a = 0.f, pending = 1.0f;
    bool have_abs = false;
    if (has_colored) {
        const float rms = sqrtf(block_sum_n(sumsq, scratch, NW) / (float)T);
        const float ac = rms / exp10f(p.colored_snr_db * 0.05f);
        const float* base = colored_bases + (int64_t)p.colored_index * kColoredBase;
        for (int n = tid; n < M; n += NT) {
            const float2 c = __ldg(reinterpret_cast<const float2*>(base + ((2 * n) % kColoredBase)));
            float2 v = buf0[sk(n)];
            v.x = (v.x + ac * c.x) * p.gain;
            v.y = (v.y + ac * c.y) * p.gain;
            buf0[sk(n)] = v;
            e += v.x * v.x + v.y * v.y;
            a += fabsf(v.x) + fabsf(v.y);
        }
        have_abs = true;
    } else {
        pending = p.gain;
    }

    if (has_noise) {
        const float e_s = has_colored ? block_sum_n(e, scratch, NW) : p.gain * p.gain * block_sum_n(sumsq, scratch, NW);
        const float e_n = block_sum_n(nsq, scratch, NW);
        const float orig = 10.0f * (log10f(e_s) - log10f(e_n));
        const float scale = exp10f((orig - p.noise_snr_db) * 0.05f);
        a = 0.f;
        for (int n = tid; n < M; n += NT) {
            float2 v = buf0[sk(n)];
            const float2 z = buf1[n];
            v.x = fmaf(scale, z.x, v.x * pending);
            v.y = fmaf(scale, z.y, v.y * pending);
            buf0[sk(n)] = v;
            a += fabsf(v.x) + fabsf(v.y);
        }
        have_abs = true;
    } else if (pending != 1.0f) {
        a = 0.f;
        for (int n = tid; n < M; n += NT) {
            float2 v = buf0[sk(n)];
            v.x *= pending;
            v.y *= pending;
            buf0[sk(n)] = v;
            a += fabsf(v.x) + fabsf(v.y);
        }
        have_abs = true;
    }

    AUG_STAMP(2);
    float2* fin = buf1;          // kMel: the finished clip, natural order (buf1 is free once the noise row has been mixed in)
    if (!has_rir) {
        if (kMel) {
            for (int n = tid; n < M; n += NT) fin[n] = buf0[sk(n)];     // a thread re-uses only the points it read the noise from
        } else {
            for (int n = tid; n < M; n += NT) dst[n] = buf0[sk(n)];
            return;
        }
    } else {
    // ---- K4 reverb -------------------------------------------------------------------------------------
    if (!have_abs) {
        for (int n = tid; n < M; n += NT) {
            const float2 v = buf0[sk(n)];
            a += fabsf(v.x) + fabsf(v.y);
        }
    }
    const float amp_x = block_sum_n(a, scratch, NW) / (float)T;   // its barriers also publish buf0 and tw
    AUG_STAMP(3);

    const float2* H = rir_specs + (int64_t)p.rir_index * (M + 1);
    const int lo_n = plan.tw_lo;
    fast_pass_fwd<16, 1>(buf0, buf1, tw, lo_n);
    fast_pass_fwd<16, 16>(buf1, buf0, tw, lo_n);
    fast_pass_fwd<9, 256>(buf0, buf1, tw, lo_n);
    fast_middle(buf1, H, plan.tw_t, tw, lo_n);
    fast_pass_bwd<9, 256>(buf1, buf0, tw, lo_n);
    fast_pass_bwd<16, 16>(buf0, buf1, tw, lo_n);
    fast_pass_bwd<16, 1>(buf1, buf0, tw, lo_n);      // buf0 = conj(M * (y_even + i y_odd)), natural order
    AUG_STAMP(4);

    const float inv_m = 1.0f / (float)M;
    float ay = 0.f;
    float2 y[M / NT];
#pragma unroll
    for (int i = 0; i < M / NT; ++i) {
        const float2 v = buf0[sk(tid + i * NT)];
        y[i] = make_float2(v.x * inv_m, -v.y * inv_m);
        ay += fabsf(y[i].x) + fabsf(y[i].y);
    }
    const float amp_y = block_sum_n(ay, scratch, NW) / (float)T;
    const float g = amp_x / (amp_y + 1e-14f);
#pragma unroll
    for (int i = 0; i < M / NT; ++i) (kMel ? fin : dst)[tid + i * NT] = make_float2(y[i].x * g, y[i].y * g);
    }   // has_rir
    if (!kMel) return;

    // ---- K6 in place: the clip's 141 log-mel frames from shared memory -------------------------------------------
    __syncthreads();                                   // fin complete, buf0 free, mel tables loaded
    AUG_STAMP(5);
    const int warp = tid >> 5, lane = tid & 31;
    float2* tr_pair = warp < kFusedWarpsInBuf ? buf0 + warp * kFusedTr : mel_extra + (warp - kFusedWarpsInBuf) * kFusedTr;
    const int l = lane & 15;
    const int my_lo = mel_s.lo[lane], my_rot = mel_s.rot[lane];
    float* mel_clip = out + (int64_t)blockIdx.x * kFusedFrames * kMels;
    constexpr int kPairs = (kFusedFrames + 1) / 2;
    for (int pr = warp; pr < kPairs; pr += NW) {
        const int f0 = 2 * pr;
        auto load = [&](int h, int n) { return fin[(kHop / 2) * min(f0 + h, kFusedFrames - 1) + n]; };
        mel_frame_pair(mel_s, tr_pair, reinterpret_cast<float*>(tr_pair), [&](int k1) { return mel_w256[(l * k1) & 255]; }, my_lo, my_rot, load,
                       mel_scale, f0, kFusedFrames, mel_clip);
    }
    __syncthreads();
    AUG_STAMP(6);
}

// ---- K9: BandStopFilter ---------------------------------------------------------------------------------------------------------
// torch_audiomentations.BandStopFilter (reference augmented.py:102-106; restated, parity unpinned: heybuddy_b200/dataset/k9.py):
// x - (lowpass_high(x) - lowpass_low(x)) with julius' Hann-windowed-sinc low-pass pair, both 2 h + 1 taps, on the clip replicate-padded
// by h.  With d = f_high - f_low (one FIR per augmentation batch): y[n] = x[n] - sum_k d[k] xp[n + 2 h - k], xp[i] = x[clamp(i - h)].
// Uniformly partitioned overlap-save on the clip's own exact-length FFT: blocks of T samples, partitions of B = T / 2 taps.  Block
// (c, p) is u[i] = xp[(c - p - 1) B + 2 h + i], i < T; its circular convolution with partition p (spectrum from hb_rir_spectrum of
// the zero-padded taps) is exact for i in [B, 2 B) = partition p's share of the outputs [c B, (c + 1) B).  h <= B / 2 (a low cut-off
// above 11 Hz at 16 kHz, ~98 % of the draws) is one partition: two blocks = two FFT pairs per clip.
template <bool kFast>
__global__ void __launch_bounds__(kFast ? kFastThreads : kAugThreads, 1)
bandstop_kernel(float* __restrict__ clips, const int32_t* __restrict__ clip_index, const int32_t* __restrict__ meta,
                const float2* __restrict__ specs, const float* __restrict__ scratch, int T, FftPlan plan) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ TwTables tw;
    constexpr int NT = kFast ? kFastThreads : kAugThreads;
    const int M = plan.M, B = M, tid = threadIdx.x;
    float2* buf0 = reinterpret_cast<float2*>(smem_raw);
    float2* buf1 = buf0 + (kFast ? kFastBuf : M);
    auto at = [](int n) { return kFast ? sk(n) : n; };
    for (int i = tid; i < plan.tw_hi; i += NT) tw.hi[i] = __ldg(plan.tw_m + i * plan.tw_lo);
    for (int i = tid; i < plan.tw_lo; i += NT) tw.lo[i] = __ldg(plan.tw_m + i);
    const float* x = scratch + (int64_t)blockIdx.x * T;                 // the untouched copy of the clip
    float* dst = clips + (int64_t)clip_index[blockIdx.x] * T;
    const int row0 = meta[3 * blockIdx.x], n_part = meta[3 * blockIdx.x + 1], h = meta[3 * blockIdx.x + 2];
    const float inv_m = 1.0f / (float)M;
    for (int c = 0; c < 2; ++c)
        for (int p = 0; p < n_part; ++p) {
            const int a = (c - p - 1) * B + h;                          // u[i] = x[clamp(a + i, 0, T - 1)]
            for (int n = tid; n < M; n += NT) {
                const int i0 = min(max(a + 2 * n, 0), T - 1), i1 = min(max(a + 2 * n + 1, 0), T - 1);
                buf0[at(n)] = make_float2(__ldg(x + i0), __ldg(x + i1));
            }
            __syncthreads();
            const float2* H = specs + (int64_t)(row0 + p) * (M + 1);
            const float2* res;
            if constexpr (kFast) {
                const int lo_n = plan.tw_lo;
                fast_pass_fwd<16, 1>(buf0, buf1, tw, lo_n);
                fast_pass_fwd<16, 16>(buf1, buf0, tw, lo_n);
                fast_pass_fwd<9, 256>(buf0, buf1, tw, lo_n);
                fast_middle(buf1, H, plan.tw_t, tw, lo_n);
                fast_pass_bwd<9, 256>(buf1, buf0, tw, lo_n);
                fast_pass_bwd<16, 16>(buf0, buf1, tw, lo_n);
                fast_pass_bwd<16, 1>(buf1, buf0, tw, lo_n);             // buf0 = conj(M * (y_even + i y_odd)), natural order
                res = buf0;
            } else {
                float2* Z = fft_forward(buf0, buf1, plan, tw);
                float2* other = (Z == buf0) ? buf1 : buf0;
                for (int k = tid; k <= M / 2; k += NT) {
                    if (k == 0) {
                        const float x0 = Z[0].x + Z[0].y, xm = Z[0].x - Z[0].y;
                        const float y0 = x0 * __ldg(&H[0]).x, ym = xm * __ldg(&H[M]).x;
                        Z[0] = make_float2(0.5f * (y0 + ym), -0.5f * (y0 - ym));
                    } else {
                        const float2 w = __ldg(plan.tw_t + k);
                        float2 Xk, Xmk;
                        untangle(Z[k], Z[M - k], w, &Xk, &Xmk);
                        const float2 Yk = cmulf(Xk, __ldg(&H[k]));
                        const float2 Ymk = cmulf(Xmk, __ldg(&H[M - k]));
                        const float2 Ye = cscale(cadd(Yk, cconj(Ymk)), 0.5f);
                        const float2 Yo = cmulf(cscale(csub(Yk, cconj(Ymk)), 0.5f), cconj(w));
                        const float2 zk = cadd(Ye, mul_pos_i(Yo));
                        const float2 zmk = cadd(cconj(Ye), mul_pos_i(cconj(Yo)));
                        Z[k] = cconj(zk);
                        if (k != M - k) Z[M - k] = cconj(zmk);
                    }
                }
                __syncthreads();
                res = fft_forward(Z, other, plan, tw);
            }
            // v[i], i in [B, 2 B): pair mm = i / 2 in [M / 2, M) -> outputs n = c B + 2 mm - B, n + 1 (the same thread every time)
            for (int mm = M / 2 + tid; mm < M; mm += NT) {
                const float2 v = res[at(mm)];
                const int n = c * B + 2 * mm - B;
                float2 o = p == 0 ? make_float2(__ldg(x + n), __ldg(x + n + 1)) : *reinterpret_cast<const float2*>(dst + n);
                o.x -= v.x * inv_m;
                o.y += v.y * inv_m;
                *reinterpret_cast<float2*>(dst + n) = o;
            }
            __syncthreads();
        }
}

__global__ void gather_rows_kernel(const float* __restrict__ clips, const int32_t* __restrict__ clip_index, float* __restrict__ out, int T) {
    const float* src = clips + (int64_t)clip_index[blockIdx.x] * T;
    float* dst = out + (int64_t)blockIdx.x * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x) dst[i] = src[i];
}

__global__ void fix_length_kernel(const int16_t* __restrict__ samples, const int64_t* __restrict__ offsets,
                                  const int32_t* __restrict__ pad_before, float* __restrict__ out, int T) {
    const int b = blockIdx.x;
    const int64_t s0 = offsets[b];
    const int n = (int)(offsets[b + 1] - s0);
    const int pad = n >= T ? 0 : pad_before[b];
    float* dst = out + (int64_t)b * T;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const int j = i - pad;
        dst[i] = (j >= 0 && j < n) ? (float)samples[s0 + j] * (1.0f / 32768.0f) : 0.f;
    }
}

// ---- host-side plan cache ---------------------------------------------------------------------------
struct PlanKey {
    int dev, T;
    bool operator<(const PlanKey& o) const { return dev != o.dev ? dev < o.dev : T < o.T; }
};
static std::map<PlanKey, FftPlan> g_plans;
static std::mutex g_plan_mutex;

static int get_plan(int T, FftPlan* out) {
    HB_REQUIRE(T >= 4 && T % 2 == 0 && T <= kMaxT, "augment: T=%d unsupported (even, <= %d)", T, kMaxT);
    int dev = 0;
    HB_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    auto it = g_plans.find({dev, T});
    if (it != g_plans.end()) {
        *out = it->second;
        return HB_OK;
    }
    FftPlan plan;
    plan.M = T / 2;
    plan.n_passes = 0;
    int n = plan.M;
    const int order[4] = {4, 2, 3, 5};
    for (int r : order)
        while (n % r == 0) {
            HB_REQUIRE(plan.n_passes < kMaxPasses, "augment: too many FFT passes for T=%d", T);
            plan.radix[plan.n_passes++] = r;
            n /= r;
        }
    HB_REQUIRE(n == 1, "augment: T/2=%d must factor into 2,3,5 for the exact-length FFT", plan.M);
    plan.tw_lo = 1;
    while (plan.tw_lo * plan.tw_lo < plan.M) ++plan.tw_lo;          // ceil(sqrt(M)): 108 for M = 11520
    plan.tw_hi = (plan.M + plan.tw_lo - 1) / plan.tw_lo;
    HB_REQUIRE(plan.tw_lo <= kTwMax && plan.tw_hi <= kTwMax, "augment: twiddle tables too small for T=%d", T);
    std::vector<float2> twm(plan.M), twt(plan.M + 1);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < plan.M; ++k) {
        twm[k].x = (float)cos(two_pi * k / plan.M);
        twm[k].y = (float)-sin(two_pi * k / plan.M);
    }
    for (int k = 0; k <= plan.M; ++k) {
        twt[k].x = (float)cos(two_pi * k / T);
        twt[k].y = (float)-sin(two_pi * k / T);
    }
    float2 *d_m = nullptr, *d_t = nullptr;
    HB_CUDA_OK(cudaMalloc(&d_m, twm.size() * sizeof(float2)));
    HB_CUDA_OK(cudaMalloc(&d_t, twt.size() * sizeof(float2)));
    HB_CUDA_OK(cudaMemcpy(d_m, twm.data(), twm.size() * sizeof(float2), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(d_t, twt.data(), twt.size() * sizeof(float2), cudaMemcpyHostToDevice));
    plan.tw_m = d_m;
    plan.tw_t = d_t;
    HB_CUDA_OK(cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (kMaxT / 2) * (int)sizeof(float2)));
    HB_CUDA_OK(cudaFuncSetAttribute(rir_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (kMaxT / 2) * (int)sizeof(float2)));
    HB_CUDA_OK(cudaFuncSetAttribute(colored_bases_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (kColoredBase / 2) * (int)sizeof(float2)));
    HB_CUDA_OK(cudaFuncSetAttribute(augment_fast_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kFastBuf * (int)sizeof(float2)));
    HB_CUDA_OK(cudaFuncSetAttribute(augment_fast_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kFastBuf * (int)sizeof(float2)));
    HB_CUDA_OK(cudaFuncSetAttribute(augment_fast_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kFastBuf * (int)sizeof(float2)));
    HB_CUDA_OK(cudaFuncSetAttribute(bandstop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kFastBuf * (int)sizeof(float2)));
    HB_CUDA_OK(cudaFuncSetAttribute(bandstop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (kMaxT / 2) * (int)sizeof(float2)));
    g_plans[{dev, T}] = plan;
    *out = plan;
    return HB_OK;
}

}  // namespace hb

using namespace hb;

extern "C" int hb_rir_spectrum(const float* kernels_dev, float* spec_dev, int n, int T, void* stream) {
    HB_REQUIRE(kernels_dev && spec_dev && n >= 0, "hb_rir_spectrum: bad argument");
    if (n == 0) return HB_OK;
    FftPlan plan;
    int rc = get_plan(T, &plan);
    if (rc) return rc;
    const size_t smem = 2 * (size_t)plan.M * sizeof(float2);
    rir_spectrum_kernel<<<n, kAugThreads, smem, (cudaStream_t)stream>>>(kernels_dev, reinterpret_cast<float2*>(spec_dev), T, plan);
    HB_LAUNCHED();
    return HB_OK;
}

// Coloured-noise patterns of k batches (global batch ids + f_decay, device arrays) -> f32 [k][16000], unit RMS each.
// Replaces the host-side pattern generation of torch_audiomentations AddColoredNoise (reference augmented.py:107-115).
extern "C" int hb_colored_bases(uint64_t seed, const int64_t* batch_ids_dev, const float* f_decay_dev, int k, float* out_dev,
                                void* stream) {
    HB_REQUIRE(k >= 0 && (k == 0 || (batch_ids_dev && f_decay_dev && out_dev)), "hb_colored_bases: bad argument");
    HB_REQUIRE((reinterpret_cast<uintptr_t>(out_dev) & 7) == 0, "hb_colored_bases: output must be 8-byte aligned");
    if (k == 0) return HB_OK;
    FftPlan plan;
    int rc = get_plan(kColoredBase, &plan);
    if (rc) return rc;
    const size_t smem = 2 * (size_t)plan.M * sizeof(float2);
    colored_bases_kernel<<<k, kAugThreads, smem, (cudaStream_t)stream>>>(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), batch_ids_dev,
                                                                           f_decay_dev, out_dev, plan);
    HB_LAUNCHED();
    return HB_OK;
}

// CTAs ahead whose inputs a CTA prefetches into L2 = one wave (one CTA per SM); HB_AUG_PREFETCH overrides (0 = off)
static int prefetch_distance() {
    static int d = -1;
    if (d < 0) {
        const char* e = getenv("HB_AUG_PREFETCH");
        if (e) d = atoi(e);
        else {
            int dev = 0, n_sm = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            d = n_sm;
        }
    }
    return d;
}

extern "C" int hb_augment_clips_f32(const float* clips_dev, const float* noise_bank_dev, const float* colored_bases_dev,
                                    const float* rir_spec_bank_dev, const hb_clip_aug* params_dev, float* out_dev, int n,
                                    int T, void* stream) {
    HB_REQUIRE(clips_dev && params_dev && out_dev && n >= 0, "hb_augment_clips_f32: bad argument");
    if (n == 0) return HB_OK;
    FftPlan plan;
    int rc = get_plan(T, &plan);
    if (rc) return rc;
    const bool aligned = ((reinterpret_cast<uintptr_t>(clips_dev) | reinterpret_cast<uintptr_t>(out_dev) |
                           reinterpret_cast<uintptr_t>(colored_bases_dev)) & 7) == 0;
    if (T == kFastT && aligned) {
        augment_fast_kernel<false, false><<<n, kFastThreads, 2 * kFastBuf * sizeof(float2), (cudaStream_t)stream>>>(
            clips_dev, nullptr, nullptr, nullptr, noise_bank_dev, colored_bases_dev, reinterpret_cast<const float2*>(rir_spec_bank_dev),
            params_dev, out_dev, plan, prefetch_distance(), nullptr, 1.0f);
    } else {
        const size_t smem = 2 * (size_t)plan.M * sizeof(float2);
        augment_kernel<<<n, kAugThreads, smem, (cudaStream_t)stream>>>(clips_dev, noise_bank_dev, colored_bases_dev,
                                                                         reinterpret_cast<const float2*>(rir_spec_bank_dev),
                                                                         params_dev, out_dev, T, plan);
    }
    HB_LAUNCHED();
    return HB_OK;
}

extern "C" int hb_augment_clips_i16(const int16_t* samples_dev, const int64_t* offsets_dev, const int32_t* pad_before_dev,
                                    const float* noise_bank_dev, const float* colored_bases_dev, const float* rir_spec_bank_dev,
                                    const hb_clip_aug* params_dev, float* out_dev, int n, int T, void* stream) {
    HB_REQUIRE(samples_dev && offsets_dev && pad_before_dev && params_dev && out_dev && n >= 0, "hb_augment_clips_i16: bad argument");
    if (T != kFastT || ((reinterpret_cast<uintptr_t>(out_dev) | reinterpret_cast<uintptr_t>(colored_bases_dev)) & 7) != 0) {
        set_error("hb_augment_clips_i16: only T = 23040 with 8-byte aligned buffers is fused; use hb_fix_length_i16 + hb_augment_clips_f32");
        return HB_ERR_UNSUPPORTED;
    }
    if (n == 0) return HB_OK;
    FftPlan plan;
    int rc = get_plan(T, &plan);
    if (rc) return rc;
    augment_fast_kernel<true, false><<<n, kFastThreads, 2 * kFastBuf * sizeof(float2), (cudaStream_t)stream>>>(
        nullptr, samples_dev, offsets_dev, pad_before_dev, noise_bank_dev, colored_bases_dev,
        reinterpret_cast<const float2*>(rir_spec_bank_dev), params_dev, out_dev, plan, prefetch_distance(), nullptr, 1.0f);
    HB_LAUNCHED();
    return HB_OK;
}

// Production mode: a1 + K1-K4 + K6 in ONE kernel -- ragged int16 clips -> augmented clip (shared memory only) -> log-mel
// mel_dev f32 [n][141][32] (T = 23040).  Bit-identical to hb_augment_clips_i16 followed by hb_mel_f32(scale).
extern "C" int hb_augment_mel_i16(const int16_t* samples_dev, const int64_t* offsets_dev, const int32_t* pad_before_dev,
                                  const float* noise_bank_dev, const float* colored_bases_dev, const float* rir_spec_bank_dev,
                                  const hb_clip_aug* params_dev, float scale, float* mel_dev, int n, int T, void* stream) {
    HB_REQUIRE(samples_dev && offsets_dev && pad_before_dev && params_dev && mel_dev && n >= 0, "hb_augment_mel_i16: bad argument");
    if (T != kFastT || (reinterpret_cast<uintptr_t>(colored_bases_dev) & 7) != 0) {
        set_error("hb_augment_mel_i16: only T = 23040 with 8-byte aligned buffers is fused; use hb_augment_clips_* + hb_mel_f32");
        return HB_ERR_UNSUPPORTED;
    }
    HB_REQUIRE(mel_tables_ready(), "hb_augment_mel_i16: hb_init_tables has not been called on this device");
    if (n == 0) return HB_OK;
    FftPlan plan;
    int rc = get_plan(T, &plan);
    if (rc) return rc;
    const MelTables* tables = mel_tables_device();
    HB_REQUIRE(tables != nullptr, "hb_augment_mel_i16: mel tables unavailable");
    augment_fast_kernel<true, true><<<n, kFastThreads, 2 * kFastBuf * sizeof(float2), (cudaStream_t)stream>>>(
        nullptr, samples_dev, offsets_dev, pad_before_dev, noise_bank_dev, colored_bases_dev,
        reinterpret_cast<const float2*>(rir_spec_bank_dev), params_dev, mel_dev, plan, prefetch_distance(), tables, scale);
    HB_LAUNCHED();
    return HB_OK;
}

// K9 BandStopFilter, in place on the k clips listed in clip_index (rows of clips_dev f32 [n][T]): meta i32 [k][3] = (first
// spectrum row, partitions, half size h) per clip, spec_dev = hb_rir_spectrum of the partition rows, scratch_dev f32 [k][T].
extern "C" int hb_k9_bandstop_f32(float* clips_dev, const int32_t* clip_index_dev, const int32_t* meta_dev, const float* spec_dev,
                                  float* scratch_dev, int k, int T, void* stream) {
    HB_REQUIRE(k >= 0 && T > 0 && T % 4 == 0 && (k == 0 || (clips_dev && clip_index_dev && meta_dev && spec_dev && scratch_dev)),
               "hb_k9_bandstop_f32: bad argument (T must be a multiple of 4)");
    HB_REQUIRE(((reinterpret_cast<uintptr_t>(clips_dev) | reinterpret_cast<uintptr_t>(spec_dev)) & 7) == 0, "hb_k9_bandstop_f32: buffers must be 8-byte aligned");
    if (k == 0) return HB_OK;
    FftPlan plan;
    int rc = get_plan(T, &plan);
    if (rc) return rc;
    gather_rows_kernel<<<k, 256, 0, (cudaStream_t)stream>>>(clips_dev, clip_index_dev, scratch_dev, T);
    HB_LAUNCHED();
    if (T == kFastT)
        bandstop_kernel<true><<<k, kFastThreads, 2 * kFastBuf * sizeof(float2), (cudaStream_t)stream>>>(
            clips_dev, clip_index_dev, meta_dev, reinterpret_cast<const float2*>(spec_dev), scratch_dev, T, plan);
    else
        bandstop_kernel<false><<<k, kAugThreads, 2 * (size_t)plan.M * sizeof(float2), (cudaStream_t)stream>>>(
            clips_dev, clip_index_dev, meta_dev, reinterpret_cast<const float2*>(spec_dev), scratch_dev, T, plan);
    HB_LAUNCHED();
    return HB_OK;
}

extern "C" int hb_debug_aug_times(long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, g_aug_times, sizeof(long long) * 8 * 8) == cudaSuccess ? 0 : -2;
}

extern "C" int hb_fix_length_i16(const int16_t* samples_dev, const int64_t* offsets_dev, const int32_t* pad_before_dev,
                                 float* out_dev, int n, int T, void* stream) {
    HB_REQUIRE(samples_dev && offsets_dev && pad_before_dev && out_dev && n >= 0 && T > 0, "hb_fix_length_i16: bad argument");
    if (n == 0) return HB_OK;
    fix_length_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(samples_dev, offsets_dev, pad_before_dev, out_dev, T);
    HB_LAUNCHED();
    return HB_OK;
}
