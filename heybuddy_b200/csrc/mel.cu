// K6: fused STFT-power -> 32-bin mel (dense fp32 projection) -> log10 + 2.
//
// Replaces the ORT run of mel-spectrogram.onnx behind MelSpectrogramModel.__call__
// (reference src/python/heybuddy/spectrogram.py:23-32).  Arithmetic spec: SURVEY.md A.4 /
// heybuddy_b200/spec.py.
//
// Layout: one CTA per (clip, 16-frame chunk).  The chunk's 2912 samples are staged once in
// shared memory with coalesced 128-bit loads (each sample is reused by 3.2 frames).  Each of
// the 8 warps owns 2 frames: 512-point real FFT computed as a 256-point complex Stockham
// radix-4 FFT in the warp's private shared-memory ping-pong buffers, power of bins [2,122)
// (the only filterbank rows that are non-zero for 60..3800 Hz), then the [120 x 32] mel
// projection as a dense fp32 product (lane m owns mel bin m), log10, +2, one coalesced
// 128-byte store per frame.
//
// Roofline: HBM-bound on paper (92,160 B in + 18,048 B out per 23040-sample clip); the FFT
// makes it FP32/shared-memory-bound in practice.
#include "hb_common.cuh"

namespace hb {

struct MelTables {
    float window[kWinLength];        // Hann(400); the 56-sample zero pads are implicit
    float2 w256[256];                // exp(-2 pi i k / 256)
    float2 w512[129];                // exp(-2 pi i k / 512), k = 0..128
    float fb[kMelBand * kMels];      // filterbank rows [2,122) x 32
};

__device__ MelTables g_mel_tables;
static bool g_tables_ready[64] = {false};

constexpr int kFramesPerCta = 16;
constexpr int kMelThreads = 256;
constexpr int kMelWarps = kMelThreads / 32;
constexpr int kFramesPerWarp = kFramesPerCta / kMelWarps;  // 2
constexpr int kChunkSamples = (kFramesPerCta - 1) * kHop + kNFFT;  // 2912

struct MelSmem {
    float fb[kMelBand * kMels];          // 15360 B
    float window[kWinLength];            // 1600 B
    float2 w256[256];                    // 2048 B
    float2 w512[132];                    // 1056 B (129 used)
    float samples[kChunkSamples];        // 11648 B
    float2 fft[kMelWarps][2][256];       // 32768 B
    float power[kMelWarps][128];         // 4096 B
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// One Stockham radix-4 pass over 256 complex points; lane handles butterflies j = lane, lane + 32.
template <int NS>
__device__ __forceinline__ void stockham_r4(const float2* __restrict__ in, float2* __restrict__ out,
                                            const float2* __restrict__ w256, int lane) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        const int k = j & (NS - 1);
        float2 v0 = in[j], v1 = in[j + 64], v2 = in[j + 128], v3 = in[j + 192];
        if (NS > 1) {
            constexpr int step = 64 / NS;
            v1 = cmul(v1, w256[k * step]);
            v2 = cmul(v2, w256[2 * k * step]);
            v3 = cmul(v3, w256[3 * k * step]);
        }
        const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
        const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
        const float2 a3 = make_float2(v1.x - v3.x, v1.y - v3.y);
        const int d = ((j - k) << 2) + k;  // (j / NS) * NS * 4 + k
        out[d] = make_float2(a0.x + a2.x, a0.y + a2.y);
        out[d + NS] = make_float2(a1.x + a3.y, a1.y - a3.x);      // a1 - i a3
        out[d + 2 * NS] = make_float2(a0.x - a2.x, a0.y - a2.y);
        out[d + 3 * NS] = make_float2(a1.x - a3.y, a1.y + a3.x);  // a1 + i a3
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kMelThreads)
mel_kernel(const float* __restrict__ audio, int64_t row_stride, float scale, float* __restrict__ mel,
           int T, int F, int chunks_per_clip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MelSmem& s = *reinterpret_cast<MelSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int clip = blockIdx.x / chunks_per_clip;
    const int frame0 = (blockIdx.x - clip * chunks_per_clip) * kFramesPerCta;
    const int sample0 = frame0 * kHop;

    // ---- stage tables + samples -----------------------------------------------------------
    for (int i = tid; i < kMelBand * kMels; i += kMelThreads) s.fb[i] = g_mel_tables.fb[i];
    for (int i = tid; i < kWinLength; i += kMelThreads) s.window[i] = g_mel_tables.window[i];
    for (int i = tid; i < 256; i += kMelThreads) s.w256[i] = g_mel_tables.w256[i];
    for (int i = tid; i < 129; i += kMelThreads) s.w512[i] = g_mel_tables.w512[i];
    {
        const float* src = audio + (int64_t)clip * row_stride + sample0;
        const int avail = min(kChunkSamples, T - sample0);
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        if (vec_ok) {
            const int nvec = avail >> 2;
            const float4* src4 = reinterpret_cast<const float4*>(src);
            for (int i = tid; i < nvec; i += kMelThreads) {
                float4 v = __ldg(src4 + i);
                s.samples[4 * i + 0] = v.x * scale;
                s.samples[4 * i + 1] = v.y * scale;
                s.samples[4 * i + 2] = v.z * scale;
                s.samples[4 * i + 3] = v.w * scale;
            }
            for (int i = (nvec << 2) + tid; i < avail; i += kMelThreads) s.samples[i] = __ldg(src + i) * scale;
        } else {
            for (int i = tid; i < avail; i += kMelThreads) s.samples[i] = __ldg(src + i) * scale;
        }
    }
    __syncthreads();

    float2* buf0 = s.fft[warp][0];
    float2* buf1 = s.fft[warp][1];
    float* pw = s.power[warp];

    for (int fi = 0; fi < kFramesPerWarp; ++fi) {
        const int fl = warp * kFramesPerWarp + fi;  // frame within the chunk
        const int f = frame0 + fl;
        if (f >= F) break;  // warp-uniform
        const float* x = s.samples + fl * kHop;

        // z[n] = (x[2n] w[2n], x[2n+1] w[2n+1]); the window is zero outside [56, 456)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = lane + 32 * i;
            const int i0 = 2 * n - kWinPad;
            float2 z = make_float2(0.f, 0.f);
            if (i0 >= 0 && i0 < kWinLength) {  // 56 and 400 are even, so both samples share the test
                const float2 xv = *reinterpret_cast<const float2*>(x + 2 * n);
                z.x = xv.x * s.window[i0];
                z.y = xv.y * s.window[i0 + 1];
            }
            buf0[n] = z;
        }
        __syncwarp();
        stockham_r4<1>(buf0, buf1, s.w256, lane);
        stockham_r4<4>(buf1, buf0, s.w256, lane);
        stockham_r4<16>(buf0, buf1, s.w256, lane);
        stockham_r4<64>(buf1, buf0, s.w256, lane);

        // real-FFT post-processing + power for bins [2, 122) -> pw[0..120)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = lane + 32 * i;
            if (k >= kMelBinLo && k < kMelBinHi) {
                const float2 a = buf0[k];
                const float2 b = buf0[256 - k];
                const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));   // (A + conj B)/2
                const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));   // (A - conj B)/2
                const float2 o = make_float2(d.y, -d.x);                                 // -i d
                const float2 wo = cmul(s.w512[k], o);
                const float re = e.x + wo.x, im = e.y + wo.y;
                pw[k - kMelBinLo] = re * re + im * im;
            }
        }
        __syncwarp();

        // dense [120 x 32] projection: lane = mel bin
        float acc = 0.f;
#pragma unroll 8
        for (int k = 0; k < kMelBand; ++k) acc = fmaf(pw[k], s.fb[k * kMels + lane], acc);
        mel[((int64_t)clip * F + f) * kMels + lane] = log10f(acc < 1e-10f ? 1e-10f : acc) + 2.0f;  // NaN-propagating clamp, like np.maximum / torch.clamp
        __syncwarp();
    }
}

}  // namespace hb

using namespace hb;

extern "C" int hb_mel_frames(int T) { return T < kNFFT ? 0 : 1 + (T - kNFFT) / kHop; }

extern "C" int hb_init_tables(const float* hann_host, const float* melfb_host) {
    HB_REQUIRE(hann_host && melfb_host, "hb_init_tables: null table");
    static MelTables t;  // host staging
    for (int i = 0; i < kWinLength; ++i) t.window[i] = hann_host[kWinPad + i];
    for (int i = 0; i < kNFFT; ++i)
        if (i < kWinPad || i >= kWinPad + kWinLength)
            HB_REQUIRE(hann_host[i] == 0.f, "hb_init_tables: window must be zero outside [56,456)");
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < 256; ++k) {
        t.w256[k].x = (float)cos(two_pi * k / 256.0);
        t.w256[k].y = (float)-sin(two_pi * k / 256.0);
    }
    for (int k = 0; k <= 128; ++k) {
        t.w512[k].x = (float)cos(two_pi * k / 512.0);
        t.w512[k].y = (float)-sin(two_pi * k / 512.0);
    }
    for (int k = 0; k < 257; ++k)
        for (int m = 0; m < kMels; ++m) {
            const float v = melfb_host[k * kMels + m];
            if (k >= kMelBinLo && k < kMelBinHi) {
                t.fb[(k - kMelBinLo) * kMels + m] = v;
            } else {
                HB_REQUIRE(v == 0.f, "hb_init_tables: filterbank row %d outside the compiled band [%d,%d) is non-zero",
                           k, kMelBinLo, kMelBinHi);
            }
        }
    HB_CUDA_OK(cudaMemcpyToSymbol(g_mel_tables, &t, sizeof(MelTables)));
    HB_CUDA_OK(cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MelSmem)));
    int dev = 0;
    HB_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64) g_tables_ready[dev] = true;
    return HB_OK;
}

extern "C" int hb_mel_f32(const float* audio_dev, int64_t audio_row_stride, float scale, float* mel_dev,
                          int B, int T, void* stream) {
    HB_REQUIRE(audio_dev && mel_dev, "hb_mel_f32: null pointer");
    HB_REQUIRE(B >= 0 && T >= 0 && audio_row_stride >= T, "hb_mel_f32: bad shape B=%d T=%d stride=%lld", B, T,
               (long long)audio_row_stride);
    int dev = 0;
    HB_CUDA_OK(cudaGetDevice(&dev));
    HB_REQUIRE(dev < 64 && g_tables_ready[dev], "hb_mel_f32: hb_init_tables has not been called on device %d", dev);
    const int F = hb_mel_frames(T);
    if (B == 0 || F == 0) return HB_OK;
    const int chunks = ceil_div(F, kFramesPerCta);
    HB_REQUIRE((int64_t)B * chunks < (1ll << 31), "hb_mel_f32: B=%d too large for one launch", B);
    mel_kernel<<<B * chunks, kMelThreads, sizeof(MelSmem), (cudaStream_t)stream>>>(audio_dev, audio_row_stride, scale,
                                                                                     mel_dev, T, F, chunks);
    HB_LAUNCHED();
    return HB_OK;
}
