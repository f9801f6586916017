// K6: fused STFT-power -> 32-bin mel (banded fp32 projection) -> log10 + 2.
//
// Replaces the ORT run of mel-spectrogram.onnx behind MelSpectrogramModel.__call__
// (reference src/python/heybuddy/spectrogram.py:23-32).  Arithmetic spec: SURVEY.md A.4 /
// heybuddy_b200/spec.py.
//
// Layout: a warp owns TWO frames at a time (half-warp h = lane / 16 -> frame 2p + h) and keeps the
// whole 512-point real FFT in registers:
//   * the frame is packed into 256 complex points z[n] = (x[2n] w[2n], x[2n+1] w[2n+1]); lane l
//     loads the 16 points n = l + 16 i straight from global memory (a half-warp reads 128
//     contiguous bytes per i; every sample is re-read by 3.2 frames out of L1),
//   * 256 = 16 x 16: a 16-point DFT in registers (two radix-4 levels), the W256^(l k1) twiddles
//     (16 register constants per lane), ONE 16 x 16 transpose through padded shared memory, a second
//     16-point DFT -> lane l holds X[l + 16 k2],
//   * the real-FFT untangle pairs X[k] with X[256 - k], which lives in lane 16 - l: one shuffle,
//   * power of bins [2, 122) goes to shared memory; lane m = mel bin m sums its own <= 16-tap band of
//     the triangular filterbank (each FFT bin feeds at most two mel bins, so the dense [120 x 32]
//     product is 88 % zeros), log10, + 2, one coalesced 128-byte store per frame.
// The previous version (radix-4 Stockham in shared memory + dense projection) was shared-memory
// wavefront bound: l1tex 94 % busy, 536 wavefronts per frame (profiles/README.md); this one needs ~100.
//
// Roofline: HBM-bound on paper (92,160 B in + 18,048 B out per 23040-sample clip).
#include "hb_common.cuh"

namespace hb {

constexpr int kMelTaps = 16;         // widest supported filterbank band (the 60..3800 Hz HTK bank needs 15)

struct MelTables {
    float2 win2[256];                // (w[2n], w[2n+1]) of the 512-sample padded Hann window
    float2 w256[256];                // exp(-2 pi i k / 256)
    float2 w512[128];                // exp(-2 pi i k / 512), k = 0..127
    float fbw[kMelTaps * kMels];     // banded filterbank, taps rotated per mel bin: fbw[j][m] = fb[lo[m] + (j + rot[m]) % 16][m]
    int lo[kMels];                   // first FFT bin of mel bin m
    int rot[kMels];                  // tap rotation of mel bin m: lane m reads power[lo[m] + (j + rot[m]) % 16] at step j
};

__device__ MelTables g_mel_tables;
static bool g_tables_ready[64] = {false};

constexpr int kMelThreads = 256;
constexpr int kMelWarps = kMelThreads / 32;
constexpr int kPowerRow = 128 + 16;
constexpr int kTrStride = 17;        // float2 row stride of the transpose tile: conflict-free both ways

struct MelSmem {
    float2 win2[kWinLength / 2];     // the non-zero part of the window: points [kWinPad / 2, (kWinPad + kWinLength) / 2)
    float2 w512[128];
    float fbw[kMelTaps * kMels];
    int lo[kMels];
    int rot[kMels];
    float2 tr[kMelWarps][2][16 * kTrStride];
    float power[kMelWarps][2][kPowerRow];  // bins [0,128) (+16: the two frames of a pair start 16 banks apart); lo + kMelTaps <= 128
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
    const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
    const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
    const float2 a3 = make_float2(v1.x - v3.x, v1.y - v3.y);
    v0 = make_float2(a0.x + a2.x, a0.y + a2.y);
    v1 = make_float2(a1.x + a3.y, a1.y - a3.x);      // a1 - i a3
    v2 = make_float2(a0.x - a2.x, a0.y - a2.y);
    v3 = make_float2(a1.x - a3.y, a1.y + a3.x);      // a1 + i a3
}

// Forward 16-point DFT in registers.  Input v[n]; output X[k] is left at v[rev16(k)], rev16(k) = 4 (k % 4) + k / 4.
__host__ __device__ constexpr int rev16(int k) { return 4 * (k & 3) + (k >> 2); }

__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
#pragma unroll
    for (int a = 0; a < 4; ++a) dft4(v[a], v[a + 4], v[a + 8], v[a + 12]);   // v[a + 4 kb] = Y[a][kb]
    // Y[a][kb] *= W16^(a kb)
    v[5] = cmul(v[5], make_float2(c1, -s1));                 // W^1
    v[6] = make_float2(h * (v[6].x + v[6].y), h * (v[6].y - v[6].x));       // W^2 = (h, -h)
    v[7] = cmul(v[7], make_float2(s1, -c1));                 // W^3
    v[9] = make_float2(h * (v[9].x + v[9].y), h * (v[9].y - v[9].x));       // W^2
    v[10] = make_float2(v[10].y, -v[10].x);                  // W^4 = -i
    v[11] = make_float2(h * (v[11].y - v[11].x), -h * (v[11].x + v[11].y)); // W^6 = (-h, -h)
    v[13] = cmul(v[13], make_float2(s1, -c1));               // W^3
    v[14] = make_float2(h * (v[14].y - v[14].x), -h * (v[14].x + v[14].y)); // W^6
    v[15] = cmul(v[15], make_float2(-c1, s1));               // W^9 = -W^1
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) dft4(v[4 * kb], v[4 * kb + 1], v[4 * kb + 2], v[4 * kb + 3]);  // v[4 kb + ka] = X[kb + 4 ka]
}

__global__ void __launch_bounds__(kMelThreads, 2)
mel_kernel(const float* __restrict__ audio, int64_t row_stride, float scale, float* __restrict__ mel,
           int B, int F, int pairs_per_clip) {
    __shared__ MelSmem s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int h = lane >> 4, l = lane & 15;
    for (int i = tid; i < kWinLength / 2; i += kMelThreads) s.win2[i] = g_mel_tables.win2[i + kWinPad / 2];
    for (int i = tid; i < 128; i += kMelThreads) s.w512[i] = g_mel_tables.w512[i];
    for (int i = tid; i < kMelTaps * kMels; i += kMelThreads) s.fbw[i] = g_mel_tables.fbw[i];
    if (tid < kMels) s.lo[tid] = g_mel_tables.lo[tid];
    if (tid < kMels) s.rot[tid] = g_mel_tables.rot[tid];
    for (int i = tid; i < kMelWarps * 2 * kPowerRow; i += kMelThreads) (&s.power[0][0][0])[i] = 0.f;
    float2 tw[16];                                   // W256^(l k1)
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) tw[k1] = g_mel_tables.w256[(l * k1) & 255];
    __syncthreads();
    const int my_lo = s.lo[lane], my_rot = s.rot[lane];
    float2* tr = s.tr[warp][h];
    float* pw = s.power[warp][h];
    const bool vec_ok = ((row_stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(audio) & 7) == 0);
    const int64_t n_pairs = (int64_t)B * pairs_per_clip;
    const int64_t warp_stride = (int64_t)gridDim.x * kMelWarps;

    for (int64_t p = (int64_t)blockIdx.x * kMelWarps + warp; p < n_pairs; p += warp_stride) {
        const int clip = (int)(p / pairs_per_clip);
        const int f0 = (int)(p - (int64_t)clip * pairs_per_clip) * 2;
        const int f = min(f0 + h, F - 1);            // an odd F repeats the last frame in the idle half-warp
        const float* x = audio + (int64_t)clip * row_stride + (int64_t)f * kHop;
        {
            // pull the NEXT pair's 672 samples (21 lines of 128 B) into L1 while this pair is transformed: with 8 warps per CTA
            // there is not enough parallelism to hide the L2 latency of the sample loads otherwise
            const int64_t pn = p + warp_stride;
            if (pn < n_pairs && lane < 22) {
                const int cn = (int)(pn / pairs_per_clip);
                const float* xn = audio + (int64_t)cn * row_stride + (int64_t)((int)(pn - (int64_t)cn * pairs_per_clip) * 2) * kHop;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(xn + lane * 32));
            }
        }

        // z[i] = point n = l + 16 i; the window is zero outside samples [56, 456) = points [28, 228)
        float2 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = l + 16 * i;
            v[i] = make_float2(0.f, 0.f);
            if (i >= 1 && i <= 14 && n >= kWinPad / 2 && n < (kWinPad + kWinLength) / 2) {
                float2 xv;
                if (vec_ok) xv = __ldg(reinterpret_cast<const float2*>(x) + n);
                else xv = make_float2(__ldg(x + 2 * n), __ldg(x + 2 * n + 1));
                const float2 w = s.win2[n - kWinPad / 2];
                v[i] = make_float2(xv.x * scale * w.x, xv.y * scale * w.y);
            }
        }
        dft16(v);                                     // v[rev16(k1)] = sum_i z[l + 16 i] W16^(i k1)
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) tr[l * kTrStride + k1] = (k1 == 0) ? v[rev16(k1)] : cmul(v[rev16(k1)], tw[k1]);
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) v[n1] = tr[n1 * kTrStride + l];
        dft16(v);                                     // v[rev16(k2)] = X[l + 16 k2]

        // real-FFT untangle + power, bins k = l + 16 k2 < 128 (only [2, 122) is read back)
        const int partner = ((16 - l) & 15) + 16 * h;
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            const float2 a = v[rev16(k2)];
            float2 b;
            b.x = __shfl_sync(0xffffffffu, v[rev16(15 - k2)].x, partner);
            b.y = __shfl_sync(0xffffffffu, v[rev16(15 - k2)].y, partner);
            if (l == 0) b = v[rev16((16 - k2) & 15)];  // X[256 - 16 k2] is in this lane (k2 = 0 -> bin 0, unused)
            const int k = l + 16 * k2;
            const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));   // (A + conj B)/2
            const float2 d = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y + b.y));   // (A - conj B)/2
            const float2 o = make_float2(d.y, -d.x);                                 // -i d
            const float2 wo = cmul(s.w512[k], o);
            const float re = e.x + wo.x, im = e.y + wo.y;
            pw[k] = re * re + im * im;
        }
        __syncwarp();

        // banded projection: lane = mel bin, both frames of the pair
        float fw[kMelTaps];                           // this lane's filterbank taps: loaded once for both frames of the pair
#pragma unroll
        for (int j = 0; j < kMelTaps; ++j) fw[j] = s.fbw[j * kMels + lane];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const float* q = s.power[warp][hh] + my_lo;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < kMelTaps; ++j) acc = fmaf(q[(j + my_rot) & (kMelTaps - 1)], fw[j], acc);   // rotated: no bank conflicts
            const int fo = f0 + hh;
            if (fo < F)   // NaN-propagating clamp, like np.maximum / torch.clamp
                mel[((int64_t)clip * F + fo) * kMels + lane] = log10f(acc < 1e-10f ? 1e-10f : acc) + 2.0f;
        }
        __syncwarp();
    }
}

}  // namespace hb

using namespace hb;

extern "C" int hb_mel_frames(int T) { return T < kNFFT ? 0 : 1 + (T - kNFFT) / kHop; }

extern "C" int hb_init_tables(const float* hann_host, const float* melfb_host) {
    HB_REQUIRE(hann_host && melfb_host, "hb_init_tables: null table");
    static MelTables t;  // host staging
    for (int i = 0; i < kNFFT; ++i)
        if (i < kWinPad || i >= kWinPad + kWinLength)
            HB_REQUIRE(hann_host[i] == 0.f, "hb_init_tables: window must be zero outside [56,456)");
    for (int n = 0; n < 256; ++n) t.win2[n] = make_float2(hann_host[2 * n], hann_host[2 * n + 1]);
    const double two_pi = 6.283185307179586476925286766559;
    for (int k = 0; k < 256; ++k) {
        t.w256[k].x = (float)cos(two_pi * k / 256.0);
        t.w256[k].y = (float)-sin(two_pi * k / 256.0);
    }
    for (int k = 0; k < 128; ++k) {
        t.w512[k].x = (float)cos(two_pi * k / 512.0);
        t.w512[k].y = (float)-sin(two_pi * k / 512.0);
    }
    for (int m = 0; m < kMels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < 257; ++k) {
            if (melfb_host[k * kMels + m] == 0.f) continue;
            HB_REQUIRE(k >= kMelBinLo && k < kMelBinHi,
                       "hb_init_tables: filterbank row %d outside the compiled band [%d,%d) is non-zero", k, kMelBinLo, kMelBinHi);
            if (first < 0) first = k;
            last = k;
        }
        if (first < 0) first = last = kMelBinLo;
        HB_REQUIRE(last - first < kMelTaps, "hb_init_tables: mel bin %d spans %d FFT bins, at most %d supported", m,
                   last - first + 1, kMelTaps);
        if (first + kMelTaps > 128) first = 128 - kMelTaps;   // keep the tap window inside the power row
        t.lo[m] = first;
        t.rot[m] = 0;
        for (int j = 0; j < kMelTaps; ++j) t.fbw[j * kMels + m] = (first + j <= last) ? melfb_host[(first + j) * kMels + m] : 0.f;
    }
    // Lane m reads power[lo[m] + j] at step j of the projection: mel bins whose bands start a multiple of 32 bins apart collide
    // in a shared-memory bank (32 excess wavefronts per frame with the 60..3800 Hz bank, a third of the kernel's shared-memory
    // traffic).  The order of a lane's taps is free, so rotate it per lane -- step j reads tap (j + rot[m]) % 16 -- and search the
    // rotations (deterministic local search) for the assignment with the fewest same-bank, different-address reads.
    {
        static_assert(kMelTaps == 16, "tap rotation assumes 16 taps");
        auto cost = [&](const int* rot) {
            int c = 0;
            for (int j = 0; j < kMelTaps; ++j) {
                int worst = 1;
                for (int b = 0; b < 32; ++b) {
                    int addrs[kMels], n = 0;
                    for (int m = 0; m < kMels; ++m) {
                        const int a = t.lo[m] + ((j + rot[m]) & (kMelTaps - 1));
                        if ((a & 31) != b) continue;
                        bool seen = false;
                        for (int i = 0; i < n; ++i) seen |= (addrs[i] == a);
                        if (!seen) addrs[n++] = a;
                    }
                    if (n > worst) worst = n;
                }
                c += worst - 1;
            }
            return c;
        };
        int best[kMels] = {0};
        int best_cost = cost(best);
        uint32_t lcg = 12345u;
        for (int it = 0; it < 200000 && best_cost > 0; ++it) {
            int trial[kMels];
            for (int m = 0; m < kMels; ++m) trial[m] = best[m];
            lcg = lcg * 1664525u + 1013904223u;
            const int m = (int)((lcg >> 8) % kMels);
            lcg = lcg * 1664525u + 1013904223u;
            trial[m] = (int)((lcg >> 8) % kMelTaps);
            const int c = cost(trial);
            if (c <= best_cost) {
                best_cost = c;
                for (int i = 0; i < kMels; ++i) best[i] = trial[i];
            }
        }
        float rotated[kMelTaps * kMels];
        for (int m = 0; m < kMels; ++m) {
            t.rot[m] = best[m];
            for (int j = 0; j < kMelTaps; ++j) rotated[j * kMels + m] = t.fbw[((j + best[m]) & (kMelTaps - 1)) * kMels + m];
        }
        for (int i = 0; i < kMelTaps * kMels; ++i) t.fbw[i] = rotated[i];
    }
    HB_CUDA_OK(cudaMemcpyToSymbol(g_mel_tables, &t, sizeof(MelTables)));
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
    const int pairs = ceil_div(F, 2);
    const int64_t n_pairs = (int64_t)B * pairs;
    static int n_sm[64] = {0};
    if (n_sm[dev] == 0) HB_CUDA_OK(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
    // persistent warps: two CTAs per SM, each warp strides over (clip, frame pair) items
    const int64_t want = (n_pairs + kMelWarps - 1) / kMelWarps;
    const int grid = (int)(want < 2 * n_sm[dev] ? want : 2 * n_sm[dev]);
    mel_kernel<<<grid, kMelThreads, 0, (cudaStream_t)stream>>>(audio_dev, audio_row_stride, scale, mel_dev, B, F, pairs);
    HB_LAUNCHED();
    return HB_OK;
}
