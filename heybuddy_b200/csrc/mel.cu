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
#include "mel_core.cuh"

namespace hb {

__device__ MelTables g_mel_tables;
static bool g_tables_ready[64] = {false};

constexpr int kMelThreads = 256;
constexpr int kMelWarps = kMelThreads / 32;

struct MelSmem {
    MelShared t;
    MelWarpScratch w[kMelWarps];
};

__global__ void __launch_bounds__(kMelThreads, 2)
mel_kernel(const float* __restrict__ audio, int64_t row_stride, float scale, float* __restrict__ mel,
           int B, int F, int pairs_per_clip) {
    __shared__ MelSmem s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = lane & 15;
    mel_load_shared(s.t, g_mel_tables, tid, kMelThreads);
    for (int i = tid; i < kMelWarps * 2 * kPowerRow; i += kMelThreads) s.w[i / (2 * kPowerRow)].power[(i / kPowerRow) & 1][i % kPowerRow] = 0.f;
    float2 tw[16];                                   // W256^(l k1)
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) tw[k1] = g_mel_tables.w256[(l * k1) & 255];
    __syncthreads();
    const int my_lo = s.t.lo[lane], my_rot = s.t.rot[lane];
    const bool vec_ok = ((row_stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(audio) & 7) == 0);
    const int64_t n_pairs = (int64_t)B * pairs_per_clip;
    const int64_t warp_stride = (int64_t)gridDim.x * kMelWarps;

    for (int64_t p = (int64_t)blockIdx.x * kMelWarps + warp; p < n_pairs; p += warp_stride) {
        const int clip = (int)(p / pairs_per_clip);
        const int f0 = (int)(p - (int64_t)clip * pairs_per_clip) * 2;
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
        const float* clip_audio = audio + (int64_t)clip * row_stride;
        auto load = [&](int h, int n) {
            const float* x = clip_audio + (int64_t)min(f0 + h, F - 1) * kHop;       // an odd F repeats the last frame in the idle half-warp
            if (vec_ok) return __ldg(reinterpret_cast<const float2*>(x) + n);
            return make_float2(__ldg(x + 2 * n), __ldg(x + 2 * n + 1));
        };
        mel_frame_pair(s.t, &s.w[warp].tr[0][0], &s.w[warp].power[0][0], [&](int k1) { return tw[k1]; }, my_lo, my_rot, load, scale, f0, F,
                       mel + (int64_t)clip * F * kMels);
    }
}

const MelTables* mel_tables_device() {
    void* p = nullptr;
    return cudaGetSymbolAddress(&p, g_mel_tables) == cudaSuccess ? reinterpret_cast<const MelTables*>(p) : nullptr;
}
bool mel_tables_ready() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && g_tables_ready[dev];
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
