// Shared core of the log-mel front end (K6): tables, the register-resident 512-point real FFT of a frame PAIR and the banded
// mel projection -- one definition used by the stand-alone kernel (mel.cu: samples from global memory) and by the fused
// augmentation + mel kernel (augment.cu: samples from the clip in shared memory), so the two produce bit-identical frames.
// Algorithm notes: mel.cu's header comment.
#pragma once
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

constexpr int kPowerRow = 128 + 16;
constexpr int kTrStride = 17;        // float2 row stride of the transpose tile: conflict-free both ways

// per-CTA copy of the tables the frame loop reads
struct MelShared {
    float2 win2[kWinLength / 2];     // the non-zero part of the window: points [kWinPad / 2, (kWinPad + kWinLength) / 2)
    float2 w512[128];
    float fbw[kMelTaps * kMels];
    int lo[kMels];
    int rot[kMels];
};
// per-warp scratch: one transpose tile and one power row per frame of the pair
struct MelWarpScratch {
    float2 tr[2][16 * kTrStride];
    float power[2][kPowerRow];       // bins [0,128) (+16: the two frames of a pair start 16 banks apart); lo + kMelTaps <= 128
};

__device__ __forceinline__ float2 mel_cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ void mel_dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
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

__device__ __forceinline__ void mel_dft16(float2 (&v)[16]) {
    constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
#pragma unroll
    for (int a = 0; a < 4; ++a) mel_dft4(v[a], v[a + 4], v[a + 8], v[a + 12]);   // v[a + 4 kb] = Y[a][kb]
    // Y[a][kb] *= W16^(a kb)
    v[5] = mel_cmul(v[5], make_float2(c1, -s1));                 // W^1
    v[6] = make_float2(h * (v[6].x + v[6].y), h * (v[6].y - v[6].x));       // W^2 = (h, -h)
    v[7] = mel_cmul(v[7], make_float2(s1, -c1));                 // W^3
    v[9] = make_float2(h * (v[9].x + v[9].y), h * (v[9].y - v[9].x));       // W^2
    v[10] = make_float2(v[10].y, -v[10].x);                  // W^4 = -i
    v[11] = make_float2(h * (v[11].y - v[11].x), -h * (v[11].x + v[11].y)); // W^6 = (-h, -h)
    v[13] = mel_cmul(v[13], make_float2(s1, -c1));               // W^3
    v[14] = make_float2(h * (v[14].y - v[14].x), -h * (v[14].x + v[14].y)); // W^6
    v[15] = mel_cmul(v[15], make_float2(-c1, s1));               // W^9 = -W^1
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) mel_dft4(v[4 * kb], v[4 * kb + 1], v[4 * kb + 2], v[4 * kb + 3]);  // v[4 kb + ka] = X[kb + 4 ka]
}

// tables -> shared memory (all threads of the CTA; the caller synchronises)
__device__ __forceinline__ void mel_load_shared(MelShared& s, const MelTables& t, int tid, int n_threads) {
    for (int i = tid; i < kWinLength / 2; i += n_threads) s.win2[i] = t.win2[i + kWinPad / 2];
    for (int i = tid; i < 128; i += n_threads) s.w512[i] = t.w512[i];
    for (int i = tid; i < kMelTaps * kMels; i += n_threads) s.fbw[i] = t.fbw[i];
    if (tid < kMels) s.lo[tid] = t.lo[tid];
    if (tid < kMels) s.rot[tid] = t.rot[tid];
}

// One frame pair (frames f0, f0 + 1 of a clip with F frames) by one warp: half-warp h = lane / 16 owns frame min(f0 + h, F - 1).
// load(h, n) returns the sample pair (x[2n], x[2n+1]) of that frame, n in [28, 228) (the window is zero elsewhere).
// tw(k1) = W256^(l k1) for this lane's l = lane % 16 (k1 is a compile-time constant after unrolling: the stand-alone kernel keeps
// the 16 values in registers, the fused kernel -- 80 registers per thread -- reads them from shared memory).
// Writes mel_clip[(f0 + h) * 32 + lane] for the frames < F.
// tr_pair: the warp's two 16 x 17 float2 transpose tiles; pw_pair: its two kPowerRow-float power rows -- which MAY alias the
// transpose tiles (the fused kernel does that to fit 24 warps' scratch next to the clip): the tiles are dead once every lane has
// read its column back, and a warp barrier separates that read from the first power store.
template <class Load, class Twiddle>
__device__ __forceinline__ void mel_frame_pair(const MelShared& s, float2* __restrict__ tr_pair, float* pw_pair, Twiddle tw, int my_lo, int my_rot,
                                               Load load, float scale, int f0, int F, float* __restrict__ mel_clip) {
    const int lane = threadIdx.x & 31;
    const int h = lane >> 4, l = lane & 15;
    float2* tr = tr_pair + h * (16 * kTrStride);
    float* pw = pw_pair + h * kPowerRow;
    // z[i] = point n = l + 16 i; the window is zero outside samples [56, 456) = points [28, 228)
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int n = l + 16 * i;
        v[i] = make_float2(0.f, 0.f);
        if (i >= 1 && i <= 14 && n >= kWinPad / 2 && n < (kWinPad + kWinLength) / 2) {
            const float2 xv = load(h, n);
            const float2 w = s.win2[n - kWinPad / 2];
            v[i] = make_float2(xv.x * scale * w.x, xv.y * scale * w.y);
        }
    }
    mel_dft16(v);                                     // v[rev16(k1)] = sum_i z[l + 16 i] W16^(i k1)
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) tr[l * kTrStride + k1] = (k1 == 0) ? v[rev16(k1)] : mel_cmul(v[rev16(k1)], tw(k1));
    __syncwarp();
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = tr[n1 * kTrStride + l];
    __syncwarp();                                     // the tiles may be overwritten by the power rows from here on
    mel_dft16(v);                                     // v[rev16(k2)] = X[l + 16 k2]

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
        const float2 wo = mel_cmul(s.w512[k], o);
        const float re = e.x + wo.x, im = e.y + wo.y;
        pw[k] = re * re + im * im;
    }
    __syncwarp();

    // banded projection: lane = mel bin, both frames of the pair (each tap of this lane's band is loaded once for both frames)
    const float* q0 = pw_pair + my_lo;
    const float* q1 = pw_pair + kPowerRow + my_lo;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int j = 0; j < kMelTaps; ++j) {
        const float fw = s.fbw[j * kMels + lane];
        const int at = (j + my_rot) & (kMelTaps - 1);            // rotated: no bank conflicts
        acc0 = fmaf(q0[at], fw, acc0);
        acc1 = fmaf(q1[at], fw, acc1);
    }
    // NaN-propagating clamp, like np.maximum / torch.clamp
    if (f0 < F) mel_clip[(int64_t)f0 * kMels + lane] = log10f(acc0 < 1e-10f ? 1e-10f : acc0) + 2.0f;
    if (f0 + 1 < F) mel_clip[(int64_t)(f0 + 1) * kMels + lane] = log10f(acc1 < 1e-10f ? 1e-10f : acc1) + 2.0f;
    __syncwarp();
}

// device address of the tables uploaded by hb_init_tables on the current device (mel.cu)
const MelTables* mel_tables_device();
bool mel_tables_ready();

}  // namespace hb
