// K9: torch_audiomentations.PitchShift on the device (reference src/python/heybuddy/dataset/augmented.py:94-101, mode per_batch).
// torch_audiomentations / torch_pitch_shift are absent offline: the chain is restated from the library's published code (parity
// unpinned for the glue; torch.stft / torchaudio's phase vocoder / torch.istft / torchaudio's sinc resampler are present here and
// pin the four stages, tests/test_k9.py).  heybuddy_b200/dataset/k9.py holds the arithmetic spec.
//
//   torch_pitch_shift.pitch_shift(x, shift, sr):  n_fft = sr // 64 = 250, hop = n_fft // 32 = 7, rectangular window
//     S  = torch.stft(x, 250, 7)                       centre = True, reflect padding: F = 1 + T // 7 frames x 126 bins
//     S' = phase_vocoder(S, rate = 1 / shift)          F' = ceil(F / rate) frames
//     y  = torch.istft(S', 250, 7)                     7 (F' - 1) samples
//     z  = resample(y, sr -> int(sr / shift))          windowed-sinc polyphase filter
//     out = z[:T], zero-padded when shorter
//
// Three kernels per group of clips:
//   ps_stft_polar_kernel    one warp per frame: the 250-point real FFT as a 125-point complex FFT (three radix-5 Stockham passes
//                           through 2 KB of shared memory) + untangle; stores (|S|, angle S) -- all the vocoder reads
//   ps_vocoder_kernel       one thread per (clip, bin, chunk of frames), two passes: magnitude interpolation and the phase
//                           accumulation with the reference's rounding points (float32 increments in torchaudio's operation order,
//                           float64 running sum rounded to float32 per frame = torch.cumsum on the CPU; the phase of the top bins
//                           reaches 7e4 rad, where float32 resolves 8e-3 rad, so these rounding points are part of the result)
//   ps_istft_resample_kernel one CTA per (quarter of the clip's outputs, clip): 16 frames per round (inverse real FFT per warp),
//                           overlap-add into a shared-memory buffer in frame order (deterministic), envelope division, then the
//                           polyphase resampler straight out of shared memory into the clip's row.
#include "cplx.cuh"
#include "hb_common.cuh"

#include <math.h>

#include <vector>

struct hb_pitch_plan {
    int T, frames_in, frames_out, istft_len, orig, up, width, taps, out_len, dev;
    int32_t* idx0;        // [frames_out]  floor(time step)
    int32_t* idx1;        // [frames_out]  floor(time step + 1) (float32 arithmetic: not always idx0 + 1)
    float* alpha;         // [frames_out]  time step mod 1
    float* phase_adv;     // [126]
    float* kernel_t;      // [taps][up]    resampling kernel, transposed
    float2* w125;         // exp(-2 pi i m / 125)
    float2* w250;         // exp(-2 pi i k / 250), k <= 125
};

namespace hb {

constexpr int kPsNfft = 250, kPsHop = 7, kPsM = 125, kPsBins = 126;
constexpr int kPsStftWarps = 8;
constexpr int kPsIstftWarps = 16;
constexpr int kPsGroup = 512;          // clips per pass over the workspace

__device__ __forceinline__ int reflect_index(int j, int T) {
    if (j < 0) j = -j;
    if (j >= T) j = 2 * (T - 1) - j;
    return j;
}

// 125-point forward FFT of one warp's buffer pair (lanes 0..24 work): three radix-5 Stockham passes; returns the result buffer.
__device__ __forceinline__ float2* fft125(float2* a, float2* b, const float2* __restrict__ w125, int lane) {
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const int Ns = pass == 0 ? 1 : (pass == 1 ? 5 : 25);
        if (lane < 25) {
            const int j = lane, k = j % Ns;
            float2 v[5];
#pragma unroll
            for (int r = 0; r < 5; ++r) v[r] = a[j + 25 * r];
            if (k != 0) {
                const float2 w1 = w125[k * (25 / Ns)];
                float2 w = w1;
                v[1] = cmulf(v[1], w);
#pragma unroll
                for (int r = 2; r < 5; ++r) {
                    w = cmulf(w, w1);
                    v[r] = cmulf(v[r], w);
                }
            }
            dft<5>(v);
            const int base = (j - k) * 5 + k;
#pragma unroll
            for (int r = 0; r < 5; ++r) b[base + r * Ns] = v[r];
        }
        __syncwarp();
        float2* t = a; a = b; b = t;
    }
    return a;
}

__global__ void __launch_bounds__(kPsStftWarps * 32)
ps_stft_polar_kernel(const float* __restrict__ clips, const int32_t* __restrict__ clip_index, float2* __restrict__ S, int T, int F,
                     const float2* __restrict__ w125_g, const float2* __restrict__ w250_g) {
    __shared__ float2 buf[kPsStftWarps][2][128];
    __shared__ float2 w125[kPsM], w250[kPsBins];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kPsM; i += blockDim.x) w125[i] = w125_g[i];
    for (int i = tid; i < kPsBins; i += blockDim.x) w250[i] = w250_g[i];
    __syncthreads();
    const int t = blockIdx.x * kPsStftWarps + warp;
    if (t >= F) return;
    const float* x = clips + (int64_t)clip_index[blockIdx.y] * T;
    float2* a = buf[warp][0];
    float2* b = buf[warp][1];
    bool any = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = lane + 32 * i;
        if (m < kPsM) {
            const int j0 = kPsHop * t + 2 * m - kPsM;          // centre = True: the frame starts n_fft / 2 before sample hop * t
            const float2 v = make_float2(__ldg(x + reflect_index(j0, T)), __ldg(x + reflect_index(j0 + 1, T)));
            a[m] = v;
            any |= v.x != 0.f || v.y != 0.f;
        }
    }
    float2* out = S + ((int64_t)blockIdx.y * F + t) * kPsBins;
    if (!__any_sync(0xffffffffu, any)) {
        // a frame of exact zeros (the zero padding of a length-fixed clip): its spectrum is exactly zero, |S| = 0 and angle = atan2(0, 0) = 0
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (lane + 32 * i < kPsBins) out[lane + 32 * i] = make_float2(0.f, 0.f);
        return;
    }
    __syncwarp();
    const float2* Z = fft125(a, b, w125, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = lane + 32 * i;
        if (k < kPsBins) {
            const float2 zk = Z[k % kPsM], zm = cconj(Z[(kPsM - k) % kPsM]);
            const float2 e = cscale(cadd(zk, zm), 0.5f);
            const float2 o = cmulf(cscale(mul_neg_i(csub(zk, zm)), 0.5f), w250[k]);      // -i / 2 (Z[k] - conj Z[M-k]) W^k
            const float2 X = cadd(e, o);
            out[k] = make_float2(sqrtf(fmaf(X.x, X.x, X.y * X.y)), atan2f(X.y, X.x));
        }
    }
}

// torchaudio.functional.phase_vocoder with the spectrogram in polar form.  S [k][F][126] (norm, angle) -> Y [k][F'][126] complex.
// The only serial part is the running phase sum; the frames are cut into kPsChunks chunks per (clip, bin): pass 0 sums each chunk's
// increments (float64), pass 1 starts every chunk from the sum of the chunks before it and writes the frames.  (A single thread per
// bin over all 3372 frames is a chain of dependent L2 latencies: 2.5 ms per 512 clips; float64 partial sums re-associate the
// running sum at 1e-16 relative, far below the float32 rounding of each output.)
constexpr int kPsChunks = 16;

struct VocoderStep {
    float mag, inc;      // magnitude of output frame j; the increment frame j contributes to the phase of frame j + 1
};

__device__ __forceinline__ VocoderStep vocoder_step(const float2* __restrict__ s, int F, int i0, int i1, float al, float pa) {
    const float two_pi = 6.283185307179586f;          // float32(2 * math.pi): the scalar joins a float32 tensor operation
    const float2 p0 = i0 < F ? __ldg(s + (int64_t)i0 * kPsBins) : make_float2(0.f, 0.f);        // F.pad(spec, [0, 2]): zero frames
    const float2 p1 = i1 < F ? __ldg(s + (int64_t)i1 * kPsBins) : make_float2(0.f, 0.f);
    VocoderStep r;
    r.mag = __fadd_rn(__fmul_rn(al, p1.x), __fmul_rn(__fsub_rn(1.0f, al), p0.x));
    float ph = __fsub_rn(__fsub_rn(p1.y, p0.y), pa);                   // torchaudio's operation order, no contraction
    ph = __fsub_rn(ph, __fmul_rn(two_pi, rintf(__fdiv_rn(ph, two_pi))));
    r.inc = __fadd_rn(ph, pa);
    return r;
}

// pass 0: part[clip][chunk][bin] = sum of phase[j] over the chunk's frames, phase[0] = angle of frame 0, phase[j] = increment of j - 1
// pass 1: frames of the chunk
template <int kPass>
__global__ void __launch_bounds__(128)
ps_vocoder_kernel(const float2* __restrict__ S, float2* __restrict__ Y, double* __restrict__ part, int F, int Fo,
                  const int32_t* __restrict__ idx0, const int32_t* __restrict__ idx1, const float* __restrict__ alpha,
                  const float* __restrict__ phase_adv) {
    const int k = threadIdx.x, chunk = blockIdx.y;
    if (k >= kPsBins) return;
    const int per = (Fo + kPsChunks - 1) / kPsChunks, j_lo = chunk * per, j_hi = min(Fo, j_lo + per);
    const float2* s = S + (int64_t)blockIdx.x * F * kPsBins + k;
    double* my_part = part + ((int64_t)blockIdx.x * kPsChunks) * kPsBins + k;
    const float pa = phase_adv[k];
    // phase[j_lo]: the angle of frame 0 for the first chunk, else the increment of frame j_lo - 1
    float phase = j_lo == 0 ? __ldg(s).y : (j_lo < Fo ? vocoder_step(s, F, idx0[j_lo - 1], idx1[j_lo - 1], alpha[j_lo - 1], pa).inc : 0.f);
    double acc = 0.0;
    if (kPass == 1)
        for (int c = 0; c < chunk; ++c) acc += my_part[(int64_t)c * kPsBins];
    float2* y = Y + (int64_t)blockIdx.x * Fo * kPsBins + k;
#pragma unroll 4
    for (int j = j_lo; j < j_hi; ++j) {
        const VocoderStep st = vocoder_step(s, F, idx0[j], idx1[j], alpha[j], pa);
        acc += (double)phase;                         // torch.cumsum on the CPU: float64 running sum, float32 outputs
        if (kPass == 1) {
            const float pacc = (float)acc;
            float sn, cs;
            sincosf(pacc, &sn, &cs);
            y[(int64_t)j * kPsBins] = make_float2(st.mag * cs, st.mag * sn);
        }
        phase = st.inc;
    }
    if (kPass == 0) my_part[(int64_t)chunk * kPsBins] = acc;
}

// One CTA per (segment, clip): the clip's T output samples are cut into kPsSegments ranges; a CTA overlap-adds only the frames that
// reach the stretch of y its range resamples from (3 % of the frames are computed twice), so its buffer is a quarter of the clip and
// several CTAs share an SM.
constexpr int kPsSegments = 4;

__global__ void __launch_bounds__(kPsIstftWarps * 32)
ps_istft_resample_kernel(const float2* __restrict__ Y, float* __restrict__ clips, const int32_t* __restrict__ clip_index, int T, int Fo,
                         int istft_len, int orig, int up, int width, int taps, int out_len, const float* __restrict__ kernel_t,
                         const float2* __restrict__ w125_g, const float2* __restrict__ w250_g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float2 buf[kPsIstftWarps][2][128];
    __shared__ float2 w125[kPsM], w250[kPsBins];
    __shared__ bool frame_live[kPsIstftWarps];
    float* ola = reinterpret_cast<float*>(smem_raw);                 // padded coordinates [s_lo, s_hi)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // this CTA's outputs [o_lo, o_hi) read y[(o / up) orig - width + j], j < taps
    const int seg_len = (T + kPsSegments - 1) / kPsSegments;
    const int o_lo = blockIdx.x * seg_len, o_hi = min(T, o_lo + seg_len), o_last = min(o_hi, out_len) - 1;
    float* dst = clips + (int64_t)clip_index[blockIdx.y] * T;
    if (o_last < o_lo) {                                             // the whole range lies past the resampled signal: zeros
        for (int o = o_lo + tid; o < o_hi; o += blockDim.x) dst[o] = 0.f;
        return;
    }
    const int y_lo = max(0, (o_lo / up) * orig - width), y_hi = min(istft_len, (o_last / up) * orig - width + taps);
    const int s_lo = y_lo + kPsNfft / 2, s_hi = y_hi + kPsNfft / 2;  // istft trims n_fft / 2 from the front
    const int j_first = max(0, (s_lo - (kPsNfft - 1) + kPsHop - 1) / kPsHop), j_last = min(Fo - 1, (s_hi - 1) / kPsHop);
    for (int i = tid; i < kPsM; i += blockDim.x) w125[i] = w125_g[i];
    for (int i = tid; i < kPsBins; i += blockDim.x) w250[i] = w250_g[i];
    for (int i = tid; i < s_hi - s_lo; i += blockDim.x) ola[i] = 0.f;
    __syncthreads();
    const float2* y = Y + (int64_t)blockIdx.y * Fo * kPsBins;
    for (int j0 = j_first; j0 <= j_last; j0 += kPsIstftWarps) {
        const int j = j0 + warp;
        float2* a = buf[warp][0];
        float2* b = buf[warp][1];
        bool live = false;
        if (j <= j_last) {
            // a frame whose 126 bins are all exactly zero (both source frames were zero padding) adds nothing: skip its transform
            const float2* Yz = y + (int64_t)j * kPsBins;
            bool nz = false;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (lane + 32 * i < kPsBins) {
                    const float2 v = __ldg(Yz + lane + 32 * i);
                    nz |= v.x != 0.f || v.y != 0.f;
                }
            live = __any_sync(0xffffffffu, nz);
        }
        if (lane == 0) frame_live[warp] = live;
        if (live) {
            // irfft: Z[k] = E[k] + i O[k], E = (X[k] + conj X[M-k]) / 2, O = conj(W^k) (X[k] - conj X[M-k]) / 2; the imaginary parts
            // of the DC and Nyquist bins are ignored (C2R); z = conj(FFT(conj Z)) / M, frame[2 m] = Re z[m], frame[2 m + 1] = Im z[m]
            const float2* Yj = y + (int64_t)j * kPsBins;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = lane + 32 * i;
                if (k < kPsM) {
                    float2 xk = __ldg(Yj + k), xm = __ldg(Yj + (kPsM - k));
                    if (k == 0) { xk.y = 0.f; xm.y = 0.f; }
                    xm = cconj(xm);
                    const float2 e = cscale(cadd(xk, xm), 0.5f);
                    const float2 o = cmulf(cscale(csub(xk, xm), 0.5f), cconj(w250[k]));
                    a[k] = cconj(cadd(e, mul_pos_i(o)));
                }
            }
            __syncwarp();
            float2* z = fft125(a, b, w125, lane);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = lane + 32 * i;
                if (m < kPsM) {
                    const float2 v = z[m];
                    z[m] = make_float2(v.x * (1.0f / kPsM), -v.y * (1.0f / kPsM));
                }
            }
        }
        __syncthreads();
        // overlap-add of the round's frames, in frame order: sample s = hop j0 + i, frame f covers [hop f, hop f + n_fft)
        constexpr int span = kPsHop * (kPsIstftWarps - 1) + kPsNfft;
        const int sidx = kPsHop * j0 + tid;
        if (tid < span && sidx >= s_lo && sidx < s_hi) {
            float sum = ola[sidx - s_lo];
#pragma unroll
            for (int f = 0; f < kPsIstftWarps; ++f) {
                const int off = tid - kPsHop * f;
                // three passes leave a warp's result in its second buffer
                if (off >= 0 && off < kPsNfft && frame_live[f]) sum += reinterpret_cast<const float*>(buf[f][1])[off];
            }
            ola[sidx - s_lo] = sum;
        }
        __syncthreads();
    }
    // window envelope of a rectangular window = the number of frames covering the sample (over ALL frames of the clip)
    for (int sidx = s_lo + tid; sidx < s_hi; sidx += blockDim.x) {
        const int t_hi = min(sidx / kPsHop, Fo - 1), t_lo = max((sidx - (kPsNfft - 1) + kPsHop - 1) / kPsHop, 0);
        ola[sidx - s_lo] = ola[sidx - s_lo] / (float)(t_hi - t_lo + 1);
    }
    __syncthreads();
    // torchaudio resample: out[b up + i] = sum_j kernel[i][j] ypad[b orig + j], ypad = y shifted right by `width`, zeros outside
    for (int o = o_lo + tid; o < o_hi; o += blockDim.x) {
        float acc = 0.f;
        if (o < out_len) {
            const int bq = o / up, i = o - bq * up;
            const int base = bq * orig - width;
            const int j_lo = max(0, -base), j_hi = min(taps, istft_len - base);
            const float* yv = ola + (base - y_lo);
            for (int jj = j_lo; jj < j_hi; ++jj) acc = fmaf(__ldg(kernel_t + (int64_t)jj * up + i), yv[jj], acc);
        }
        dst[o] = acc;
    }
}

}  // namespace hb

using namespace hb;

// floats of y one segment of outputs reads (+ slack)
static size_t istft_smem_bytes(int T, int orig, int up, int taps) {
    const int seg_len = (T + kPsSegments - 1) / kPsSegments;
    return ((size_t)(seg_len / up + 2) * orig + taps + 16) * sizeof(float);
}

// One plan per (clip length, pitch ratio): the host computes the tables (heybuddy_b200/dataset/k9.py: time steps of the vocoder,
// phase advance, resampling kernel) and this copies them to the current device.
extern "C" int hb_pitch_plan_create(hb_pitch_plan** out, int T, int n_fft, int hop, int frames_out, const int32_t* idx0_host,
                                    const int32_t* idx1_host, const float* alpha_host, const float* phase_advance_host, int orig, int up, int width,
                                    const float* kernel_host) {
    HB_REQUIRE(out && idx0_host && idx1_host && alpha_host && phase_advance_host && kernel_host, "hb_pitch_plan_create: null argument");
    HB_REQUIRE(n_fft == kPsNfft && hop == kPsHop, "hb_pitch_plan_create: only n_fft = 250, hop = 7 (16 kHz) is built, got %d / %d", n_fft, hop);
    HB_REQUIRE(T >= kPsNfft && frames_out >= 1 && orig >= 1 && up >= 1 && width >= 0, "hb_pitch_plan_create: bad geometry");
    hb_pitch_plan* p = new hb_pitch_plan();
    p->T = T;
    p->frames_in = 1 + T / hop;
    p->frames_out = frames_out;
    p->istft_len = hop * (frames_out - 1);
    p->orig = orig;
    p->up = up;
    p->width = width;
    p->taps = 2 * width + orig;
    p->out_len = (int)(((int64_t)up * p->istft_len + orig - 1) / orig);
    HB_CUDA_OK(cudaGetDevice(&p->dev));
    HB_REQUIRE(istft_smem_bytes(T, orig, up, p->taps) <= 160 * 1024, "hb_pitch_plan_create: ratio %d/%d needs too much shared memory", up, orig);
    std::vector<float> kt((size_t)p->taps * up);
    for (int i = 0; i < up; ++i)
        for (int j = 0; j < p->taps; ++j) kt[(size_t)j * up + i] = kernel_host[(size_t)i * p->taps + j];
    std::vector<float2> w125(kPsM), w250(kPsBins);
    const double two_pi = 6.283185307179586476925286766559;
    for (int m = 0; m < kPsM; ++m) w125[m] = make_float2((float)cos(two_pi * m / kPsM), (float)-sin(two_pi * m / kPsM));
    for (int k = 0; k < kPsBins; ++k) w250[k] = make_float2((float)cos(two_pi * k / kPsNfft), (float)-sin(two_pi * k / kPsNfft));
    HB_CUDA_OK(cudaMalloc(&p->idx0, frames_out * sizeof(int32_t)));
    HB_CUDA_OK(cudaMalloc(&p->idx1, frames_out * sizeof(int32_t)));
    HB_CUDA_OK(cudaMalloc(&p->alpha, frames_out * sizeof(float)));
    HB_CUDA_OK(cudaMalloc(&p->phase_adv, kPsBins * sizeof(float)));
    HB_CUDA_OK(cudaMalloc(&p->kernel_t, kt.size() * sizeof(float)));
    HB_CUDA_OK(cudaMalloc(&p->w125, kPsM * sizeof(float2)));
    HB_CUDA_OK(cudaMalloc(&p->w250, kPsBins * sizeof(float2)));
    HB_CUDA_OK(cudaMemcpy(p->idx0, idx0_host, frames_out * sizeof(int32_t), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(p->idx1, idx1_host, frames_out * sizeof(int32_t), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(p->alpha, alpha_host, frames_out * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(p->phase_adv, phase_advance_host, kPsBins * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(p->kernel_t, kt.data(), kt.size() * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(p->w125, w125.data(), kPsM * sizeof(float2), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(p->w250, w250.data(), kPsBins * sizeof(float2), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaFuncSetAttribute(ps_istft_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    *out = p;
    return HB_OK;
}

extern "C" int hb_pitch_plan_destroy(hb_pitch_plan* p) {
    if (!p) return HB_OK;
    cudaFree(p->idx0);
    cudaFree(p->idx1);
    cudaFree(p->alpha);
    cudaFree(p->phase_adv);
    cudaFree(p->kernel_t);
    cudaFree(p->w125);
    cudaFree(p->w250);
    delete p;
    return HB_OK;
}

// Bytes of workspace hb_k9_pitch_f32 needs for k clips: the polar spectrogram and the stretched spectrogram of one group.
extern "C" int64_t hb_k9_pitch_workspace_bytes(const hb_pitch_plan* p, int k) {
    if (!p || k < 0) return HB_ERR_INVALID;
    const int64_t g = k < kPsGroup ? k : kPsGroup;
    return g * (((int64_t)p->frames_in + p->frames_out) * kPsBins * (int64_t)sizeof(float2) + (int64_t)kPsChunks * kPsBins * sizeof(double)) + 256;
}

// In place on the k clips listed in clip_index_dev (rows of clips_dev f32 [n][T]).  After the call the workspace holds, for the last
// group of <= 512 clips, the polar spectrogram f32 [g][frames_in][126][2] = (|S|, angle S) followed by the stretched complex
// spectrogram f32 [g][frames_out][126][2] (the parity tests read them).
extern "C" int hb_k9_pitch_f32(const hb_pitch_plan* p, float* clips_dev, const int32_t* clip_index_dev, int k, void* workspace_dev,
                               int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(p && k >= 0 && (k == 0 || (clips_dev && clip_index_dev && workspace_dev)), "hb_k9_pitch_f32: bad argument");
    if (k == 0) return HB_OK;
    HB_REQUIRE(workspace_bytes >= hb_k9_pitch_workspace_bytes(p, k), "hb_k9_pitch_f32: workspace too small");
    HB_REQUIRE((reinterpret_cast<uintptr_t>(workspace_dev) & 15) == 0, "hb_k9_pitch_f32: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int F = p->frames_in, Fo = p->frames_out;
    float2* S = reinterpret_cast<float2*>(workspace_dev);
    const size_t ola_bytes = istft_smem_bytes(p->T, p->orig, p->up, p->taps);
    for (int g0 = 0; g0 < k; g0 += kPsGroup) {
        const int g = k - g0 < kPsGroup ? k - g0 : kPsGroup;
        float2* Y = S + (int64_t)g * F * kPsBins;
        double* part = reinterpret_cast<double*>(Y + (int64_t)g * Fo * kPsBins);
        ps_stft_polar_kernel<<<dim3(ceil_div(F, kPsStftWarps), g), kPsStftWarps * 32, 0, st>>>(clips_dev, clip_index_dev + g0, S, p->T, F, p->w125,
                                                                                                 p->w250);
        HB_LAUNCHED();
        ps_vocoder_kernel<0><<<dim3(g, kPsChunks), 128, 0, st>>>(S, Y, part, F, Fo, p->idx0, p->idx1, p->alpha, p->phase_adv);
        HB_LAUNCHED();
        ps_vocoder_kernel<1><<<dim3(g, kPsChunks), 128, 0, st>>>(S, Y, part, F, Fo, p->idx0, p->idx1, p->alpha, p->phase_adv);
        HB_LAUNCHED();
        ps_istft_resample_kernel<<<dim3(kPsSegments, g), kPsIstftWarps * 32, ola_bytes, st>>>(Y, clips_dev, clip_index_dev + g0, p->T, Fo, p->istft_len, p->orig, p->up,
                                                                            p->width, p->taps, p->out_len, p->kernel_t, p->w125, p->w250);
        HB_LAUNCHED();
    }
    return HB_OK;
}
