// K7 (parity mode): speech-embedding conv stack as fp32 CUDA-core direct convolutions, plus the
// C-ABI entry points of the embedding model (they dispatch to embed_tc.cu for HB_EMBED_F16).
//
// Replaces the ORT run of speech-embedding.onnx behind SpeechEmbeddingModel.__call__
// (reference src/python/heybuddy/embeddings.py:32-42).  Layer table: embed_common.cuh.
//
// This is the bit-for-bit-style parity path (fp32 FMA, no operand rounding): one kernel per conv
// layer over NHWC activations, weights streamed from L2, a block computes 4 output time rows x all
// (freq, cout) so each weight load feeds 4 FMAs.  It is deliberately simple; the product path is
// the tcgen05 implicit GEMM in embed_tc.cu.  Both evaluate a clip fully convolutionally
// (SURVEY.md A.5): the 16 overlapping 76-frame windows of a 141-frame clip share every conv up to
// conv2d_15, only the last 2x2 pool and block 5 run per pool phase.
#include "embed_common.cuh"

#include <vector>

namespace hb {

constexpr int kRowsPerBlock = 4;
constexpr int kConvThreads = 256;

struct ConvArgs {
    const float* x;   // [n][T][F][Cin]
    const float* w;   // [KH][KW][Cin][Cout]
    const float* b;   // [Cout]
    float* y;         // [n][To][Fo][Cout]
    int T, F, Cin, Cout, To, Fo, KH, KW, padw, leaky;
};

__global__ void __launch_bounds__(kConvThreads) conv_rows_kernel(ConvArgs a) {
    extern __shared__ float xs[];  // [(RB+KH-1)][Fw][Cin]
    const int n = blockIdx.y;
    const int t0 = blockIdx.x * kRowsPerBlock;
    const int Fw = a.F + 2 * a.padw;
    const int rows = kRowsPerBlock + a.KH - 1;
    const int row_elems = Fw * a.Cin;
    const float* xin = a.x + (int64_t)n * a.T * a.F * a.Cin;
    for (int i = threadIdx.x; i < rows * row_elems; i += kConvThreads) {
        const int r = i / row_elems;
        const int rem = i - r * row_elems;
        const int fw = rem / a.Cin;
        const int c = rem - fw * a.Cin;
        const int t = t0 + r, f = fw - a.padw;
        float v = 0.f;
        if (t < a.T && f >= 0 && f < a.F) v = xin[((int64_t)t * a.F + f) * a.Cin + c];
        xs[i] = v;
    }
    __syncthreads();
    const int outs = a.Fo * a.Cout;
    for (int o = threadIdx.x; o < outs; o += kConvThreads) {
        const int f = o / a.Cout;
        const int co = o - f * a.Cout;
        float acc[kRowsPerBlock];
#pragma unroll
        for (int r = 0; r < kRowsPerBlock; ++r) acc[r] = 0.f;
        for (int kh = 0; kh < a.KH; ++kh)
            for (int kw = 0; kw < a.KW; ++kw) {
                const float* wp = a.w + ((int64_t)(kh * a.KW + kw) * a.Cin) * a.Cout + co;
                const float* xp = xs + (kh * Fw + f + kw) * a.Cin;
                for (int ci = 0; ci < a.Cin; ++ci) {
                    const float wv = __ldg(wp + (int64_t)ci * a.Cout);
#pragma unroll
                    for (int r = 0; r < kRowsPerBlock; ++r) acc[r] = fmaf(xp[r * row_elems + ci], wv, acc[r]);
                }
            }
        const float bias = __ldg(a.b + co);
#pragma unroll
        for (int r = 0; r < kRowsPerBlock; ++r) {
            const int t = t0 + r;
            if (t < a.To) {
                float v = acc[r] + bias;
                if (a.leaky) v = v > 0.f ? v : kLeaky * v;
                a.y[(((int64_t)n * a.To + t) * a.Fo + f) * a.Cout + co] = v;
            }
        }
    }
}

__global__ void maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int T, int F, int C,
                               int pt, int pf, int phase, int To, int Fo) {
    const int64_t total = (int64_t)n * To * Fo * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t r = i / C;
        const int fo = (int)(r % Fo);
        r /= Fo;
        const int to = (int)(r % To);
        const int b = (int)(r / To);
        float m = -INFINITY;
        for (int dt = 0; dt < pt; ++dt)
            for (int df = 0; df < pf; ++df) {
                const float v = x[(((int64_t)b * T + to * pt + phase + dt) * F + fo * pf + df) * C + c];
                m = (v > m || v != v) ? v : m;  // NaN-propagating max, like torch.max_pool2d
            }
        y[i] = m;
    }
}

// out[b][s][:] = tmp_phase(s)[b][j(s)][:]
__global__ void gather_slots_kernel(const float* __restrict__ tmp0, const float* __restrict__ tmp1, int J0, int J1,
                                    const int* __restrict__ slot_m, int n_slots, float* __restrict__ out, int B) {
    const int64_t total = (int64_t)B * n_slots * kEmbDim;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % kEmbDim);
        const int64_t r = i / kEmbDim;
        const int s = (int)(r % n_slots);
        const int b = (int)(r / n_slots);
        const int m = slot_m[s];
        const int j = m >> 1;
        out[i] = (m & 1) ? tmp1[((int64_t)b * J1 + j) * kEmbDim + c] : tmp0[((int64_t)b * J0 + j) * kEmbDim + c];
    }
}

static int launch_conv(const hb_embed_model* m, int li, const float* x, float* y, int n, int T, int F, int* To,
                       int* Fo, cudaStream_t st) {
    const ConvLayer& L = kLayers[li];
    ConvArgs a;
    a.x = x;
    a.w = m->w32 + m->w_off[li];
    a.b = m->w32 + m->b_off[li];
    a.y = y;
    a.T = T; a.F = F; a.Cin = L.cin; a.Cout = L.cout; a.KH = L.kh; a.KW = L.kw;
    a.padw = L.same ? L.kw / 2 : 0;
    a.To = T - L.kh + 1;
    a.Fo = L.same ? F : F - L.kw + 1;
    a.leaky = L.leaky;
    *To = a.To; *Fo = a.Fo;
    if (a.To <= 0 || a.Fo <= 0 || n == 0) return HB_OK;
    const size_t smem = (size_t)(kRowsPerBlock + L.kh - 1) * (F + 2 * a.padw) * L.cin * sizeof(float);
    dim3 grid(ceil_div(a.To, kRowsPerBlock), n);
    conv_rows_kernel<<<grid, kConvThreads, smem, st>>>(a);
    HB_LAUNCHED();
    return HB_OK;
}

static int launch_pool(const float* x, float* y, int n, int T, int F, int C, int pt, int pf, int phase, int* To,
                       int* Fo, cudaStream_t st) {
    *To = (T - phase) / pt;
    *Fo = F / pf;
    const int64_t total = (int64_t)n * (*To) * (*Fo) * C;
    if (total <= 0) return HB_OK;
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), 148 * 16);
    maxpool_kernel<<<blocks, 256, 0, st>>>(x, y, n, T, F, C, pt, pf, phase, *To, *Fo);
    HB_LAUNCHED();
    return HB_OK;
}

// Largest activation (floats per clip) for a strip of F frames: conv2d..conv2d_3 outputs.
static int64_t fp32_buf_floats(int B, int F) { return (int64_t)B * F * kMels * 24; }

int64_t fp32_workspace_bytes(int B, int F) {
    // two ping-pong activation buffers + two per-phase output buffers + slot table
    const int64_t act = fp32_buf_floats(B, F) * sizeof(float);
    const int64_t tail = (int64_t)B * (F / 8 + 2) * kEmbDim * sizeof(float);
    return 2 * act + 2 * tail + 4096;
}

// Runs layers [0, stop_layer] (phase 0 only past conv2d_15) and returns the last activation.
static int run_fp32(const hb_embed_model* m, const float* mel, int B, int F, int stop_layer, int final_phase,
                    float* bufA, float* bufB, const float** out_act, int* oT, int* oF, int* oC, cudaStream_t st) {
    const float* cur = mel;
    float* bufs[2] = {bufA, bufB};
    int which = 0;
    int T = F, Fq = kMels, C = 1;
    for (int li = 0; li <= stop_layer; ++li) {
        const ConvLayer& L = kLayers[li];
        int To, Fo;
        int rc = launch_conv(m, li, cur, bufs[which], B, T, Fq, &To, &Fo, st);
        if (rc) return rc;
        cur = bufs[which];
        which ^= 1;
        T = To; Fq = Fo; C = L.cout;
        if (L.pool_t > 1 || L.pool_f > 1) {
            const int phase = (li == 15) ? final_phase : 0;
            rc = launch_pool(cur, bufs[which], B, T, Fq, C, L.pool_t, L.pool_f, phase, &To, &Fo, st);
            if (rc) return rc;
            cur = bufs[which];
            which ^= 1;
            T = To; Fq = Fo;
        }
        if (T <= 0) break;
    }
    *out_act = cur; *oT = T; *oF = Fq; *oC = C;
    return HB_OK;
}

// Last 2x2 max-pool (both phases) + block 5 (conv2d_16..19) + slot gather, from the pre-pool conv2d_15 output
// f32 [B][T15][4][96].  Window offset 8j uses pool phase 0 position j, offset 8j+4 pool phase 1 position j.
int fp32_tail_from_l15(const hb_embed_model* m, const float* pre_pool, int B, int T15, const int32_t* slot_offsets_host,
                       int n_slots, float* out, float* scratch, int64_t scratch_floats, cudaStream_t st) {
    const int64_t rows = T15 / 2 + 1;
    const int64_t tail_floats = (int64_t)B * rows * kEmbDim;
    HB_REQUIRE(scratch_floats >= 6 * tail_floats + 1024, "hb_embed: tail scratch too small (%lld < %lld)",
               (long long)scratch_floats, (long long)(6 * tail_floats + 1024));
    float* tmp[2] = {scratch, scratch + tail_floats};
    float* s0 = scratch + 2 * tail_floats;
    float* s1 = s0 + 2 * tail_floats;
    int* slot_m_dev = reinterpret_cast<int*>(s1 + 2 * tail_floats);

    bool need_phase[2] = {false, false};
    std::vector<int> slot_m(n_slots);
    HB_REQUIRE(n_slots <= 1024, "hb_embed_clips: too many slots (%d)", n_slots);
    for (int s = 0; s < n_slots; ++s) {
        const int off = slot_offsets_host[s];
        HB_REQUIRE(off >= 0 && off % 4 == 0, "hb_embed_clips: slot offset %d must be a non-negative multiple of 4", off);
        slot_m[s] = off / 4;
        need_phase[slot_m[s] & 1] = true;
    }
    HB_CUDA_OK(cudaMemcpyAsync(slot_m_dev, slot_m.data(), n_slots * sizeof(int), cudaMemcpyHostToDevice, st));
    int J[2] = {0, 0};
    for (int p = 0; p < 2; ++p) {
        if (!need_phase[p]) continue;
        int To, Fo;
        int rc = launch_pool(pre_pool, s0, B, T15, 4, kEmbDim, 2, 2, p, &To, &Fo, st);
        if (rc) return rc;
        int Tc = To, Fc = Fo;
        const float* c = s0;
        float* nxt = s1;
        for (int li = 16; li < kNumConv; ++li) {
            float* dst = (li == kNumConv - 1) ? tmp[p] : nxt;
            rc = launch_conv(m, li, c, dst, B, Tc, Fc, &To, &Fo, st);
            if (rc) return rc;
            c = dst;
            nxt = (nxt == s1) ? s0 : s1;
            Tc = To; Fc = Fo;
        }
        J[p] = Tc;
    }
    for (int s = 0; s < n_slots; ++s) {
        const int p = slot_m[s] & 1, j = slot_m[s] >> 1;
        HB_REQUIRE(j < J[p], "hb_embed_clips: slot %d (offset %d) beyond the strip (phase %d has %d outputs)", s,
                   slot_offsets_host[s], p, J[p]);
    }
    const int64_t total = (int64_t)B * n_slots * kEmbDim;
    if (total > 0) {
        const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), 148 * 8);
        gather_slots_kernel<<<blocks, 256, 0, st>>>(tmp[0], tmp[1], J[0], J[1], slot_m_dev, n_slots, out, B);
        HB_LAUNCHED();
    }
    return HB_OK;
}

int fp32_gather_slots(const float* tmp0, const float* tmp1, int J0, int J1, const int* slot_m_dev, int n_slots, float* out, int B,
                      cudaStream_t st) {
    const int64_t total = (int64_t)B * n_slots * kEmbDim;
    if (total <= 0) return HB_OK;
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), 148 * 8);
    gather_slots_kernel<<<blocks, 256, 0, st>>>(tmp0, tmp1, J0, J1, slot_m_dev, n_slots, out, B);
    HB_LAUNCHED();
    return HB_OK;
}

int fp32_pool_public(const float* x, float* y, int n, int T, int F, int C, int pt, int pf, int phase, cudaStream_t st) {
    int To, Fo;
    return launch_pool(x, y, n, T, F, C, pt, pf, phase, &To, &Fo, st);
}

int fp32_embed_clips(const hb_embed_model* m, const float* mel, int B, int F, const int32_t* slot_offsets_host,
                     int n_slots, float* out, void* ws, int64_t ws_bytes, cudaStream_t st) {
    HB_REQUIRE(ws_bytes >= fp32_workspace_bytes(B, F), "hb_embed: workspace too small (%lld < %lld)",
               (long long)ws_bytes, (long long)fp32_workspace_bytes(B, F));
    float* bufA = reinterpret_cast<float*>(ws);
    float* bufB = bufA + fp32_buf_floats(B, F);
    // shared trunk: conv2d .. conv2d_14, then conv2d_15 without its pool so both phases can pool from it
    const float* act; int T, Fq, C;
    int rc = run_fp32(m, mel, B, F, 14, 0, bufA, bufB, &act, &T, &Fq, &C, st);
    if (rc) return rc;
    float* free_buf = (act == bufA) ? bufB : bufA;
    int T15, F15;
    rc = launch_conv(m, 15, act, free_buf, B, T, Fq, &T15, &F15, st);
    if (rc) return rc;
    float* scratch = (free_buf == bufA) ? bufB : bufA;  // the trunk's input is no longer needed
    return fp32_tail_from_l15(m, free_buf, B, T15, slot_offsets_host, n_slots, out, scratch, fp32_buf_floats(B, F), st);
}

int64_t fp32_activation(const hb_embed_model* m, const float* mel, int B, int F, int layer, float* out,
                        int64_t cap, void* ws, int64_t ws_bytes, cudaStream_t st) {
    if (ws_bytes < fp32_workspace_bytes(B, F)) {
        set_error("hb_embed_activation: workspace too small");
        return HB_ERR_INVALID;
    }
    float* bufA = reinterpret_cast<float*>(ws);
    float* bufB = bufA + fp32_buf_floats(B, F);
    const float* act; int T, Fq, C;
    int rc = run_fp32(m, mel, B, F, layer, 0, bufA, bufB, &act, &T, &Fq, &C, st);
    if (rc) return rc;
    const int64_t n = (int64_t)B * T * Fq * C;
    if (n > cap) {
        set_error("hb_embed_activation: output capacity %lld < %lld", (long long)cap, (long long)n);
        return HB_ERR_INVALID;
    }
    if (cudaMemcpyAsync(out, act, n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("hb_embed_activation: copy failed");
        return HB_ERR_CUDA;
    }
    return n;
}

}  // namespace hb

using namespace hb;

extern "C" int64_t hb_embed_num_params(void) { return total_weight_floats(); }

extern "C" int hb_embed_create(hb_embed_model** out, const float* weights_host, int64_t n_floats) {
    HB_REQUIRE(out && weights_host, "hb_embed_create: null pointer");
    HB_REQUIRE(n_floats == total_weight_floats(), "hb_embed_create: expected %lld floats, got %lld",
               (long long)total_weight_floats(), (long long)n_floats);
    hb_embed_model* m = new hb_embed_model();
    int64_t off = 0;
    for (int i = 0; i < kNumConv; ++i) {
        m->w_off[i] = off;
        off += layer_weight_floats(kLayers[i]);
        m->b_off[i] = off;
        off += kLayers[i].cout;
    }
    HB_CUDA_OK(cudaGetDevice(&m->device));
    HB_CUDA_OK(cudaMalloc(&m->w32, n_floats * sizeof(float)));
    HB_CUDA_OK(cudaMemcpy(m->w32, weights_host, n_floats * sizeof(float), cudaMemcpyHostToDevice));
    int rc = tc_prepare(m, weights_host);
    if (rc != HB_OK) {
        cudaFree(m->w32);
        delete m;
        return rc;
    }
    *out = m;
    return HB_OK;
}

extern "C" int hb_embed_destroy(hb_embed_model* m) {
    if (!m) return HB_OK;
    tc_release(m);
    if (m->w32) cudaFree(m->w32);
    delete m;
    return HB_OK;
}

extern "C" int64_t hb_embed_clips_workspace_bytes(int B, int F, int mode) {
    if (B < 0 || F < 0) return HB_ERR_INVALID;
    return mode == HB_EMBED_F16 ? tc_workspace_bytes(B, F) : fp32_workspace_bytes(B, F);
}

extern "C" int64_t hb_embed_windows_workspace_bytes(int n, int mode) {
    return hb_embed_clips_workspace_bytes(n, kEmbWindow, mode);
}

extern "C" int hb_embed_clips(const hb_embed_model* m, int mode, const float* mel_dev, int B, int F,
                              const int32_t* slot_offsets_host, int n_slots, float* out_dev, void* workspace_dev,
                              int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && mel_dev && out_dev && workspace_dev && slot_offsets_host, "hb_embed_clips: null pointer");
    HB_REQUIRE(B >= 0 && F >= kEmbWindow && n_slots > 0, "hb_embed_clips: bad shape B=%d F=%d n_slots=%d", B, F, n_slots);
    for (int s = 0; s < n_slots; ++s)
        HB_REQUIRE(slot_offsets_host[s] >= 0 && slot_offsets_host[s] % 4 == 0 && slot_offsets_host[s] + kEmbWindow <= F,
                   "hb_embed_clips: slot offset %d invalid for F=%d (must be a multiple of 4, offset+76 <= F)",
                   slot_offsets_host[s], F);
    HB_REQUIRE(B <= 65535, "hb_embed_clips: B=%d exceeds one launch", B);
    if (B == 0) return HB_OK;
    if (mode == HB_EMBED_FP32)
        return fp32_embed_clips(m, mel_dev, B, F, slot_offsets_host, n_slots, out_dev, workspace_dev, workspace_bytes,
                                (cudaStream_t)stream);
    if (mode == HB_EMBED_F16)
        return tc_embed_clips(m, mel_dev, B, F, slot_offsets_host, n_slots, out_dev, workspace_dev, workspace_bytes,
                              (cudaStream_t)stream);
    set_error("hb_embed_clips: unknown mode %d", mode);
    return HB_ERR_INVALID;
}

extern "C" int hb_embed_windows(const hb_embed_model* m, int mode, const float* windows_dev, float* out_dev, int n,
                                void* workspace_dev, int64_t workspace_bytes, void* stream) {
    const int32_t zero = 0;
    return hb_embed_clips(m, mode, windows_dev, n, kEmbWindow, &zero, 1, out_dev, workspace_dev, workspace_bytes, stream);
}

extern "C" int64_t hb_embed_activation(const hb_embed_model* m, int mode, const float* mel_dev, int B, int F, int layer,
                                       float* out_dev, int64_t out_capacity, void* workspace_dev,
                                       int64_t workspace_bytes, void* stream) {
    if (!m || !mel_dev || !out_dev || !workspace_dev || layer < 0 || layer >= kNumConv || B <= 0 || F < kEmbWindow) {
        set_error("hb_embed_activation: bad argument");
        return HB_ERR_INVALID;
    }
    if (mode == HB_EMBED_FP32)
        return fp32_activation(m, mel_dev, B, F, layer, out_dev, out_capacity, workspace_dev, workspace_bytes,
                               (cudaStream_t)stream);
    if (mode == HB_EMBED_F16)
        return tc_activation(m, mel_dev, B, F, layer, out_dev, out_capacity, workspace_dev, workspace_bytes,
                             (cudaStream_t)stream);
    set_error("hb_embed_activation: unknown mode %d", mode);
    return HB_ERR_INVALID;
}
