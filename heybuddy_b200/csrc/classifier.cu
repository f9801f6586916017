// K8: wake-word gated-MLP classifier -- batched forward, and one fused training step (forward,
// high-loss selection, weighted BCE, backward, Adam) -- as hand-written fp32 CUDA kernels.
//
// Replaces WakeWordMLPModel.forward (reference src/python/heybuddy/wakeword.py:334-348, with
// GatedMultiLayerPerceptron.forward, modules/multi_layer_perceptron.py:115-124) and the
// loss / backward / optimizer lines of WakeWordTrainer.train_epoch (trainer.py:405-462).
//
//   x[B,1536] -> LN(1536) -> gated(1536->64->96) -> 2 x [LN(96) -> gated(96->64->96)] -> LN(96)
//             -> gated(96->64->1) -> sigmoid            gated(u) = W_o (silu(W_h u + b_h) * (W_g u + b_g)) + b_o
//
// Parameters are packed in state_dict order (heybuddy_b200/spec.py classifier_param_shapes) so .pt
// checkpoints and the in-repo ONNX initializers interchange.  The step is launch-latency / HBM
// bound (25 MB of input per 4096-row batch, 1 MB of parameters); fp32 FMA keeps logits within 1e-3
// of the reference.
#include "mlp_common.cuh"

#include <math.h>

namespace hb {

static const MlpLayout kLayout = make_mlp_layout();

// ---- kernels -------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row.
__global__ void ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                              float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int B, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* xr = x + (int64_t)row * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i];
    const float mu = warp_sum(s) / D;
    float v = 0.f;
    for (int i = lane; i < D; i += 32) { const float d = xr[i] - mu; v += d * d; }
    const float rs = rsqrtf(warp_sum(v) / D + kLnEps);
    float* yr = y + (int64_t)row * D;
    for (int i = lane; i < D; i += 32) yr[i] = (xr[i] - mu) * rs * w[i] + b[i];
    if (lane == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
}

constexpr int kTile = 64, kTk = 16;   // tile of the model-batched linear kernel (config 5)

// ------------------------------------------------------------------------------------------------------------------------
// Register-blocked fp32 SGEMM for the step's large products: C[m,n] = sum_k A(m,k) B(k,n) (+ bias[n]).
//   * BM x BN x BK tiles (128 x 128 x 8 or 64 x 64 x 16), (BM/TM) x (BN/TN) threads, every thread a TM x TN micro-tile made of 4 x 4 blocks BM/2 (BN/2) apart, so
//     that each shared-memory read is a conflict-free LDS.128; double-buffered shared memory, the next k-tile's global loads in
//     flight (registers) while the current one is multiplied.
//   * either operand may be k-contiguous or m/n-contiguous (A_K / B_K): NT (activations x weights^T), TN (weight gradients:
//     d^T u) and NN (input gradients: d W) are the same kernel.
//   * the B operand (and the bias, and the output rows) may come from TWO parameter blocks -- the hidden and the gate linears of a
//     gated MLP share their input, so [W_h; W_g] is one stacked 128-row operand and the A tile is read once for both.
//   * gridDim.z > 1 = split-K: slice z writes its partial tile to part + z * M * N; sgemm_reduce_kernel adds the slices in a fixed
//     order (deterministic, unlike atomics) and applies bias / destinations.
struct SgemmArgs {
    const float* A; int lda;
    const float* B0; const float* B1; int bsplit; int ldb;     // B row r (r = n when B_K, else k): r < bsplit ? B0 + r ldb : B1 + (r - bsplit) ldb
    float* C0; float* C1; int csplit; int ldc;                   // C row m: m < csplit ? C0 + m ldc : C1 + (m - csplit) ldc
    const float* bias0; const float* bias1; int biassplit;       // bias[n]: n < biassplit ? bias0[n] : bias1[n - biassplit]; bias0 == nullptr: none
    float* part;                                                 // split-K partial buffer
    int M, N, K, k_per_split;
};

template <int BM, int BN, int BK, int TM, int TN, bool A_K, bool B_K>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) sgemm_kernel(const SgemmArgs a) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int A_LOADS = BM * BK / 4 / NT, B_LOADS = BN * BK / 4 / NT;   // float4 slots per thread per k-tile
    static_assert(TM % 4 == 0 && TN % 4 == 0 && A_LOADS >= 1 && B_LOADS >= 1, "tile shape");
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * a.k_per_split, k_end = min(a.K, k_begin + a.k_per_split);

    float4 ra[A_LOADS], rb[B_LOADS];
    // slot -> (row along the non-contiguous dim, 4 elements along the contiguous dim)
    auto load_a = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_LOADS; ++i) {
            const int slot = tid + i * NT;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (A_K) {
                const int m = m0 + slot / (BK / 4), k = k0 + (slot % (BK / 4)) * 4;
                if (m < a.M) {
                    const float* p = a.A + (int64_t)m * a.lda + k;
                    if (k + 3 < k_end && (reinterpret_cast<uintptr_t>(p) & 15) == 0) *reinterpret_cast<float4*>(v) = __ldg(reinterpret_cast<const float4*>(p));
                    else
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (k + j < k_end) v[j] = __ldg(p + j);
                }
            } else {
                const int k = k0 + slot / (BM / 4), m = m0 + (slot % (BM / 4)) * 4;
                if (k < k_end) {
                    const float* p = a.A + (int64_t)k * a.lda + m;
                    if (m + 3 < a.M && (reinterpret_cast<uintptr_t>(p) & 15) == 0) *reinterpret_cast<float4*>(v) = __ldg(reinterpret_cast<const float4*>(p));
                    else
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (m + j < a.M) v[j] = __ldg(p + j);
                }
            }
            ra[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
    };
    auto load_b = [&](int k0) {
#pragma unroll
        for (int i = 0; i < B_LOADS; ++i) {
            const int slot = tid + i * NT;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (B_K) {
                const int n = n0 + slot / (BK / 4), k = k0 + (slot % (BK / 4)) * 4;
                if (n < a.N) {
                    const float* p = (n < a.bsplit ? a.B0 + (int64_t)n * a.ldb : a.B1 + (int64_t)(n - a.bsplit) * a.ldb) + k;
                    if (k + 3 < k_end && (reinterpret_cast<uintptr_t>(p) & 15) == 0) *reinterpret_cast<float4*>(v) = __ldg(reinterpret_cast<const float4*>(p));
                    else
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (k + j < k_end) v[j] = __ldg(p + j);
                }
            } else {
                const int k = k0 + slot / (BN / 4), n = n0 + (slot % (BN / 4)) * 4;
                if (k < k_end) {
                    const float* p = (k < a.bsplit ? a.B0 + (int64_t)k * a.ldb : a.B1 + (int64_t)(k - a.bsplit) * a.ldb) + n;
                    if (n + 3 < a.N && (reinterpret_cast<uintptr_t>(p) & 15) == 0) *reinterpret_cast<float4*>(v) = __ldg(reinterpret_cast<const float4*>(p));
                    else
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (n + j < a.N) v[j] = __ldg(p + j);
                }
            }
            rb[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_LOADS; ++i) {
            const int slot = tid + i * NT;
            if (A_K) {
                const int m = slot / (BK / 4), k = (slot % (BK / 4)) * 4;
                As[buf][k][m] = ra[i].x; As[buf][k + 1][m] = ra[i].y; As[buf][k + 2][m] = ra[i].z; As[buf][k + 3][m] = ra[i].w;
            } else {
                const int k = slot / (BM / 4), m = (slot % (BM / 4)) * 4;
                *reinterpret_cast<float4*>(&As[buf][k][m]) = ra[i];
            }
        }
#pragma unroll
        for (int i = 0; i < B_LOADS; ++i) {
            const int slot = tid + i * NT;
            if (B_K) {
                const int n = slot / (BK / 4), k = (slot % (BK / 4)) * 4;
                Bs[buf][k][n] = rb[i].x; Bs[buf][k + 1][n] = rb[i].y; Bs[buf][k + 2][n] = rb[i].z; Bs[buf][k + 3][n] = rb[i].w;
            } else {
                const int k = slot / (BN / 4), n = (slot % (BN / 4)) * 4;
                *reinterpret_cast<float4*>(&Bs[buf][k][n]) = rb[i];
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    int buf = 0;
    if (k_begin < k_end) {
        load_a(k_begin);
        load_b(k_begin);
        store_tiles(0);
    }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool more = k0 + BK < k_end;
        if (more) {
            load_a(k0 + BK);
            load_b(k0 + BK);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM / 4; ++i)
                *reinterpret_cast<float4*>(&av[4 * i]) = *reinterpret_cast<const float4*>(&As[buf][k][i * (BM / (TM / 4)) + ty * 4]);
#pragma unroll
            for (int j = 0; j < TN / 4; ++j)
                *reinterpret_cast<float4*>(&bv[4 * j]) = *reinterpret_cast<const float4*>(&Bs[buf][k][j * (BN / (TN / 4)) + tx * 4]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            store_tiles(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    // epilogue: micro-tile row i -> m0 + (i / 4) * (BM / (TM / 4)) + ty * 4 + i % 4, same for columns
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * (BM / (TM / 4)) + ty * 4 + (i % 4);
        if (m >= a.M) continue;
        float* crow = split ? a.part + ((int64_t)blockIdx.z * a.M + m) * a.N
                            : (m < a.csplit ? a.C0 + (int64_t)m * a.ldc : a.C1 + (int64_t)(m - a.csplit) * a.ldc);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + (j / 4) * (BN / (TN / 4)) + tx * 4 + (j % 4);
            if (n >= a.N) continue;
            float v = acc[i][j];
            if (!split && a.bias0 != nullptr) v += (n < a.biassplit ? a.bias0[n] : a.bias1[n - a.biassplit]);
            crow[n] = v;
        }
    }
}

__global__ void sgemm_reduce_kernel(const SgemmArgs a, int S) {
    const int64_t total = (int64_t)a.M * a.N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / a.N), n = (int)(i - (int64_t)m * a.N);
        float v = a.bias0 != nullptr ? (n < a.biassplit ? a.bias0[n] : a.bias1[n - a.biassplit]) : 0.f;
        for (int z = 0; z < S; ++z) v += a.part[(int64_t)z * total + i];
        float* crow = m < a.csplit ? a.C0 + (int64_t)m * a.ldc : a.C1 + (int64_t)(m - a.csplit) * a.ldc;
        crow[n] = v;
    }
}

__global__ void gate_fwd_kernel(const float* __restrict__ h, const float* __restrict__ g, float* __restrict__ a, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float hv = h[i];
        a[i] = hv / (1.f + expf(-hv)) * g[i];
    }
}

// stacked layout: hg [B][128] = [hidden | gate] pre-activations of a gated MLP; a [B][64] = silu(h) * g
__global__ void gate_hg_fwd_kernel(const float* __restrict__ hg, float* __restrict__ a, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / kHid;
        const int j = (int)(i - r * kHid);
        const float hv = hg[r * 2 * kHid + j], gv = hg[r * 2 * kHid + kHid + j];
        a[i] = hv / (1.f + expf(-hv)) * gv;
    }
}
__global__ void gate_hg_bwd_kernel(const float* __restrict__ hg, const float* __restrict__ da, float* __restrict__ dhg, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / kHid;
        const int j = (int)(i - r * kHid);
        const float hv = hg[r * 2 * kHid + j], gv = hg[r * 2 * kHid + kHid + j];
        const float sg = 1.f / (1.f + expf(-hv));
        const float d = da[i];
        dhg[r * 2 * kHid + kHid + j] = d * hv * sg;                                   // d gate
        dhg[r * 2 * kHid + j] = d * gv * (sg * (1.f + hv * (1.f - sg)));              // d hidden
    }
}

__global__ void sigmoid_kernel(const float* __restrict__ z, float* __restrict__ p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1.f / (1.f + expf(-z[i]));
}

// stats = {loss_sum -> loss, n_selected, stepped, high_loss_rate}
__global__ void head_select_kernel(const float* __restrict__ z, const int64_t* __restrict__ y, float* __restrict__ p,
                                   float* __restrict__ stats, int B, float thr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int sel = 0;
    if (i < B) {
        const float pr = 1.f / (1.f + expf(-z[i]));
        p[i] = pr;
        sel = (y[i] == 0 && pr >= thr) || (y[i] == 1 && pr < 1.f - thr);
    }
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&stats[1], (float)__popc(m));
}

// n_total: rows selected over the whole (possibly multi-rank) batch = the normaliser of the mean loss
__global__ void head_grad_kernel(const float* __restrict__ p, const int64_t* __restrict__ y, float* __restrict__ dz,
                                 float* __restrict__ stats, const float* __restrict__ n_total, int B, float thr, float neg_w, float loss_scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float n_sel = *n_total / loss_scale;     // loss_scale = 1 / accumulation_steps of the reference's loop
    float l = 0.f;
    if (i < B) {
        const float pr = p[i];
        const bool pos = y[i] == 1;
        const bool sel = (!pos && pr >= thr) || (pos && pr < 1.f - thr);
        float g = 0.f;
        if (sel && n_sel > 0.f) {
            const float w = pos ? 1.f : neg_w;
            // F.binary_cross_entropy clamps log() at -100
            const float lp = fmaxf(logf(pr), -100.f), l1p = fmaxf(logf(1.f - pr), -100.f);
            l = -w * (pos ? lp : l1p) / n_sel;
            g = w * (pr - (pos ? 1.f : 0.f)) / n_sel;   // d loss / d logit
        }
        dz[i] = g;
    }
    l = warp_sum(l);
    if ((threadIdx.x & 31) == 0 && l != 0.f) atomicAdd(&stats[0], l);
}

__global__ void head_finish_kernel(float* __restrict__ stats, const float* __restrict__ n_total, int B, int min_selected) {
    stats[2] = *n_total >= (float)min_selected ? 1.f : 0.f;
    stats[3] = stats[1] / (float)B;
}

// The single-device training step's whole head in one CTA: sigmoid, high-loss selection, the selected count as the normaliser, weighted
// BCE and its gradient, the step's statistics (what memset + head_select + head_grad + head_finish do in four launches when the count
// has to cross ranks in between).  Sums in a fixed order: the reported loss is deterministic.
__global__ void __launch_bounds__(1024) head_train_kernel(const float* __restrict__ z, const int64_t* __restrict__ y, float* __restrict__ p,
                                                          float* __restrict__ dz, float* __restrict__ stats, int B, float thr, float neg_w,
                                                          float loss_scale, int min_selected, float* __restrict__ exchange_tail) {
    __shared__ float red[32];
    __shared__ float total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto block_sum = [&](float v) {
        v = warp_sum(v);
        __syncthreads();                // the previous use of red / total is over
        if (lane == 0) red[warp] = v;
        __syncthreads();
        if (warp == 0) {
            float t = warp_sum(red[lane]);
            if (lane == 0) total = t;
        }
        __syncthreads();
        return total;
    };
    float cnt = 0.f;
    for (int i = tid; i < B; i += 1024) {
        const float pr = 1.f / (1.f + expf(-z[i]));
        p[i] = pr;
        cnt += ((y[i] == 0 && pr >= thr) || (y[i] == 1 && pr < 1.f - thr)) ? 1.f : 0.f;
    }
    const float n_total = block_sum(cnt);
    // exchange_tail: the one-collective data-parallel form -- sums stay unnormalised (normaliser 1), the division follows the all-reduce
    const float n_sel = (exchange_tail != nullptr ? 1.f : n_total) / loss_scale;
    float l = 0.f;
    for (int i = tid; i < B; i += 1024) {
        const float pr = p[i];
        const bool pos = y[i] == 1;
        const bool sel = (!pos && pr >= thr) || (pos && pr < 1.f - thr);
        float g = 0.f;
        if (sel && n_sel > 0.f) {
            const float w = pos ? 1.f : neg_w;
            const float lp = fmaxf(logf(pr), -100.f), l1p = fmaxf(logf(1.f - pr), -100.f);
            l += -w * (pos ? lp : l1p) / n_sel;
            g = w * (pr - (pos ? 1.f : 0.f)) / n_sel;
        }
        dz[i] = g;
    }
    const float loss = block_sum(l);
    if (tid == 0) {
        stats[0] = loss;
        stats[1] = n_total;
        stats[2] = n_total >= (float)min_selected ? 1.f : 0.f;
        stats[3] = n_total / (float)B;
        if (exchange_tail != nullptr) { exchange_tail[0] = loss; exchange_tail[1] = n_total; }
    }
}

// Column reductions over the batch, two deterministic passes: blockIdx.y owns a slice of the rows and writes its partial sums to
// scratch [gridDim.y][2][N]; colred_finish_kernel adds the slices in order.  MODE 0: sum_m X[m,n] (bias gradients);
// MODE 1: LayerNorm parameter gradients, first = sum_m dy * xhat (gamma), second = sum_m dy (beta).
constexpr int kColSlices = 16;
template <int MODE>
__global__ void colred_kernel(const float* __restrict__ X, const float* __restrict__ x_in, const float* __restrict__ mean,
                              const float* __restrict__ rstd, float* __restrict__ scratch, int M, int N) {
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const int part = threadIdx.x >> 5;  // 8 row partitions per slice
    const int rows = (M + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * rows, r1 = min(M, r0 + rows);
    __shared__ float r_a[8][33], r_b[8][33];
    float a = 0.f, b = 0.f;
    if (n < N)
        for (int m = r0 + part; m < r1; m += 8) {
            const float d = X[(int64_t)m * N + n];
            if (MODE == 1) { a += d * (x_in[(int64_t)m * N + n] - mean[m]) * rstd[m]; b += d; }
            else a += d;
        }
    r_a[part][threadIdx.x & 31] = a;
    r_b[part][threadIdx.x & 31] = b;
    __syncthreads();
    if (part == 0 && n < N) {
        float ta = 0.f, tb = 0.f;
        for (int i = 0; i < 8; ++i) { ta += r_a[i][threadIdx.x & 31]; tb += r_b[i][threadIdx.x & 31]; }
        scratch[((int64_t)blockIdx.y * 2 + 0) * N + n] = ta;
        scratch[((int64_t)blockIdx.y * 2 + 1) * N + n] = tb;
    }
}
// first sums: column n < split -> out0[n], else out1[n - split]; second sums (MODE 1 only) -> out_b[n]
__global__ void colred_finish_kernel(const float* __restrict__ scratch, float* __restrict__ out0, float* __restrict__ out1, int split,
                                     float* __restrict__ out_b, int slices, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float ta = 0.f, tb = 0.f;
    for (int s = 0; s < slices; ++s) { ta += scratch[((int64_t)s * 2 + 0) * N + n]; tb += scratch[((int64_t)s * 2 + 1) * N + n]; }
    if (n < split) out0[n] = ta; else out1[n - split] = ta;
    if (out_b != nullptr) out_b[n] = tb;
}

// dx = rstd * (dyg - mean(dyg) - xhat * mean(dyg * xhat)), dyg = dy * gamma; one warp per row
__global__ void ln_bwd_input_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
                                    int B, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float mu = mean[row], rs = rstd[row];
    const float* dyr = dy + (int64_t)row * D;
    const float* xr = x + (int64_t)row * D;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float g = dyr[i] * gamma[i], xh = (xr[i] - mu) * rs;
        s1 += g;
        s2 += g * xh;
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    for (int i = lane; i < D; i += 32) {
        const float g = dyr[i] * gamma[i], xh = (xr[i] - mu) * rs;
        dx[(int64_t)row * D + i] = rs * (g - s1 - xh * s2);
    }
}

// torch.optim.Adam defaults (betas 0.9/0.999, eps 1e-8, no weight decay); a no-op when stats[2] == 0
// One launch: every block reads the step counter before it can change -- the last block to finish (a ticket kept next to the counter,
// step[1]) advances it.
// exchange != nullptr (the one-collective data-parallel form): the gradients are exchange[0, n) / n_total with n_total = exchange[n + 1]
// (summed over the ranks), the loss exchange[n] / n_total; the model's gradient buffer and stats[0..2] are filled on the way.
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int* __restrict__ step, float* __restrict__ stats, float lr, int n, const float* __restrict__ exchange,
                            int min_selected) {
    float gscale = 1.f;
    bool stepped;
    if (exchange != nullptr) {
        const float n_total = exchange[n + 1];
        gscale = n_total > 0.f ? 1.f / n_total : 0.f;
        stepped = n_total >= (float)min_selected;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            stats[0] = exchange[n] * gscale;
            stats[1] = n_total;
            stats[2] = stepped ? 1.f : 0.f;
        }
    } else {
        stepped = stats[2] != 0.f;
    }
    const float* __restrict__ src = exchange != nullptr ? exchange : g;
    if (!stepped) {
        if (exchange != nullptr)
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) g[i] = src[i] * gscale;
        return;
    }
    const int t = *reinterpret_cast<volatile int*>(step) + 1;
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float bc1 = 1.f - powf(b1, (float)t), bc2 = 1.f - powf(b2, (float)t);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float gi = src[i] * gscale;
        if (exchange != nullptr) g[i] = gi;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= (lr / bc1) * mi / (sqrtf(vi) / sqrtf(bc2) + eps);
    }
    __syncthreads();                                     // every thread of this block has read the counter
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int* ticket = reinterpret_cast<unsigned int*>(step + 1);
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0;
            *step = t;
        }
    }
}

}  // namespace hb


namespace hb {

constexpr int64_t kPartFloats = 32ll * 128 * 1536;    // split-K partial buffer: up to 32 slices of the largest (stacked) weight gradient

// ---- SGEMM driver ------------------------------------------------------------------------------------------------------
static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

template <int BM, int BN, int BK, int TM, int TN>
static void sgemm_launch(const SgemmArgs& a, bool a_k, bool b_k, dim3 grid, cudaStream_t st) {
    constexpr int NT = (BM / TM) * (BN / TN);
    if (a_k && b_k) sgemm_kernel<BM, BN, BK, TM, TN, true, true><<<grid, NT, 0, st>>>(a);
    else if (a_k) sgemm_kernel<BM, BN, BK, TM, TN, true, false><<<grid, NT, 0, st>>>(a);
    else if (b_k) sgemm_kernel<BM, BN, BK, TM, TN, false, true><<<grid, NT, 0, st>>>(a);
    else sgemm_kernel<BM, BN, BK, TM, TN, false, false><<<grid, NT, 0, st>>>(a);
}

// part / part_floats: split-K scratch (nullptr: never split).  Large outputs get 128 x 128 tiles, small ones 64 x 64; products
// whose tiles do not fill the GPU and whose K is long (the weight gradients: K = batch; the first layer: K = 1536) are split over K.
// HB_MLP_FMA=1: keep every product on the fp32 FMA kernels (parity mode / A-B comparisons)
static bool tensor_path_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("HB_MLP_FMA");
        on = (e && e[0] == '1') ? 0 : 1;
    }
    return on == 1;
}

static int sgemm(SgemmArgs a, bool a_k, bool b_k, cudaStream_t st, float* part, int64_t part_floats) {
    if (a.M <= 0 || a.N <= 0) return HB_OK;
    // The first-layer x W^T products (K = 1536; both operands K-major) go to the tensor cores, three TF32 passes = fp32-level accuracy
    // (gemm_tf32.cu) -- whatever the batch size, so that a row's result does not depend on how many rows travel with it (the streaming
    // service and the offline strip evaluation agree).  The 96- and 64-wide products stay here: at K < 512 the FMA kernel is faster.
    if (a_k && b_k && a.C1 == nullptr && a.K >= 512 && tensor_path_enabled() &&
        gemm_tf32x3_ok(a.A, a.lda, a.B0, a.B1, a.ldb, a.C0, a.ldc, a.M, a.N, a.K))
        return gemm_tf32x3_tn(a.A, a.lda, a.B0, a.B1, a.bsplit, a.ldb, a.bias0, a.bias1, a.biassplit, a.C0, a.ldc, a.M, a.N, a.K, st);
    if (a.C1 == nullptr) { a.C1 = a.C0; a.csplit = a.M; }
    if (a.B1 == nullptr) { a.B1 = a.B0; a.bsplit = 1 << 30; }
    if (a.bias1 == nullptr) { a.bias1 = a.bias0; a.biassplit = 1 << 30; }
    // 128 x 128 tiles when they (times the K slices a long K allows) fill the GPU; otherwise 64 x 64 tiles for more CTAs
    const int tiles128 = ceil_div(a.M, 128) * ceil_div(a.N, 128);
    const bool big = a.M >= 128 && a.N >= 128 && (tiles128 >= sm_count() * 3 / 4 || (part != nullptr && a.K >= 1024));
    const int bm = big ? 128 : 64, bn = big ? 128 : 64;
    dim3 grid(ceil_div(a.N, bn), ceil_div(a.M, bm), 1);
    const int tiles = (int)(grid.x * grid.y);
    int S = 1;
    if (part != nullptr && tiles < sm_count() && a.K >= 512) {
        S = std::min(std::min(a.K / 256, 32), ceil_div(2 * sm_count(), tiles));     // slices of at least 256 k
        while (S > 1 && (int64_t)S * a.M * a.N > part_floats) --S;
    }
    a.k_per_split = S > 1 ? ceil_div(ceil_div(a.K, S), 16) * 16 : a.K;
    a.part = part;
    grid.z = S > 1 ? ceil_div(a.K, a.k_per_split) : 1;
    if (big) sgemm_launch<128, 128, 8, 8, 8>(a, a_k, b_k, grid, st);
    else sgemm_launch<64, 64, 16, 4, 4>(a, a_k, b_k, grid, st);
    HB_LAUNCHED();
    if (grid.z > 1) {
        const int64_t total = (int64_t)a.M * a.N;
        sgemm_reduce_kernel<<<(int)std::min<int64_t>(ceil_div64(total, 256), 8 * sm_count()), 256, 0, st>>>(a, (int)grid.z);
        HB_LAUNCHED();
    }
    return HB_OK;
}

static SgemmArgs sgemm_args(const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias, int M, int N, int K) {
    SgemmArgs a;
    a.A = A; a.lda = lda;
    a.B0 = B; a.B1 = nullptr; a.bsplit = 0; a.ldb = ldb;
    a.C0 = C; a.C1 = nullptr; a.csplit = 0; a.ldc = ldc;
    a.bias0 = bias; a.bias1 = nullptr; a.biassplit = 0;
    a.part = nullptr;
    a.M = M; a.N = N; a.K = K; a.k_per_split = K;
    return a;
}

// bias gradients: column sums of X [M][N] -> out0 (columns < split) / out1
static int colsum(const float* X, int M, int N, float* out0, float* out1, int split, float* scratch, cudaStream_t st) {
    const int slices = std::max(1, std::min(kColSlices, M / 64));
    colred_kernel<0><<<dim3(ceil_div(N, 32), slices), 256, 0, st>>>(X, nullptr, nullptr, nullptr, scratch, M, N);
    HB_LAUNCHED();
    colred_finish_kernel<<<ceil_div(N, 256), 256, 0, st>>>(scratch, out0, out1 ? out1 : out0, out1 ? split : N, nullptr, slices, N);
    HB_LAUNCHED();
    return HB_OK;
}
// LayerNorm parameter gradients: dgamma[n] = sum_m dy * xhat, dbeta[n] = sum_m dy
static int ln_param_grads(const float* dy, const float* x, const float* mean, const float* rstd, float* dgamma, float* dbeta, int M, int N,
                          float* scratch, cudaStream_t st) {
    const int slices = std::max(1, std::min(kColSlices, M / 64));
    colred_kernel<1><<<dim3(ceil_div(N, 32), slices), 256, 0, st>>>(dy, x, mean, rstd, scratch, M, N);
    HB_LAUNCHED();
    colred_finish_kernel<<<ceil_div(N, 256), 256, 0, st>>>(scratch, dgamma, dgamma, N, dbeta, slices, N);
    HB_LAUNCHED();
    return HB_OK;
}

// y[B,N] = x[B,K] W[N,K]^T + b  (kept for the stacked multi-model forward)
static int linear_fwd(const float* x, const float* W, const float* b, float* y, int B, int N, int K, cudaStream_t st, float* part = nullptr) {
    return sgemm(sgemm_args(x, K, W, K, y, N, b, B, N, K), true, true, st, part, kPartFloats);
}

// workspace carve-up (floats)
struct Ws {
    float *xn, *mean[kStages], *rstd[kStages];
    float *u[kStages];       // LN output = stage input (u[0] = xn)
    float *hg[kStages];      // [B][128]: hidden | gate pre-activations (one stacked product)
    float *a[kStages], *o[kStages];
    float *dz, *d_o[2], *d_a, *d_hg, *d_u;
    float* part;             // split-K partial sums
    float* colred;           // column-reduction partials [kColSlices][2][1536]
    float* fused;            // workspace of the fused step (mlp_fused.cu), behind the staged path's buffers
    int training;
};
static int64_t carve(Ws* w, float* base, int B, int training) {
    int64_t off = 0;
    auto take = [&](int64_t n) { float* p = base ? base + off : nullptr; off += (n + 63) & ~63ll; return p; };
    for (int s = 0; s < kStages; ++s) {
        const int in_dim = kLayout.s[s].in_dim, out_dim = kLayout.s[s].out_dim;
        w->u[s] = take((int64_t)B * in_dim);
        w->mean[s] = take(B);
        w->rstd[s] = take(B);
        w->hg[s] = take((int64_t)B * 2 * kHid);
        w->a[s] = take((int64_t)B * kHid);
        w->o[s] = take((int64_t)B * out_dim);
    }
    w->xn = w->u[0];
    if (training) {
        w->dz = take(B);
        w->d_o[0] = take((int64_t)B * kDim);
        w->d_o[1] = take((int64_t)B * kDim);
        w->d_a = take((int64_t)B * kHid);
        w->d_hg = take((int64_t)B * 2 * kHid);
        w->d_u = take((int64_t)B * kIn);
    }
    w->part = take(kPartFloats);
    w->colred = take((int64_t)kColSlices * 2 * kIn);
    w->fused = base ? base + off : nullptr;
    w->training = training;
    off += mlp_fused_ws_floats(B, training);
    return off;
}

// HB_MLP_STAGED=1 (or HB_MLP_FMA=1): the one-kernel-per-operation path below instead of the fused step of mlp_fused.cu (parity mode)
static bool fused_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("HB_MLP_STAGED");
        on = ((e && e[0] == '1') || !tensor_path_enabled()) ? 0 : 1;
    }
    return on == 1;
}

// logits of the batch -> *logits (a workspace buffer)
static int forward_impl(const hb_mlp_model* m, const float* x, int B, const Ws& w, cudaStream_t st, const float** logits) {
    if (fused_enabled()) return mlp_fused_forward(m, x, B, w.fused, w.training, logits, st);
    *logits = w.o[kStages - 1];
    const float* cur = x;
    for (int s = 0; s < kStages; ++s) {
        const StageOff& L = kLayout.s[s];
        ln_fwd_kernel<<<ceil_div(B, 8), 256, 0, st>>>(cur, m->p + L.ln_w, m->p + L.ln_b, w.u[s], w.mean[s], w.rstd[s], B, L.in_dim);
        HB_LAUNCHED();
        int rc;
        // hidden and gate linears share their input: ONE product against the stacked [W_h; W_g]
        SgemmArgs hg = sgemm_args(w.u[s], L.in_dim, m->p + L.hw, L.in_dim, w.hg[s], 2 * kHid, m->p + L.hb, B, 2 * kHid, L.in_dim);
        hg.B1 = m->p + L.gw; hg.bsplit = kHid;
        hg.bias1 = m->p + L.gb; hg.biassplit = kHid;
        if ((rc = sgemm(hg, true, true, st, w.part, kPartFloats))) return rc;
        const int64_t n = (int64_t)B * kHid;
        gate_hg_fwd_kernel<<<(int)std::min<int64_t>(ceil_div64(n, 256), 1184), 256, 0, st>>>(w.hg[s], w.a[s], n);
        HB_LAUNCHED();
        if ((rc = sgemm(sgemm_args(w.a[s], kHid, m->p + L.ow, kHid, w.o[s], L.out_dim, m->p + L.ob, B, L.out_dim, kHid), true, true, st, nullptr, 0))) return rc;
        cur = w.o[s];
    }
    return HB_OK;
}

}  // namespace hb

using namespace hb;

extern "C" int64_t hb_mlp_num_params(void) { return kLayout.total; }

extern "C" int hb_mlp_create(hb_mlp_model** out, const float* params_host, int64_t n_floats) {
    HB_REQUIRE(out && params_host, "hb_mlp_create: null pointer");
    HB_REQUIRE(n_floats == kLayout.total, "hb_mlp_create: expected %d floats (default architecture), got %lld", kLayout.total,
               (long long)n_floats);
    hb_mlp_model* m = new hb_mlp_model();
    const size_t bytes = (size_t)n_floats * sizeof(float);
    HB_CUDA_OK(cudaMalloc(&m->p, bytes));
    HB_CUDA_OK(cudaMalloc(&m->g, bytes));
    HB_CUDA_OK(cudaMalloc(&m->m, bytes));
    HB_CUDA_OK(cudaMalloc(&m->v, bytes));
    HB_CUDA_OK(cudaMalloc(&m->step, 2 * sizeof(int)));      // [step, the Adam kernel's block ticket]
    HB_CUDA_OK(cudaMemcpy(m->p, params_host, bytes, cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemset(m->g, 0, bytes));
    HB_CUDA_OK(cudaMemset(m->m, 0, bytes));
    HB_CUDA_OK(cudaMemset(m->v, 0, bytes));
    HB_CUDA_OK(cudaMemset(m->step, 0, 2 * sizeof(int)));
    *out = m;
    return HB_OK;
}

extern "C" int hb_mlp_destroy(hb_mlp_model* m) {
    if (!m) return HB_OK;
    cudaFree(m->p); cudaFree(m->g); cudaFree(m->m); cudaFree(m->v); cudaFree(m->step);
    delete m;
    return HB_OK;
}

extern "C" int hb_mlp_get_params(const hb_mlp_model* m, float* params_host, int64_t n_floats) {
    HB_REQUIRE(m && params_host && n_floats == kLayout.total, "hb_mlp_get_params: bad argument");
    HB_CUDA_OK(cudaMemcpy(params_host, m->p, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return HB_OK;
}

extern "C" int hb_mlp_set_params(hb_mlp_model* m, const float* params_host, int64_t n_floats) {
    HB_REQUIRE(m && params_host && n_floats == kLayout.total, "hb_mlp_set_params: bad argument");
    HB_CUDA_OK(cudaMemcpy(m->p, params_host, (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice));
    return HB_OK;
}

extern "C" int hb_mlp_get_grads(const hb_mlp_model* m, float* grads_host, int64_t n_floats) {
    HB_REQUIRE(m && grads_host && n_floats == kLayout.total, "hb_mlp_get_grads: bad argument");
    HB_CUDA_OK(cudaMemcpy(grads_host, m->g, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return HB_OK;
}

extern "C" int hb_mlp_set_loss_scale(hb_mlp_model* m, float scale) {
    HB_REQUIRE(m && scale > 0.f, "hb_mlp_set_loss_scale: bad argument");
    m->loss_scale = scale;
    return HB_OK;
}

// Adam state (torch.optim.Adam's exp_avg / exp_avg_sq / step) <-> host, packed like the parameters: checkpoint + resume
extern "C" int hb_mlp_get_adam(const hb_mlp_model* m, float* exp_avg_host, float* exp_avg_sq_host, int* step_host, int64_t n_floats) {
    HB_REQUIRE(m && exp_avg_host && exp_avg_sq_host && step_host && n_floats == kLayout.total, "hb_mlp_get_adam: bad argument");
    HB_CUDA_OK(cudaMemcpy(exp_avg_host, m->m, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    HB_CUDA_OK(cudaMemcpy(exp_avg_sq_host, m->v, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    HB_CUDA_OK(cudaMemcpy(step_host, m->step, sizeof(int), cudaMemcpyDeviceToHost));
    return HB_OK;
}

extern "C" int hb_mlp_set_adam(hb_mlp_model* m, const float* exp_avg_host, const float* exp_avg_sq_host, int step, int64_t n_floats) {
    HB_REQUIRE(m && exp_avg_host && exp_avg_sq_host && step >= 0 && n_floats == kLayout.total, "hb_mlp_set_adam: bad argument");
    HB_CUDA_OK(cudaMemcpy(m->m, exp_avg_host, (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(m->v, exp_avg_sq_host, (size_t)n_floats * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMemcpy(m->step, &step, sizeof(int), cudaMemcpyHostToDevice));
    return HB_OK;
}

namespace hb {
// one Philox4x32-10 block per four elements: counter = (index / 4, 7, call lo, call hi), key = seed (same generator as the draw table)
__device__ __forceinline__ uint4 philox_cls(uint4 c, uint2 k) {
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
// nn.Dropout(p) on the classifier input (wakeword.py:197,338): y = x * keep / (1 - p), keep ~ Bernoulli(1 - p) per element
__global__ void dropout_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t n4, float p, uint2 key, uint64_t call) {
    const float scale = 1.0f / (1.0f - p);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 r = philox_cls(make_uint4((uint32_t)i, 7u ^ ((uint32_t)(i >> 32) << 8), (uint32_t)call, (uint32_t)(call >> 32)), key);
        const float4 v = x[i];
        const float t = p * 16777216.0f;     // keep when the 24-bit uniform >= p
        y[i] = make_float4((float)(r.x >> 8) >= t ? v.x * scale : 0.f, (float)(r.y >> 8) >= t ? v.y * scale : 0.f,
                           (float)(r.z >> 8) >= t ? v.z * scale : 0.f, (float)(r.w >> 8) >= t ? v.w * scale : 0.f);
    }
}
}  // namespace hb

extern "C" int hb_mlp_dropout(const float* x_dev, float* y_dev, int64_t n, float p, uint64_t seed, uint64_t call, void* stream) {
    HB_REQUIRE(x_dev && y_dev && n >= 0 && n % 4 == 0 && p >= 0.f && p < 1.f, "hb_mlp_dropout: bad argument (n must be a multiple of 4)");
    HB_REQUIRE(((reinterpret_cast<uintptr_t>(x_dev) | reinterpret_cast<uintptr_t>(y_dev)) & 15) == 0, "hb_mlp_dropout: 16-byte aligned buffers");
    if (n == 0) return HB_OK;
    const int64_t n4 = n / 4;
    dropout_kernel<<<(int)std::min<int64_t>(ceil_div64(n4, 256), 2368), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x_dev), reinterpret_cast<float4*>(y_dev), n4, p, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), call);
    HB_LAUNCHED();
    return HB_OK;
}

extern "C" int64_t hb_mlp_workspace_bytes(int B, int training) {
    if (B < 0) return HB_ERR_INVALID;
    Ws w;
    return carve(&w, nullptr, B, training) * (int64_t)sizeof(float) + 256;
}

extern "C" int hb_mlp_forward(const hb_mlp_model* m, const float* x_dev, float* prob_dev, int B, void* workspace_dev,
                              int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && x_dev && prob_dev && workspace_dev, "hb_mlp_forward: null pointer");
    HB_REQUIRE(workspace_bytes >= hb_mlp_workspace_bytes(B, 0), "hb_mlp_forward: workspace too small");
    if (B == 0) return HB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    Ws w;
    carve(&w, reinterpret_cast<float*>(workspace_dev), B, 0);
    const float* logits = nullptr;
    int rc = forward_impl(m, x_dev, B, w, st, &logits);
    if (rc) return rc;
    sigmoid_kernel<<<ceil_div(B, 256), 256, 0, st>>>(logits, prob_dev, B);
    HB_LAUNCHED();
    return HB_OK;
}

// ---- config 5: M models on the same inputs, one launch chain --------------------------------------------------------------
// The models differ in every parameter, including norm_in's affine -- but LN(x) = xhat * gamma + beta with xhat = (x - mu) * rstd
// shared by all of them, so gamma folds into the first-layer weights and beta into their bias:
//     W'_m[n][k] = W_m[n][k] * gamma_m[k],   b'_m[n] = b_m[n] + sum_k W_m[n][k] * beta_m[k]
// and the 2 M first-layer products (hidden and gate of every model; 97 % of a forward pass's FLOPs) become ONE stacked GEMM
// xhat[B,1536] x W'[M*128,1536]^T that reads the input once.  The 96-wide remainder of every model runs as GEMMs batched over the
// models (blockIdx.z = model).  The fold is redone on every call (it reads the models' live parameters; 1 MB per model).
namespace hb {

constexpr int kMaxPack = kMlpMaxPack;
struct PtrPack { const float* p[kMaxPack]; };

// xhat = (x - mean) * rstd, one warp per row (LayerNorm without the affine)
__global__ void xhat_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* xr = x + (int64_t)row * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i];
    const float mu = warp_sum(s) / D;
    float v = 0.f;
    for (int i = lane; i < D; i += 32) { const float d = xr[i] - mu; v += d * d; }
    const float rs = rsqrtf(warp_sum(v) / D + kLnEps);
    float* yr = y + (int64_t)row * D;
    for (int i = lane; i < D; i += 32) yr[i] = (xr[i] - mu) * rs;
}

// one warp per stacked row (model m, n in [0,128): n < 64 hidden, else gate): W' row and b'
__global__ void fold_first_layer_kernel(PtrPack pk, int m0, int M, StageOff L, float* __restrict__ Wst, float* __restrict__ bst) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M * 2 * kHid) return;
    const int m = row / (2 * kHid), n = row - m * 2 * kHid;
    const float* p = pk.p[m];
    const float* W = p + (n < kHid ? L.hw + n * kIn : L.gw + (n - kHid) * kIn);
    const float* gamma = p + L.ln_w;
    const float* beta = p + L.ln_b;
    float* out = Wst + ((int64_t)(m0 + m) * 2 * kHid + n) * kIn;
    float acc = 0.f;
    for (int k = lane; k < kIn; k += 32) {
        const float w = W[k];
        out[k] = w * gamma[k];
        acc = fmaf(w, beta[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) bst[(int64_t)(m0 + m) * 2 * kHid + n] = acc + p[(n < kHid ? L.hb + n : L.gb + n - kHid)];
}

// HG [B][Mtot*128] -> A [Mtot][B][64] = silu(h) * g
__global__ void gate_stacked_kernel(const float* __restrict__ hg, float* __restrict__ a, int B, int M) {
    const int64_t n = (int64_t)B * M * kHid;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % kHid);
        const int64_t r = i / kHid;
        const int m = (int)(r % M);
        const int64_t b = r / M;
        const float hv = hg[(b * M + m) * 2 * kHid + j], gv = hg[(b * M + m) * 2 * kHid + kHid + j];
        a[((int64_t)m * B + b) * kHid + j] = hv / (1.f + expf(-hv)) * gv;
    }
}

// LayerNorm with per-model affine: x, y [M][B][D]; blockIdx.y = model
__global__ void ln_fwd_batched_kernel(const float* __restrict__ x, PtrPack pk, int w_off, int b_off, float* __restrict__ y, int B, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, m = blockIdx.y;
    if (row >= B) return;
    const float* w = pk.p[m] + w_off;
    const float* b = pk.p[m] + b_off;
    const float* xr = x + ((int64_t)m * B + row) * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i];
    const float mu = warp_sum(s) / D;
    float v = 0.f;
    for (int i = lane; i < D; i += 32) { const float d = xr[i] - mu; v += d * d; }
    const float rs = rsqrtf(warp_sum(v) / D + kLnEps);
    float* yr = y + ((int64_t)m * B + row) * D;
    for (int i = lane; i < D; i += 32) yr[i] = (xr[i] - mu) * rs * w[i] + b[i];
}

// y[m] [B][N] = x[m] [B][K] W_m[N][K]^T + b_m, batched over the models (blockIdx.z = model); same tile loop as gemm_kernel
__global__ void __launch_bounds__(256) linear_batched_kernel(const float* __restrict__ X, PtrPack pk, int w_off, int b_off,
                                                             float* __restrict__ Y, int B, int N, int K) {
    __shared__ float As[kTk][kTile + 1];
    __shared__ float Bs[kTk][kTile + 1];
    const int m = blockIdx.z;
    const float* A = X + (int64_t)m * B * K;
    const float* W = pk.p[m] + w_off;
    const float* bias = pk.p[m] + b_off;
    float* C = Y + (int64_t)m * B * N;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += kTk) {
        for (int i = threadIdx.x; i < kTile * kTk; i += 256) {
            const int k = i % kTk, r = i / kTk;
            As[k][r] = (m0 + r < B && k0 + k < K) ? A[(int64_t)(m0 + r) * K + k0 + k] : 0.f;
            Bs[k][r] = (n0 + r < N && k0 + k < K) ? W[(int64_t)(n0 + r) * K + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kTk; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty + 16 * i]; b[i] = Bs[k][tx + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gm = m0 + ty + 16 * i, gn = n0 + tx + 16 * j;
            if (gm < B && gn < N) C[(int64_t)gm * N + gn] = acc[i][j] + bias[gn];
        }
}

struct MultiWs {
    float *xhat, *Wst, *bst, *hg, *a, *o, *u, *h, *g;
    float* part;
    float* wt;       // transposed small weights of one pack of models (mlp_fused.cu)
};
static int64_t carve_multi(MultiWs* w, float* base, int M, int B) {
    int64_t off = 0;
    auto take = [&](int64_t n) { float* p = base ? base + off : nullptr; off += (n + 63) & ~63ll; return p; };
    w->xhat = take((int64_t)B * kIn);
    w->Wst = take((int64_t)M * 2 * kHid * kIn);
    w->bst = take((int64_t)M * 2 * kHid);
    w->hg = take((int64_t)B * M * 2 * kHid);
    w->a = take((int64_t)M * B * kHid);
    w->o = take((int64_t)M * B * kDim);
    w->u = take((int64_t)M * B * kDim);
    w->h = take((int64_t)M * B * kHid);
    w->g = take((int64_t)M * B * kHid);
    w->part = take(kPartFloats);
    w->wt = take(mlp_fused_multi_wt_floats(std::min(M, kMaxPack)));
    return off;
}

}  // namespace hb

extern "C" int64_t hb_mlp_multi_workspace_bytes(int M, int B) {
    if (M < 0 || B < 0) return HB_ERR_INVALID;
    MultiWs w;
    return carve_multi(&w, nullptr, M, B) * (int64_t)sizeof(float) + 256;
}

extern "C" int hb_mlp_forward_multi(hb_mlp_model* const* models, int M, const float* x_dev, float* prob_dev, int B,
                                    void* workspace_dev, int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(models && M >= 0 && B >= 0, "hb_mlp_forward_multi: bad argument");
    if (M == 0 || B == 0) return HB_OK;
    HB_REQUIRE(x_dev && prob_dev && workspace_dev, "hb_mlp_forward_multi: null pointer");
    HB_REQUIRE(workspace_bytes >= hb_mlp_multi_workspace_bytes(M, B), "hb_mlp_forward_multi: workspace too small (hb_mlp_multi_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    MultiWs w;
    carve_multi(&w, reinterpret_cast<float*>(workspace_dev), M, B);
    int rc;
    // 1. shared normalised input
    xhat_kernel<<<ceil_div(B, 8), 256, 0, st>>>(x_dev, w.xhat, B, kIn);
    HB_LAUNCHED();
    // 2. fold norm_in into the stacked first layer (packs of up to 64 models per launch: the pointers travel as kernel arguments)
    for (int m0 = 0; m0 < M; m0 += kMaxPack) {
        PtrPack pk;
        const int n = std::min(kMaxPack, M - m0);
        for (int i = 0; i < n; ++i) {
            HB_REQUIRE(models[m0 + i] != nullptr, "hb_mlp_forward_multi: null model %d", m0 + i);
            pk.p[i] = models[m0 + i]->p;
        }
        fold_first_layer_kernel<<<ceil_div(n * 2 * kHid, 8), 256, 0, st>>>(pk, m0, n, kLayout.s[0], w.Wst, w.bst);
        HB_LAUNCHED();
    }
    // 3. ONE stacked GEMM: hidden and gate pre-activations of every model
    if ((rc = linear_fwd(w.xhat, w.Wst, w.bst, w.hg, B, M * 2 * kHid, kIn, st, w.part))) return rc;
    if (fused_enabled()) {
        // 4. the 96-wide remainder of every model: one forward-tail CTA per (32 rows, model), the stacked output read in place
        for (int m0 = 0; m0 < M; m0 += kMaxPack) {
            const int n = std::min(kMaxPack, M - m0);
            const float* params[kMaxPack];
            for (int i = 0; i < n; ++i) params[i] = models[m0 + i]->p;
            float* logits = w.o + (int64_t)m0 * B;      // [n][B]
            if ((rc = mlp_fused_multi_tail(params, n, w.hg + (int64_t)m0 * 2 * kHid, (int64_t)M * 2 * kHid, B, w.wt, logits, st))) return rc;
            const int total = n * B;
            sigmoid_kernel<<<ceil_div(total, 256), 256, 0, st>>>(logits, prob_dev + (int64_t)m0 * B, total);
            HB_LAUNCHED();
        }
        return HB_OK;
    }
    const int64_t n_gate = (int64_t)B * M * kHid;
    const int gate_grid = (int)std::min<int64_t>(ceil_div64(n_gate, 256), 2368);
    gate_stacked_kernel<<<gate_grid, 256, 0, st>>>(w.hg, w.a, B, M);
    HB_LAUNCHED();
    // 4. the 96-wide remainder, batched over the models
    for (int m0 = 0; m0 < M; m0 += kMaxPack) {
        PtrPack pk;
        const int n = std::min(kMaxPack, M - m0);
        for (int i = 0; i < n; ++i) pk.p[i] = models[m0 + i]->p;
        float* a = w.a + (int64_t)m0 * B * kHid;
        float* o = w.o + (int64_t)m0 * B * kDim;
        float* u = w.u + (int64_t)m0 * B * kDim;
        float* h = w.h + (int64_t)m0 * B * kHid;
        float* g = w.g + (int64_t)m0 * B * kHid;
        const dim3 grid_o(ceil_div(kDim, kTile), ceil_div(B, kTile), n), grid_h(1, ceil_div(B, kTile), n);
        linear_batched_kernel<<<grid_o, 256, 0, st>>>(a, pk, kLayout.s[0].ow, kLayout.s[0].ob, o, B, kDim, kHid);
        HB_LAUNCHED();
        for (int s = 1; s < kStages; ++s) {
            const StageOff& L = kLayout.s[s];
            ln_fwd_batched_kernel<<<dim3(ceil_div(B, 8), n), 256, 0, st>>>(o, pk, L.ln_w, L.ln_b, u, B, kDim);
            HB_LAUNCHED();
            linear_batched_kernel<<<grid_h, 256, 0, st>>>(u, pk, L.hw, L.hb, h, B, kHid, kDim);
            HB_LAUNCHED();
            linear_batched_kernel<<<grid_h, 256, 0, st>>>(u, pk, L.gw, L.gb, g, B, kHid, kDim);
            HB_LAUNCHED();
            const int64_t ng = (int64_t)n * B * kHid;
            gate_fwd_kernel<<<(int)std::min<int64_t>(ceil_div64(ng, 256), 2368), 256, 0, st>>>(h, g, a, ng);
            HB_LAUNCHED();
            // stage 3's output is one logit per row: o is reused as [n][B][1]
            const dim3 grid_out(ceil_div(L.out_dim, kTile), ceil_div(B, kTile), n);
            linear_batched_kernel<<<grid_out, 256, 0, st>>>(a, pk, L.ow, L.ob, o, B, L.out_dim, kHid);
            HB_LAUNCHED();
        }
        const int total = n * B;   // logits [n][B] contiguous
        sigmoid_kernel<<<ceil_div(total, 256), 256, 0, st>>>(o, prob_dev + (int64_t)m0 * B, total);
        HB_LAUNCHED();
    }
    return HB_OK;
}

namespace hb {

// forward + sigmoid + high-loss selection: prob, stats[1] = rows selected in THIS batch (stats[0,2,3] zeroed)
static int select_impl(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float thr, float* prob_dev, float* stats_dev,
                       const Ws& w, cudaStream_t st) {
    const float* logits = nullptr;
    int rc = forward_impl(m, x_dev, B, w, st, &logits);
    if (rc) return rc;
    HB_CUDA_OK(cudaMemsetAsync(stats_dev, 0, 4 * sizeof(float), st));
    head_select_kernel<<<ceil_div(B, 256), 256, 0, st>>>(logits, y_dev, prob_dev, stats_dev, B, thr);
    HB_LAUNCHED();
    return HB_OK;
}

// weighted BCE over the selected rows divided by *n_total, and its gradient with respect to every parameter -> m->g
static int backward_impl(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float negative_weight, float thr,
                         const float* n_total_dev, int min_selected, const float* prob_dev, float* stats_dev, const Ws& w,
                         cudaStream_t st) {
    int rc;
    head_grad_kernel<<<ceil_div(B, 256), 256, 0, st>>>(prob_dev, y_dev, w.dz, stats_dev, n_total_dev, B, thr, negative_weight, m->loss_scale);
    HB_LAUNCHED();
    head_finish_kernel<<<1, 1, 0, st>>>(stats_dev, n_total_dev, B, min_selected);
    HB_LAUNCHED();
    if (fused_enabled()) return mlp_fused_backward(m, x_dev, B, w.fused, w.dz, st);
    const float* d_out = w.dz;   // gradient wrt the stage's output o[s]
    int flip = 0;
    for (int s = kStages - 1; s >= 0; --s) {
        const StageOff& L = kLayout.s[s];
        const int in_dim = L.in_dim, out_dim = L.out_dim;
        // output linear: dW_o [out,64] = d_out^T a, db_o = colsum(d_out), d_a [B,64] = d_out W_o
        if ((rc = sgemm(sgemm_args(d_out, out_dim, w.a[s], kHid, m->g + L.ow, kHid, nullptr, out_dim, kHid, B), false, false, st, w.part, kPartFloats))) return rc;
        if ((rc = colsum(d_out, B, out_dim, m->g + L.ob, nullptr, 0, w.colred, st))) return rc;
        if ((rc = sgemm(sgemm_args(d_out, out_dim, m->p + L.ow, kHid, w.d_a, kHid, nullptr, B, kHid, out_dim), true, false, st, nullptr, 0))) return rc;
        const int64_t n = (int64_t)B * kHid;
        gate_hg_bwd_kernel<<<(int)std::min<int64_t>(ceil_div64(n, 256), 1184), 256, 0, st>>>(w.hg[s], w.d_a, w.d_hg, n);
        HB_LAUNCHED();
        // hidden / gate linears, stacked: dW [128,in] = d_hg^T u (rows < 64 -> hidden, else gate), db = colsum(d_hg), d_u = d_hg [W_h; W_g]
        SgemmArgs dw = sgemm_args(w.d_hg, 2 * kHid, w.u[s], in_dim, m->g + L.hw, in_dim, nullptr, 2 * kHid, in_dim, B);
        dw.C1 = m->g + L.gw; dw.csplit = kHid;
        if ((rc = sgemm(dw, false, false, st, w.part, kPartFloats))) return rc;
        if ((rc = colsum(w.d_hg, B, 2 * kHid, m->g + L.hb, m->g + L.gb, kHid, w.colred, st))) return rc;
        SgemmArgs du = sgemm_args(w.d_hg, 2 * kHid, m->p + L.hw, in_dim, w.d_u, in_dim, nullptr, B, in_dim, 2 * kHid);
        du.B1 = m->p + L.gw; du.bsplit = kHid;
        if ((rc = sgemm(du, true, false, st, nullptr, 0))) return rc;
        // LayerNorm: parameter grads always, input grad unless the input is the data
        const float* ln_in = (s == 0) ? x_dev : w.o[s - 1];
        if ((rc = ln_param_grads(w.d_u, ln_in, w.mean[s], w.rstd[s], m->g + L.ln_w, m->g + L.ln_b, B, in_dim, w.colred, st))) return rc;
        if (s > 0) {
            ln_bwd_input_kernel<<<ceil_div(B, 8), 256, 0, st>>>(w.d_u, ln_in, m->p + L.ln_w, w.mean[s], w.rstd[s], w.d_o[flip], B, in_dim);
            HB_LAUNCHED();
            d_out = w.d_o[flip];
            flip ^= 1;
        }
    }
    return HB_OK;
}

static int adam_impl(hb_mlp_model* m, float lr, const float* stats_dev, cudaStream_t st) {
    adam_kernel<<<148, 256, 0, st>>>(m->p, m->g, m->m, m->v, m->step, const_cast<float*>(stats_dev), lr, kLayout.total, nullptr, 0);
    HB_LAUNCHED();
    return HB_OK;
}

}  // namespace hb

extern "C" int hb_mlp_train_step(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float lr,
                                 float negative_weight, float high_loss_threshold, int min_selected, float* prob_dev,
                                 float* stats_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && x_dev && y_dev && prob_dev && stats_dev && workspace_dev, "hb_mlp_train_step: null pointer");
    HB_REQUIRE(B > 0, "hb_mlp_train_step: empty batch");
    HB_REQUIRE(workspace_bytes >= hb_mlp_workspace_bytes(B, 1), "hb_mlp_train_step: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    Ws w;
    carve(&w, reinterpret_cast<float*>(workspace_dev), B, 1);
    int rc;
    if (fused_enabled()) {
        const float* logits = nullptr;
        if ((rc = forward_impl(m, x_dev, B, w, st, &logits))) return rc;
        head_train_kernel<<<1, 1024, 0, st>>>(logits, y_dev, prob_dev, w.dz, stats_dev, B, high_loss_threshold, negative_weight, m->loss_scale,
                                              min_selected, nullptr);
        HB_LAUNCHED();
        if ((rc = mlp_fused_backward(m, x_dev, B, w.fused, w.dz, st))) return rc;
        return adam_impl(m, lr, stats_dev, st);
    }
    if ((rc = select_impl(m, x_dev, y_dev, B, high_loss_threshold, prob_dev, stats_dev, w, st))) return rc;
    // single device: the batch's own selection count is the normaliser
    if ((rc = backward_impl(m, x_dev, y_dev, B, negative_weight, high_loss_threshold, stats_dev + 1, min_selected, prob_dev, stats_dev, w, st)))
        return rc;
    return adam_impl(m, lr, stats_dev, st);
}

// ---- data-parallel form of the same step: select -> (all-reduce n) -> backward -> (all-reduce grads) -> adam ----
extern "C" int hb_mlp_select(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float high_loss_threshold,
                             float* prob_dev, float* stats_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && x_dev && y_dev && prob_dev && stats_dev && workspace_dev && B > 0, "hb_mlp_select: bad argument");
    HB_REQUIRE(workspace_bytes >= hb_mlp_workspace_bytes(B, 1), "hb_mlp_select: workspace too small");
    Ws w;
    carve(&w, reinterpret_cast<float*>(workspace_dev), B, 1);
    return select_impl(m, x_dev, y_dev, B, high_loss_threshold, prob_dev, stats_dev, w, (cudaStream_t)stream);
}

extern "C" int hb_mlp_backward(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float negative_weight,
                               float high_loss_threshold, const float* n_selected_total_dev, int min_selected,
                               const float* prob_dev, float* stats_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && x_dev && y_dev && prob_dev && stats_dev && workspace_dev && n_selected_total_dev && B > 0, "hb_mlp_backward: bad argument");
    HB_REQUIRE(workspace_bytes >= hb_mlp_workspace_bytes(B, 1), "hb_mlp_backward: workspace too small");
    Ws w;
    carve(&w, reinterpret_cast<float*>(workspace_dev), B, 1);
    return backward_impl(m, x_dev, y_dev, B, negative_weight, high_loss_threshold, n_selected_total_dev, min_selected, prob_dev,
                         stats_dev, w, (cudaStream_t)stream);
}

extern "C" int hb_mlp_grads_copy(hb_mlp_model* m, float* buf_dev, int64_t n_floats, int to_model, void* stream) {
    HB_REQUIRE(m && buf_dev && n_floats == kLayout.total, "hb_mlp_grads_copy: expected %d floats", kLayout.total);
    HB_CUDA_OK(cudaMemcpyAsync(to_model ? m->g : buf_dev, to_model ? buf_dev : m->g, (size_t)n_floats * sizeof(float),
                               cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return HB_OK;
}

// One-collective data-parallel step, first half: forward, selection, UNNORMALISED weighted-BCE sum and gradients of this shard, packed
// for the exchange: exchange_dev f32 [hb_mlp_num_params() + 2] = {gradients | loss sum | rows selected}.
extern "C" int hb_mlp_local_step(hb_mlp_model* m, const float* x_dev, const int64_t* y_dev, int B, float negative_weight,
                                 float high_loss_threshold, float* prob_dev, float* stats_dev, float* exchange_dev, void* workspace_dev,
                                 int64_t workspace_bytes, void* stream) {
    HB_REQUIRE(m && x_dev && y_dev && prob_dev && stats_dev && exchange_dev && workspace_dev && B > 0, "hb_mlp_local_step: bad argument");
    HB_REQUIRE(workspace_bytes >= hb_mlp_workspace_bytes(B, 1), "hb_mlp_local_step: workspace too small");
    HB_REQUIRE(fused_enabled(), "hb_mlp_local_step: not available with HB_MLP_STAGED / HB_MLP_FMA (use hb_mlp_select / hb_mlp_backward)");
    cudaStream_t st = (cudaStream_t)stream;
    Ws w;
    carve(&w, reinterpret_cast<float*>(workspace_dev), B, 1);
    int rc;
    const float* logits = nullptr;
    if ((rc = forward_impl(m, x_dev, B, w, st, &logits))) return rc;
    head_train_kernel<<<1, 1024, 0, st>>>(logits, y_dev, prob_dev, w.dz, stats_dev, B, high_loss_threshold, negative_weight, m->loss_scale, 0,
                                          exchange_dev + kLayout.total);
    HB_LAUNCHED();
    return mlp_fused_backward(m, x_dev, B, w.fused, w.dz, st, exchange_dev);
}

// Second half, after ONE all-reduce (SUM) of exchange_dev: gradients and loss divided by the global count, stats_dev[0..2] = {mean loss
// over all selected rows, rows selected on all ranks, stepped}, Adam (skipped below min_selected) -- identical on every rank.
extern "C" int hb_mlp_apply_exchange(hb_mlp_model* m, const float* exchange_dev, float lr, int min_selected, float* stats_dev, void* stream) {
    HB_REQUIRE(m && exchange_dev && stats_dev, "hb_mlp_apply_exchange: null pointer");
    adam_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(m->p, m->g, m->m, m->v, m->step, stats_dev, lr, kLayout.total, exchange_dev, min_selected);
    HB_LAUNCHED();
    return HB_OK;
}

extern "C" int hb_mlp_adam(hb_mlp_model* m, float lr, const float* stats_dev, void* stream) {
    HB_REQUIRE(m && stats_dev, "hb_mlp_adam: null pointer");
    return adam_impl(m, lr, stats_dev, (cudaStream_t)stream);
}
