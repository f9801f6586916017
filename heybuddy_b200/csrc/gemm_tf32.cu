// Classifier (K8): C[M, N] = A[M, K] B[N, K]^T + bias on the tensor cores with fp32 accuracy -- the first-layer products of the gated
// MLP (reference wakeword.py:334-348: 1536 -> 64 + 64 per model; 97 % of a forward pass) for one model (hidden | gate stacked) and
// for M models stacked (hb_mlp_forward_multi).
//
// tcgen05.mma.kind::tf32 reads fp32 words from shared memory and uses their upper 19 bits.  Classifier logits have to stay within 1e-3
// of the reference; one TF32 pass rounds each operand to 2^-11, which puts a K = 1536 dot product at a few 1e-4 relative -- on the
// bound.  Three passes restore fp32: with a = a_hi + a_lo (a_hi = what the hardware sees, a_lo = a - a_hi, exact),
//     a b  ~=  a_hi b_hi + a_lo b_hi + a_hi b_lo            (the dropped a_lo b_lo term is 2^-22 relative)
// all three accumulating in the same fp32 TMEM tile.  Measured against float64: see tests/test_classifier_gpu.py.
//
// One CTA = one 128 x 128 tile of C, K steps of 32 through three shared-memory stages.  Warps 0-7 produce: each thread brings its
// share of the A and B tiles from global memory (registers, loads two steps ahead) and writes the raw words and the lo words as
// K-major core-matrix tiles ([k / 4][row][4 floats], chunk pitch 129 rows so that a quarter-warp's eight chunks fall into eight bank
// groups), then one of them arrives on the stage's `full` barrier.  Warp 8 consumes: one thread waits for `full`, issues 4 x 3 MMAs
// (M = 128, N = 128, K = 8) and commits to the stage's `empty` barrier, which the producers wait on before they overwrite the stage.
// No CTA-wide barrier inside the loop (the first version had two per step and ran at a third of this one's rate).
// The epilogue goes through shared memory so that C is written in whole 128-byte row segments.
#include "mlp_common.cuh"
#include "tc_ptx.cuh"

#include <algorithm>

namespace hb {

namespace {

constexpr int kTfBM = 128, kTfBN = 128, kTfBK = 32, kTfProducers = 256, kTfThreads = kTfProducers + 32;
constexpr int kTfChunks = kTfBK / 4;                    // 16-byte K chunks per step
constexpr int kTfPitch = 129 * 16;                      // bytes between K chunks of a tile (128 rows + one row of skew)
constexpr int kTfTile = kTfChunks * kTfPitch;           // one operand tile: 16.5 KB
constexpr int kTfStage = 4 * kTfTile;                   // A, A_lo, B, B_lo


constexpr int kTfStages = 3;
#ifndef HB_TF_AHEAD
#define HB_TF_AHEAD 2
#endif
constexpr int kTfAhead = HB_TF_AHEAD;      // K steps of global loads in flight per producer thread (register sets)

struct TfHeader {
    uint64_t full[kTfStages], empty[kTfStages];
    uint32_t tmem;
    uint32_t pad;
};

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// instruction descriptor: c_format F32 (bit 4), a / b format TF32 (2 at [7,10) and [10,13)), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kMode 0: both operands K-major (A [M][K], B [N][K]).  kMode 1: both batch-major (A [K][M], B [K][N]): the producers transpose 4 x 4
// blocks in registers on the way to the same K-major shared-memory tiles.
template <int kMode>
__device__ __forceinline__ void gemm_tf32x3_tile(const TfArgs& a, const int tile_n, const int tile_m, const int slice) {
    extern __shared__ __align__(128) unsigned char smem[];
    TfHeader& hdr = *reinterpret_cast<TfHeader*>(smem);
    unsigned char* stage0 = smem + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = tile_m * kTfBM, n0 = tile_n * kTfBN;
    if (tid == 0) {
        for (int i = 0; i < kTfStages; ++i) { mbar_init(&hdr.full[i], 1); mbar_init(&hdr.empty[i], 1); }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&hdr.tmem, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr.tmem;
    constexpr uint32_t idesc = make_idesc_tf32(kTfBM, kTfBN);

    // K slice of this CTA: steps [s_begin, s_end) of 32
    const int steps_total = (a.K + kTfBK - 1) / kTfBK;
    const int s_begin = (int)((int64_t)slice * steps_total / a.splits), s_end = (int)((int64_t)(slice + 1) * steps_total / a.splits);
    const int steps = s_end - s_begin;
    const int kbase = s_begin * kTfBK;

    // thread -> 4 (row, chunk) slots of each operand tile.  kMode 0: slot = tid + 256 i, chunk = slot % 8 (a row's 128 bytes are 8 lanes),
    // row = slot / 8.  kMode 1: chunk = warp, row = lane + 32 i (a warp reads 128 consecutive floats of four k rows).
    constexpr int kSlots = kTfBM * kTfChunks / kTfProducers;    // 4
    float4 ra[kTfAhead][kSlots], rb[kTfAhead][kSlots];          // register sets: the loads run kTfAhead steps ahead of their use
    auto slot_c = [&](int i) { return kMode == 0 ? ((tid + i * kTfProducers) & (kTfChunks - 1)) : warp; };
    auto slot_r = [&](int i) { return kMode == 0 ? ((tid + i * kTfProducers) / kTfChunks) : (lane + 32 * i); };
    float mu[kSlots], rs[kSlots];
#pragma unroll
    for (int i = 0; i < kSlots; ++i) {
        const int m = m0 + slot_r(i);
        const bool on = kMode == 0 && a.mean != nullptr && m < a.M;
        mu[i] = on ? __ldg(a.mean + m) : 0.f;
        rs[i] = on ? __ldg(a.rstd + m) : 1.f;
    }
    // loads leave the words as they came (nothing may depend on them until `store`, two steps later); the row statistics are applied there
    auto load = [&](int k0, float4 (&xa)[kSlots], float4 (&xb)[kSlots], float4& km, float4& kr) {
        if (kMode == 0) {
#pragma unroll
            for (int i = 0; i < kSlots; ++i) {
                const int c = slot_c(i), r = slot_r(i);
                const int m = m0 + r, n = n0 + r;
                xa[i] = m < a.M ? __ldg(reinterpret_cast<const float4*>(a.A + (int64_t)m * a.lda + k0) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float* brow = n < a.bsplit ? a.B0 + (int64_t)n * a.ldb : a.B1 + (int64_t)(n - a.bsplit) * a.ldb;
                xb[i] = n < a.N ? __ldg(reinterpret_cast<const float4*>(brow + k0) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            float ta[kSlots][4], tb[kSlots][4], tm[4], tr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + 4 * warp + j;
                const bool kv = k < a.K;
                tm[j] = (kv && a.mean != nullptr) ? __ldg(a.mean + k) : 0.f;
                tr[j] = (kv && a.rstd != nullptr) ? __ldg(a.rstd + k) : 1.f;
#pragma unroll
                for (int i = 0; i < kSlots; ++i) {
                    const int m = m0 + lane + 32 * i, n = n0 + lane + 32 * i;
                    ta[i][j] = (kv && m < a.M) ? __ldg(a.A + (int64_t)k * a.lda + m) : 0.f;
                    tb[i][j] = (kv && n < a.N) ? __ldg(a.B0 + (int64_t)k * a.ldb + n) : ((kv && n == a.ones_col) ? 1.f : 0.f);
                }
            }
#pragma unroll
            for (int i = 0; i < kSlots; ++i) {
                xa[i] = make_float4(ta[i][0], ta[i][1], ta[i][2], ta[i][3]);
                xb[i] = make_float4(tb[i][0], tb[i][1], tb[i][2], tb[i][3]);
            }
            km = make_float4(tm[0], tm[1], tm[2], tm[3]);
            kr = make_float4(tr[0], tr[1], tr[2], tr[3]);
        }
    };
    auto lo_of = [](float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); };       // v - (what kind::tf32 reads of v)
    auto store = [&](unsigned char* st, const float4 (&xa)[kSlots], const float4 (&xb)[kSlots], const float4& km, const float4& kr) {
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
            const int off = slot_c(i) * kTfPitch + slot_r(i) * 16;
            float4 va = xa[i], vb = xb[i];
            if (kMode == 0) va = make_float4((va.x - mu[i]) * rs[i], (va.y - mu[i]) * rs[i], (va.z - mu[i]) * rs[i], (va.w - mu[i]) * rs[i]);
            else if (a.mean != nullptr) vb = make_float4((vb.x - km.x) * kr.x, (vb.y - km.y) * kr.y, (vb.z - km.z) * kr.z, (vb.w - km.w) * kr.w);
            *reinterpret_cast<float4*>(st + off) = va;
            *reinterpret_cast<float4*>(st + kTfTile + off) = make_float4(lo_of(va.x), lo_of(va.y), lo_of(va.z), lo_of(va.w));
            *reinterpret_cast<float4*>(st + 2 * kTfTile + off) = vb;
            *reinterpret_cast<float4*>(st + 3 * kTfTile + off) = make_float4(lo_of(vb.x), lo_of(vb.y), lo_of(vb.z), lo_of(vb.w));
        }
    };

    if (warp < kTfProducers / 32) {
        // ---- producers: step s lives in shared-memory stage s % 3 and register set s % 2 ------------------------------------------
        float4 km[kTfAhead], kr[kTfAhead];
#pragma unroll
        for (int d = 0; d < kTfAhead; ++d)
            if (d < steps) load(kbase + d * kTfBK, ra[d], rb[d], km[d], kr[d]);
        auto body = [&](int s, float4 (&xa)[kSlots], float4 (&xb)[kSlots], float4& xm, float4& xr) {
            const int sg = s % kTfStages;
            // the stage is free once the MMAs of step s - 3 have completed: warp 0 polls their commit, the producers meet at a barrier
            if (s >= kTfStages) {
                if (warp == 0) mbar_wait(&hdr.empty[sg], ((s - kTfStages) / kTfStages) & 1u, 8u + sg);
                named_bar_sync(1, kTfProducers);
            }
            store(stage0 + sg * kTfStage, xa, xb, xm, xr);
            if (s + kTfAhead < steps) load(kbase + (s + kTfAhead) * kTfBK, xa, xb, xm, xr);      // in flight across the next iterations
            fence_proxy_async();                                    // generic-proxy stores -> visible to the tensor core
            named_bar_sync(1, kTfProducers);
            if (tid == 0) mbar_arrive(&hdr.full[sg]);
        };
        for (int s = 0; s < steps; s += kTfAhead) {
#pragma unroll
            for (int d = 0; d < kTfAhead; ++d)
                if (s + d < steps) body(s + d, ra[d], rb[d], km[d], kr[d]);
        }
    } else {
      if (lane == 0) {
        // ---- consumer: one thread issues every MMA -------------------------------------------------------------------------------
        for (int s = 0; s < steps; ++s) {
            const int sg = s % kTfStages;
            mbar_wait(&hdr.full[sg], (s / kTfStages) & 1u, 4u + sg);
            tc_fence_after();
            const uint32_t base = smem_u32(stage0 + sg * kTfStage);
#pragma unroll
            for (int j = 0; j < kTfBK / 8; ++j) {               // K = 8 per MMA = two chunks
                const uint64_t a_hi = make_desc(base + 2 * j * kTfPitch, kTfPitch, 128u);
                const uint64_t a_lo = make_desc(base + kTfTile + 2 * j * kTfPitch, kTfPitch, 128u);
                const uint64_t b_hi = make_desc(base + 2 * kTfTile + 2 * j * kTfPitch, kTfPitch, 128u);
                const uint64_t b_lo = make_desc(base + 3 * kTfTile + 2 * j * kTfPitch, kTfPitch, 128u);
                umma_tf32(tmem, a_lo, b_hi, idesc, (s > 0 || j > 0) ? 1u : 0u);        // small terms first
                umma_tf32(tmem, a_hi, b_lo, idesc, 1u);
                umma_tf32(tmem, a_hi, b_hi, idesc, 1u);
            }
            umma_commit(&hdr.empty[sg]);
        }
        // commits complete in issue order: the last one implies every MMA has completed
        const int last = steps - 1;
        mbar_wait(&hdr.empty[last % kTfStages], (last / kTfStages) & 1u, 12u);
      }
      __syncwarp();
    }
    __syncthreads();
    tc_fence_after();

    // ---- epilogue: TMEM lane = row of C, 128 columns; 32 columns at a time through shared memory ----------------------------------
    constexpr int OPITCH = 32 * 4 + 16;                     // 36 words: a quarter-warp's 16-byte stores are conflict-free
    float* ost = reinterpret_cast<float*>(stage0);
    const int quad = warp & 3, half = (warp >> 2) & 1, row = quad * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(16 * half);
#pragma unroll 1
    for (int q = 0; q < kTfBN / 32 && warp < kTfProducers / 32; ++q) {
        float v[16];
        tmem_ld16(taddr + 32 * q, v);
        if (a.bias0 != nullptr) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int n = n0 + 32 * q + 16 * half + e;
                if (n < a.N) v[e] += n < a.biassplit ? __ldg(a.bias0 + n) : __ldg(a.bias1 + (n - a.biassplit));
            }
        }
        unsigned char* mine = reinterpret_cast<unsigned char*>(ost) + row * OPITCH + half * 64;
#pragma unroll
        for (int e = 0; e < 4; ++e) reinterpret_cast<float4*>(mine)[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        named_bar_sync(1, kTfProducers);
#pragma unroll
        for (int k = 0; k < kTfBM * 8 / kTfProducers; ++k) {
            const int idx = tid + k * kTfProducers, r = idx >> 3, piece = idx & 7;
            const int m = m0 + r, n = n0 + 32 * q + 4 * piece;
            if (m < a.M && n < a.store_n)        // a multiple of 4 (checked by the caller)
                *reinterpret_cast<float4*>(a.C + (int64_t)slice * a.split_stride + (int64_t)m * a.ldc + n) = *reinterpret_cast<const float4*>(reinterpret_cast<unsigned char*>(ost) + r * OPITCH + piece * 16);
        }
        named_bar_sync(1, kTfProducers);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

__global__ void __launch_bounds__(kTfThreads, 1) gemm_tf32x3_kernel(const TfArgs a) { gemm_tf32x3_tile<0>(a, blockIdx.x, blockIdx.y, blockIdx.z); }

// batch-major form, a group of products in one launch: blockIdx.x = a 128-column tile of one of them, blockIdx.z = K slice
__global__ void __launch_bounds__(kTfThreads, 1) gemm_tf32x3_group_kernel(const TfGroup g) {
    int pi = 0;
    while (pi + 1 < g.nprob && (int)blockIdx.x >= g.tile0[pi + 1]) ++pi;
    // the chosen entry goes to registers: a dynamically indexed kernel parameter would be re-read from the constant bank after every barrier
    TfArgs a;
    a.A = g.pr[pi].A; a.lda = g.pr[pi].lda;
    a.B0 = g.pr[pi].B0; a.B1 = nullptr; a.bsplit = 1 << 30; a.ldb = g.pr[pi].ldb;
    a.bias0 = nullptr; a.bias1 = nullptr; a.biassplit = 1 << 30;
    a.C = g.pr[pi].C; a.ldc = g.pr[pi].ldc;
    a.M = g.pr[pi].M; a.N = g.pr[pi].N; a.K = g.pr[pi].K;
    a.mean = g.pr[pi].mean; a.rstd = g.pr[pi].rstd;
    a.splits = g.pr[pi].splits; a.split_stride = g.pr[pi].split_stride;
    a.ones_col = g.pr[pi].ones_col; a.store_n = g.pr[pi].store_n;
    gemm_tf32x3_tile<1>(a, blockIdx.x - g.tile0[pi], 0, blockIdx.z);
}

}  // namespace

// Whether gemm_tf32x3_tn can run this product (both operands K-major with 16-byte aligned rows, K a multiple of 32, C rows 16-byte aligned).
bool gemm_tf32x3_ok(const float* A, int lda, const float* B0, const float* B1, int ldb, const float* C, int ldc, int M, int N, int K) {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return M >= 1 && N >= 4 && K >= kTfBK && K % kTfBK == 0 && N % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 && al(A) && al(B0) &&
           (B1 == nullptr || al(B1)) && al(C);
}

constexpr size_t kTfSmem = 128 + kTfStages * (size_t)kTfStage + 128;
static bool tf_aligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int gemm_tf32x3_launch(const TfArgs& a0, cudaStream_t st) {
    TfArgs a = a0;
    a.ones_col = -1; a.store_n = a.N;
    HB_REQUIRE(gemm_tf32x3_ok(a.A, a.lda, a.B0, a.B1, a.ldb, a.C, a.ldc, a.M, a.N, a.K), "gemm_tf32x3: unsupported shape or alignment");
    HB_REQUIRE(a.splits >= 1 && a.splits <= a.K / kTfBK && (a.splits == 1 || (a.bias0 == nullptr && a.split_stride % 4 == 0)),
               "gemm_tf32x3: bad K split");
    static bool configured = false;
    if (!configured) {
        HB_CUDA_OK(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTfSmem));
        configured = true;
    }
    gemm_tf32x3_kernel<<<dim3(ceil_div(a.N, kTfBN), ceil_div(a.M, kTfBM), a.splits), kTfThreads, kTfSmem, st>>>(a);
    HB_LAUNCHED();
    return HB_OK;
}

int gemm_tf32x3_launch_group(TfGroup& g, cudaStream_t st) {
    HB_REQUIRE(g.nprob >= 1 && g.nprob <= kTfGroupMax, "gemm_tf32x3 (group): bad problem count");
    int tiles = 0;
    for (int i = 0; i < g.nprob; ++i) {
        TfArgs& a = g.pr[i];
        HB_REQUIRE(a.M >= 1 && a.M <= kTfBM && a.N >= 0 && (a.N >= 1 || a.ones_col >= 0) && (a.ones_col < 0 || a.mean == nullptr) && a.K >= 1 && a.K == g.pr[0].K && a.splits == g.pr[0].splits && a.ldc % 4 == 0 &&
                       a.store_n % 4 == 0 && tf_aligned(a.C) && a.bias0 == nullptr && a.split_stride % 4 == 0 && a.splits >= 1 &&
                       a.splits <= ceil_div(a.K, kTfBK),
                   "gemm_tf32x3 (group): unsupported shape, alignment or K split");
        g.tile0[i] = tiles;
        tiles += ceil_div(std::max(a.N, a.ones_col + 1), kTfBN);
    }
    static bool configured = false;
    if (!configured) {
        HB_CUDA_OK(cudaFuncSetAttribute(gemm_tf32x3_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTfSmem));
        configured = true;
    }
    gemm_tf32x3_group_kernel<<<dim3(tiles, 1, g.pr[0].splits), kTfThreads, kTfSmem, st>>>(g);
    HB_LAUNCHED();
    return HB_OK;
}

// C[M, N] = A[M, K] B[N, K]^T + bias (fp32 in and out, 3 x TF32 on tcgen05).  B rows / bias entries at or past `bsplit` come from B1 / bias1.
int gemm_tf32x3_tn(const float* A, int lda, const float* B0, const float* B1, int bsplit, int ldb, const float* bias0, const float* bias1,
                   int biassplit, float* C, int ldc, int M, int N, int K, cudaStream_t st) {
    TfArgs a;
    a.A = A; a.lda = lda;
    a.B0 = B0; a.B1 = B1 ? B1 : B0; a.bsplit = B1 ? bsplit : (1 << 30); a.ldb = ldb;
    a.bias0 = bias0; a.bias1 = bias1 ? bias1 : bias0; a.biassplit = bias1 ? biassplit : (1 << 30);
    a.C = C; a.ldc = ldc;
    a.M = M; a.N = N; a.K = K;
    a.mean = nullptr; a.rstd = nullptr;
    a.splits = 1; a.split_stride = 0;
    a.ones_col = -1; a.store_n = N;
    return gemm_tf32x3_launch(a, st);
}

int gemm_tf32_check_timeout() {
    unsigned int flag = 0;
    HB_CUDA_OK(cudaMemcpyFromSymbol(&flag, g_tc_timeout, sizeof(flag)));
    HB_REQUIRE(flag == 0, "tcgen05 classifier GEMM: an mbarrier wait timed out (pipeline bug; barrier code %u)", flag);
    return HB_OK;
}

}  // namespace hb

// y = x W^T + b with fp32 accuracy on the tensor cores (three TF32 passes): x f32 [M][K] (row stride lda), W f32 [N][K] (row stride ldb),
// bias f32 [N] or NULL, y f32 [M][N] (row stride ldc).  K a multiple of 32; N, lda, ldb, ldc multiples of 4; 16-byte aligned pointers.
// The product behind torch.nn.Linear in the classifier's gated MLPs (reference wakeword.py:334-348).
extern "C" int hb_linear_tf32x3(const float* x_dev, int lda, const float* w_dev, int ldb, const float* bias_dev, float* y_dev, int ldc, int M,
                                int N, int K, void* stream) {
    HB_REQUIRE(x_dev && w_dev && y_dev && M >= 0, "hb_linear_tf32x3: bad argument");
    if (M == 0) return HB_OK;
    return hb::gemm_tf32x3_tn(x_dev, lda, w_dev, nullptr, 0, ldb, bias_dev, nullptr, 0, y_dev, ldc, M, N, K, (cudaStream_t)stream);
}
