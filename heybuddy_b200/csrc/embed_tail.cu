// K7, the tail of the speech-embedding stack: conv2d_15's 2x2 pool (both time phases) + conv2d_16 .. conv2d_19 (reference
// embeddings.py:32-42, layers after the last pool).
//
// After the pool a clip is 13 rows x 2 bins x 96 channels per phase and conv2d_16 (1x2 VALID) collapses the bins: the tail is four
// small 1-D convolutions over (clip, phase, row) "columns" with 96 x {192, 288, 96, 288} weights.  Round 1 ran it on the generic
// block kernel: 6 clips per CTA, every layer's weights (up to 78 KB) fetched again by each of the 1366 tiles of each phase, and a
// SAME-padding column layout in which two of three accumulator columns are padding -- 0.33 ms per 8192 clips at 22 % tensor-pipe
// activity, 630 MB of L2 -> shared-memory weight traffic.  Here:
//   * one launch per LAYER over all columns of both phases; a persistent CTA keeps the layer's weights in shared memory and strides
//     over tiles of 128 columns, so the weights are read once per CTA and the activations (192 B per column) stream through HBM;
//   * positions are the M dimension (128 per MMA), the 96 output channels are N: the accumulator's lanes are positions, so an
//     epilogue thread owns one position's channel vector -- bias, LeakyReLU, fp16 pack and two 16-byte stores, no transposition;
//   * a 3x1 VALID conv is three shifted views of the same [chunk][column][8] operand planes (start address + 16 B per tap); outputs
//     that mix rows of neighbouring clips land in rows >= R - 2 (R - 4 after two such layers), which nothing valid reads.
// conv2d_16's staging does the 2x2 max-pool with the time phase straight from block 4's output.  Same rounding points as before:
// fp16 operands, fp32 accumulation, fp16 activations between layers, f32 result.
#include "tc_ptx.cuh"

#include <algorithm>
#include <vector>

namespace hb {

namespace {

constexpr int kTailC = 96, kTailChunks = 12, kTailCols = 128, kTailThreads = 256;
constexpr int kTailMaxDup = 4;                  // slots one (phase, row) may feed when conv2d_19 writes the slots itself
// bytes per operand plane: 128 columns + 2 taps of overhang + pad.  137 records: consecutive planes start 16 B apart modulo 128 B, so
// the staging stores of a quarter-warp (one column, eight chunks = eight planes) fall into eight different bank groups
constexpr int kTailPitch = 137 * 16;

struct TailWeights {
    unsigned char* w[4] = {nullptr, nullptr, nullptr, nullptr};     // B operands [K chunk][96][8] fp16 of conv2d_16 .. 19
    float* bias[4] = {nullptr, nullptr, nullptr, nullptr};
};

struct TailArgs {
    const void* in;        // POOL_IN: block 4's output, fp16 [B][12][T15][4][8]; else fp16 [n_cols][96]
    __half* out;           // fp16 [n_cols][96] (not LAST)
    float* out0;           // LAST: f32 [B][T_out][96] of pool phase 0 / 1
    float* out1;
    const unsigned char* w;
    const float* bias;
    // LAST, optional: write the caller's slots directly -- slot_map i32 [2 R][kTailMaxDup] lists the slots fed by (phase, row), -1 = none;
    // out_slots f32 [B][n_slots][96]
    const int32_t* slot_map;
    float* out_slots;
    int n_slots;
    int B, T15, R, T_out, n_cols, n_tiles;
};

struct TailHeader {
    uint64_t bar, wbar;
    uint32_t tmem;
    uint32_t pad;
};

// column <-> (clip, pool phase, row): both phases of a clip are neighbours, so the second read of block 4's rows hits L1 / L2
__device__ __forceinline__ void tail_decode(int col, int R, int& clip, int& ph, int& row) {
    const int cp = col / R;
    row = col - cp * R;
    clip = cp >> 1;
    ph = cp & 1;
}

// KC K chunks of 8 input channels, TAPS row taps (KC = 12 TAPS, or 24 = the two pooled bins of conv2d_16).
// 256 threads: everyone stages (all of a thread's loads in flight before its first store; the plain layers fetch the NEXT tile's
// records while the current tile's MMAs and epilogue run), thread 0 issues the MMAs, warp w drains TMEM lane quadrant w % 4,
// channel half w / 4.
template <int KC, int TAPS, bool POOL_IN, bool LAST>
__global__ void __launch_bounds__(kTailThreads, 2) tail_layer_kernel(const TailArgs a) {
    // NREC records of 8 channels per input column = operand planes (conv2d_16 reads the two pooled bins = 24, the others 12)
    constexpr int NREC = POOL_IN ? 24 : KC / TAPS;
    static_assert(KC % 2 == 0 && NREC % 2 == 0 && (NREC == kTailChunks || (NREC == 2 * kTailChunks && TAPS == 1)), "K layout");
    constexpr int NPL = NREC, W_BYTES = KC * kTailC * 16, ACT_BYTES = NPL * kTailPitch;
    extern __shared__ __align__(128) unsigned char smem[];
    TailHeader& hdr = *reinterpret_cast<TailHeader*>(smem);
    unsigned char* wbuf = smem + 128;
    unsigned char* act = wbuf + W_BYTES;
    float* bias_s = reinterpret_cast<float*>(act + ACT_BYTES);
    int64_t* row_off = reinterpret_cast<int64_t*>(bias_s + kTailC);     // per tile column: POOL_IN source / LAST destination offset, -1 = none
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TC_STAMP(0);
    if (tid == 0) {
        mbar_init(&hdr.bar, 1);
        mbar_init(&hdr.wbar, 1);
        fence_barrier_init();
        mbar_expect_tx(&hdr.wbar, (uint32_t)W_BYTES);               // the layer's weights: one bulk copy, under the rest of the set-up
        bulk_g2s(wbuf, a.w, (uint32_t)W_BYTES, &hdr.wbar);
    }
    if (warp == 0) tmem_alloc(&hdr.tmem, 128);
    for (int i = tid; i < kTailC; i += kTailThreads) bias_s[i] = a.bias[i];
    for (int i = tid; i < ACT_BYTES / 16; i += kTailThreads) reinterpret_cast<uint4*>(act)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = hdr.tmem, act_u32 = smem_u32(act), w_u32 = smem_u32(wbuf);
    constexpr uint32_t idesc = make_idesc(128, kTailC);
    TC_STAMP(1);
    // plain layers: records (column i, chunk c) of a tile are kRec consecutive uint4 of the [n_cols][NREC] input
    constexpr int kRec = (kTailCols + TAPS - 1) * NREC, kPer = (kRec + kTailThreads - 1) / kTailThreads;
    uint4 rec[POOL_IN ? 1 : kPer];
    auto fetch = [&](int tile) {
        if (POOL_IN || tile >= a.n_tiles) return;
        const int c0 = tile * kTailCols;
        const uint4* in = reinterpret_cast<const uint4*>(a.in) + (int64_t)c0 * NREC;
        const int64_t limit = ((int64_t)a.n_cols - c0) * NREC;                     // records that exist past c0
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int r = tid + k * kTailThreads;
            rec[k] = (r < kRec && r < limit) ? __ldg(in + r) : make_uint4(0, 0, 0, 0);
        }
    };
    fetch(blockIdx.x);
    uint32_t phase = 0;
    int n_done = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int c0 = tile * kTailCols;
        if ((POOL_IN || LAST) && tid < kTailCols) {
            // one (clip, phase, row) decode per column and tile (two integer divisions), not one per record
            const int col = c0 + tid;
            int64_t off = -1;
            if (col < a.n_cols) {
                int clip, ph, row;
                tail_decode(col, a.R, clip, ph, row);
                if (POOL_IN) {
                    const int rr = 2 * row + ph;
                    if (rr + 1 < a.T15) off = ((int64_t)clip * kTailChunks * a.T15 + rr) * 4;        // uint4 index of (clip, chunk 0, row rr, bin 0)
                } else if (a.slot_map != nullptr) {
                    off = ((int64_t)clip << 32) | (uint32_t)(ph * a.R + row);                         // clip, slot-map row
                } else if (row < a.T_out) {
                    off = (((int64_t)clip * a.T_out + row) * kTailC) * 2 + ph;                         // float offset * 2 + phase
                }
            }
            row_off[tid] = off;
        }
        if (POOL_IN) __syncthreads();
        // ---- stage the tile's columns as K-major operand planes [chunk][column][8] ---------------------------------------------
        if (POOL_IN) {
            // (column i, chunk c): 2 rows x 4 bins of block 4's output -> the two pooled bins (K chunks c and 12 + c)
            const uint4* in = reinterpret_cast<const uint4*>(a.in);
            constexpr int kItems = kTailCols * kTailChunks / kTailThreads, kBatch = 3;
            static_assert(kItems % kBatch == 0, "pool staging batches");
#pragma unroll 1
            for (int b0 = 0; b0 < kItems; b0 += kBatch) {
                uint4 x[kBatch][8];
                bool ok[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int r = tid + (b0 + u) * kTailThreads;
                    const int i = r / kTailChunks, c = r - i * kTailChunks;
                    const int64_t off = row_off[i];
                    ok[u] = off >= 0;
                    const uint4* src = in + off + (int64_t)c * a.T15 * 4;
#pragma unroll
                    for (int q = 0; q < 8; ++q) x[u][q] = ok[u] ? __ldg(src + q) : make_uint4(0, 0, 0, 0);     // rows rr, rr + 1: four bins each
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int r = tid + (b0 + u) * kTailThreads;
                    const int i = r / kTailChunks, c = r - i * kTailChunks;
                    const __half2* h = reinterpret_cast<const __half2*>(x[u]);
                    __half2 m0[4], m1[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        m0[e] = __hmax2_nan(__hmax2_nan(h[0 * 4 + e], h[1 * 4 + e]), __hmax2_nan(h[4 * 4 + e], h[5 * 4 + e]));
                        m1[e] = __hmax2_nan(__hmax2_nan(h[2 * 4 + e], h[3 * 4 + e]), __hmax2_nan(h[6 * 4 + e], h[7 * 4 + e]));
                    }
                    *reinterpret_cast<uint4*>(act + c * kTailPitch + i * 16) = *reinterpret_cast<uint4*>(m0);
                    *reinterpret_cast<uint4*>(act + (kTailChunks + c) * kTailPitch + i * 16) = *reinterpret_cast<uint4*>(m1);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int r = tid + k * kTailThreads;
                if (r < kRec) {
                    const int i = r / NREC, c = r - i * NREC;
                    *reinterpret_cast<uint4*>(act + c * kTailPitch + i * 16) = rec[k];
                }
            }
        }
        fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncthreads();
        if (n_done == 1) TC_STAMP(2);
        // ---- D[128 positions][96] = sum over K steps of A[positions][16] x W[96][16]^T -------------------------------------------
        if (tid == 0) {
            if (n_done == 0) mbar_wait(&hdr.wbar, 0u);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < KC / 2; ++j) {
                const int kk = 2 * j;
                const uint32_t a_addr = act_u32 + (kk % NREC) * kTailPitch + (kk / NREC) * 16;        // plane of the chunk, shifted by the row tap
                umma_f16(tmem, make_desc(a_addr, kTailPitch, 128u), make_desc(w_u32 + kk * (kTailC * 16), kTailC * 16, 128u), idesc, j > 0 ? 1u : 0u);
            }
            umma_commit(&hdr.bar);
        }
        if (n_done == 1) TC_STAMP(3);
        fetch(tile + (int)gridDim.x);        // the next tile's records travel while this tile's MMAs and epilogue run
        mbar_wait(&hdr.bar, phase & 1u);
        ++phase;
        tc_fence_after();
        if (n_done == 1) TC_STAMP(4);
        // ---- epilogue: lane = position, columns = channels -----------------------------------------------------------------------
        // A thread owns one position and 16 of every 32 channels.  Its values go through shared memory (the operand planes are free
        // once the MMAs have completed) so that the global stores are whole 64 / 128-byte row segments instead of one 16-byte piece
        // per lane at a 192-byte stride (which made the epilogue 5.3 k of a tile's 10 k cycles).
        constexpr int ELT = LAST ? 4 : 2, ROWB = 32 * ELT, OPITCH = ROWB + 16, PIECES = ROWB / 16;
        const int quad = warp & 3, half = warp >> 2, pos = quad * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(16 * half);
#pragma unroll
        for (int q = 0; q < kTailC / 32; ++q) {
            float v[16];
            tmem_ld16(taddr + 32 * q, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += bias_s[32 * q + 16 * half + e];
            unsigned char* mine = act + pos * OPITCH + half * (16 * ELT);
            if (LAST) {
#pragma unroll
                for (int e = 0; e < 4; ++e) reinterpret_cast<float4*>(mine)[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
            } else {
                uint32_t h[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h[e] = leaky_half2(v[2 * e], v[2 * e + 1]);
                reinterpret_cast<uint4*>(mine)[0] = make_uint4(h[0], h[1], h[2], h[3]);
                reinterpret_cast<uint4*>(mine)[1] = make_uint4(h[4], h[5], h[6], h[7]);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kTailCols * PIECES / kTailThreads; ++k) {
                const int idx = tid + k * kTailThreads, row = idx / PIECES, piece = idx - row * PIECES;
                const int col = c0 + row;
                if (col < a.n_cols) {
                    const uint4 val = *reinterpret_cast<const uint4*>(act + row * OPITCH + piece * 16);
                    if (LAST) {
                        const int64_t off = row_off[row];
                        if (off >= 0 && a.slot_map != nullptr) {
                            const int32_t* slots = a.slot_map + (int64_t)(uint32_t)off * kTailMaxDup;
                            float* dst = a.out_slots + (off >> 32) * a.n_slots * kTailC + 32 * q + 4 * piece;
#pragma unroll
                            for (int d = 0; d < kTailMaxDup; ++d) {
                                const int sl = __ldg(slots + d);
                                if (sl >= 0) *reinterpret_cast<uint4*>(dst + (int64_t)sl * kTailC) = val;
                            }
                        } else if (off >= 0) {
                            *reinterpret_cast<uint4*>(((off & 1) ? a.out1 : a.out0) + (off >> 1) + 32 * q + 4 * piece) = val;
                        }
                    } else {
                        reinterpret_cast<uint4*>(a.out)[(int64_t)col * kTailChunks + 4 * q + piece] = val;
                    }
                }
            }
            __syncthreads();
        }
        tc_fence_before();
        __syncthreads();                     // accumulator read, operand planes free for the next tile
        if (n_done == 1) TC_STAMP(5);
        if (n_done == 0) TC_STAMP(9);
        ++n_done;
    }
    TC_STAMP(6);
    if (warp == 0) tmem_dealloc(tmem, 128);
    TC_STAMP(7);
    if (tid == 0 && blockIdx.x < 8) g_tc_times[blockIdx.x][8] = n_done;
}

template <int KC, int TAPS, bool POOL_IN, bool LAST>
size_t tail_smem_bytes() {
    return 128 + (size_t)KC * kTailC * 16 + (size_t)(POOL_IN ? 24 : KC / TAPS) * kTailPitch + kTailC * sizeof(float) + kTailCols * sizeof(int64_t) + 128;
}

template <int KC, int TAPS, bool POOL_IN, bool LAST>
int tail_launch(TailArgs a, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        HB_CUDA_OK(cudaFuncSetAttribute(tail_layer_kernel<KC, TAPS, POOL_IN, LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)tail_smem_bytes<KC, TAPS, POOL_IN, LAST>()));
        configured = true;
    }
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        HB_CUDA_OK(cudaGetDevice(&dev));
        HB_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid = std::min(a.n_tiles, 2 * n_sm);
    tail_layer_kernel<KC, TAPS, POOL_IN, LAST><<<grid, kTailThreads, tail_smem_bytes<KC, TAPS, POOL_IN, LAST>(), st>>>(a);
    HB_LAUNCHED();
    return HB_OK;
}

// fp16 [n_cols][96] (columns = (clip, phase, row)) -> f32 [B][T][96] of pool phase 0 (parity hook)
__global__ void tail_dump_kernel(const __half* __restrict__ act, float* __restrict__ out, int B, int R, int T) {
    const int64_t total = (int64_t)B * T * kTailC;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % kTailC);
        const int64_t r = i / kTailC;
        const int t = (int)(r % T), b = (int)(r / T);
        out[i] = __half2float(act[((int64_t)(2 * b) * R + t) * kTailC + c]);
    }
}

}  // namespace

int tail_prepare(hb_embed_model* m, const float* weights_host) {
    TailWeights* tw = new TailWeights();
    m->tail = tw;
    for (int l = 0; l < 4; ++l) {
        const int li = 16 + l;
        const ConvLayer& L = kLayers[li];
        HB_REQUIRE(L.cin == kTailC && L.cout == kTailC && L.same == 0 && L.pool_t == 1 && L.pool_f == 1, "tail: unexpected layer table entry for conv2d_%d", li);
        const int taps = L.kh * L.kw, kc = taps * kTailChunks;
        std::vector<__half> packed((size_t)kc * kTailC * 8);
        const float* w = weights_host + m->w_off[li];          // [tap][cin][cout]
        for (int kk = 0; kk < kc; ++kk) {
            const int tap = kk / kTailChunks, c = kk % kTailChunks;
            for (int n = 0; n < kTailC; ++n)
                for (int e = 0; e < 8; ++e)
                    packed[((size_t)kk * kTailC + n) * 8 + e] = __float2half_rn(w[((int64_t)tap * kTailC + c * 8 + e) * kTailC + n]);
        }
        HB_CUDA_OK(cudaMalloc(&tw->w[l], packed.size() * 2));
        HB_CUDA_OK(cudaMemcpy(tw->w[l], packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
        HB_CUDA_OK(cudaMalloc(&tw->bias[l], kTailC * sizeof(float)));
        HB_CUDA_OK(cudaMemcpy(tw->bias[l], weights_host + m->b_off[li], kTailC * sizeof(float), cudaMemcpyHostToDevice));
    }
    return HB_OK;
}

void tail_release(hb_embed_model* m) {
    TailWeights* tw = reinterpret_cast<TailWeights*>(m->tail);
    if (!tw) return;
    for (int l = 0; l < 4; ++l) {
        cudaFree(tw->w[l]);
        cudaFree(tw->bias[l]);
    }
    delete tw;
    m->tail = nullptr;
}

int64_t tail_scratch_bytes(int B, int T15) {
    const int64_t n_cols = 2ll * B * std::max(1, T15 / 2);
    return 2 * ((n_cols * kTailC * 2 + 255) & ~255ll);
}

// block 4's output -> conv2d_19's output for both pool phases, f32 [B][T15 / 2 - 4][96] each.  `pooled`: block 4 already wrote the
// pooled operand format fp16 [(clip, phase, row)][24][8] (tcg_block4's pool2_out, one tile per clip); else fp16 [B][12][T15][4][8] and
// conv2d_16's staging pools (long strips, parity hook).
// upto: 16 .. 19 = stop after that conv (parity hook; dbg_out then receives the phase-0 activation f32 [B][T][96]).
// slot_map_dev / out_slots / n_slots (optional, upto = 19): conv2d_19 writes out_slots f32 [B][n_slots][96] itself (out0 / out1 unused).
int tail_run(const hb_embed_model* m, const __half* block4_out, bool pooled, int B, int T15, float* out0, float* out1, void* scratch,
             int64_t scratch_bytes, int upto, float* dbg_out, cudaStream_t st, const int32_t* slot_map_dev, float* out_slots, int n_slots) {
    const TailWeights* tw = reinterpret_cast<const TailWeights*>(m->tail);
    HB_REQUIRE(tw != nullptr, "tail weights missing");
    const int R = T15 / 2;
    HB_REQUIRE(R >= 5, "hb_embed: strip too short for the tail (%d rows after conv2d_15)", T15);
    HB_REQUIRE(scratch_bytes >= tail_scratch_bytes(B, T15), "hb_embed: tail scratch too small");
    const int64_t n_cols64 = 2ll * B * R;
    HB_REQUIRE(n_cols64 + kTailCols < (1ll << 31), "hb_embed: too many tail columns");
    TailArgs a;
    a.B = B; a.T15 = T15; a.R = R; a.T_out = R - 4;
    a.n_cols = (int)n_cols64;
    a.n_tiles = ceil_div(a.n_cols, kTailCols);
    a.out0 = out0; a.out1 = out1;
    a.slot_map = upto == 19 ? slot_map_dev : nullptr;
    a.out_slots = out_slots;
    a.n_slots = n_slots;
    __half* x = reinterpret_cast<__half*>(scratch);
    __half* y = reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(scratch) + tail_scratch_bytes(B, T15) / 2);
    auto dump = [&](const __half* act, int T) {
        const int64_t n = (int64_t)B * T * kTailC;
        tail_dump_kernel<<<(int)std::min<int64_t>(ceil_div64(n, 256), 148 * 8), 256, 0, st>>>(act, dbg_out, B, R, T);
        return cudaGetLastError() == cudaSuccess ? HB_OK : HB_ERR_CUDA;
    };
    int rc;
    a.in = block4_out; a.out = x; a.w = tw->w[0]; a.bias = tw->bias[0];
    if ((rc = pooled ? tail_launch<24, 1, false, false>(a, st) : tail_launch<24, 1, true, false>(a, st))) return rc;
    if (upto == 16) return dump(x, R);
    a.in = x; a.out = y; a.w = tw->w[1]; a.bias = tw->bias[1];
    if ((rc = tail_launch<36, 3, false, false>(a, st))) return rc;
    if (upto == 17) return dump(y, R - 2);
    a.in = y; a.out = x; a.w = tw->w[2]; a.bias = tw->bias[2];
    if ((rc = tail_launch<12, 1, false, false>(a, st))) return rc;
    if (upto == 18) return dump(x, R - 2);
    a.in = x; a.out = nullptr; a.w = tw->w[3]; a.bias = tw->bias[3];
    return tail_launch<36, 3, false, true>(a, st);
}

// profiling aid: phase timestamps of the first 8 CTAs of the last tail launch (0 start, 1 set up, 9 first tile done, 2..5 second tile:
// staged / MMAs issued / accumulator ready / drained, 6 all tiles, 7 deallocated, 8 = tiles this CTA ran)
int tail_debug_times(long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, g_tc_times, sizeof(long long) * 8 * 16) == cudaSuccess ? 0 : -2;
}

int tail_max_dup() { return kTailMaxDup; }

int tail_check_timeout() {
    unsigned int flag = 0;
    HB_CUDA_OK(cudaMemcpyFromSymbol(&flag, g_tc_timeout, sizeof(flag)));
    HB_REQUIRE(flag == 0, "tcgen05 embed kernel (tail): an mbarrier wait timed out (pipeline bug)");
    return HB_OK;
}

}  // namespace hb
