// K7 (product mode): speech-embedding conv stack as tcgen05 implicit GEMMs with TMEM accumulators.
//
// Replaces the ORT run of speech-embedding.onnx behind SpeechEmbeddingModel.__call__ (reference
// src/python/heybuddy/embeddings.py:32-42).  Layer table: embed_common.cuh / spec.py.
//
// Design (DESIGN.md "embedding conv stack"):
//  * Every conv of blocks 1-4 is 1x3 (freq, SAME) or 3x1 (time, VALID).  With activations stored
//    position-major -- position q = 1 + row*(F+1) + f, one shared zero pad column per row -- a tap
//    is a pure shift of the position index (+-1 for freq, +0/S/2S for time, S = F+1), so the
//    implicit GEMM   D[q, cout] = sum_tap sum_cin X[q + shift(tap), cin] * W[tap, cin, cout]
//    needs no im2col: the A operand of tap `t` is the SAME shared-memory buffer addressed through a
//    UMMA descriptor whose start address is moved by shift(t) * 16 bytes.
//  * Shared-memory layout of an activation buffer: [channel chunk of 8][position][8 x fp16]; in
//    UMMA terms a K-major, no-swizzle canonical layout: core matrix = 8 positions x 16 B
//    (contiguous 128 B), SBO = 128 B (next 8 positions), LBO = P_alloc * 16 B (next 8 channels).
//  * One CTA owns a tile (a clip's time slice, rows independent between clips) and runs the four
//    convs of a block back to back with the activations ping-ponging between two shared-memory
//    buffers: M = 128 positions per tcgen05.mma, N = Cout (32/48/80/96; 24 and 72 are zero padded),
//    K = 16 per instruction (kind::f16, fp16 operands, fp32 accumulation in TMEM).
//  * Warp roles: warp 0 = MMA issuer (one elected thread) + TMEM allocator, warp 1 = weight loader
//    (cp.async.bulk of the next layer's pre-packed B operand, double buffered, mbarrier tx-count),
//    warps 2-9 = epilogue (tcgen05.ld -> +bias -> LeakyReLU -> fp16 -> st.shared into the next
//    layer's A buffer).  Four TMEM accumulator slots let MMA of tile i+1.. overlap the epilogue of
//    tile i.  Max-pool + store to global (fp16, chunk-major) closes the block.
//  * conv2d (Cin = 1) is CUDA-core work fused into block 1's prologue; the last 2x2 pool (two
//    phases) and block 5 (1.6 % of the MACs) run on the fp32 CUDA-core kernels of embed_fp32.cu.
#include "embed_common.cuh"

#include <vector>

namespace hb {

// ------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ unsigned int g_tc_timeout = 0;  // set when a barrier wait gave up (never hang the GPU)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware instead of spinning
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {  // ~2 s: give up instead of hanging the box
            atomicExch(&g_tc_timeout, 1u);
            return;
        }
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16, issued by one thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets TMEM lane (base lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (K direction, next 8-element chunk),
//   [32,46) stride byte offset >> 4 (M/N direction, next 8 rows), [46,48) version = 1, [61,64) layout = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bit 4), a/b format F16 (0), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// Block configuration
// ------------------------------------------------------------------------------------------------
constexpr int kTcThreads = 320;   // warp 0 MMA, warp 1 loader, warps 2..9 epilogue
constexpr int kSlots = 4;         // TMEM accumulator slots
constexpr int kSlotCols = 96;     // columns per slot (max N)
constexpr int kTmemCols = 512;
constexpr int kMaxTcLayers = 4;

struct TcLayer {
    int cin_chunks;   // padded Cin / 8
    int n;            // padded Cout
    int is_time;      // taps shift by S (time conv) instead of 1 (freq conv)
    int w_bytes;      // packed B operand bytes (3 taps)
    int64_t w_off;    // byte offset in the packed weight buffer
    int bias_off;     // float offset in the packed bias buffer
};

struct TcBlockArgs {
    const void* in;        // block 1: mel f32 [clips][T_in][32]; else fp16 chunk-major [clips][cin_chunks][T_in][F][8]
    __half* out;           // fp16 chunk-major [clips][out_chunks][T_out][F_out][8]
    const unsigned char* w_packed;
    const float* bias_packed;
    const float* l0_w;     // block 1 only: conv2d kernel f32 [3][24] + bias [24]
    __half* dbg;           // optional activation dump [tiles][chunks][P_alloc][8]
    int dbg_layer;         // tc layer index within the block to dump (-1: none; 100: the block-1 conv2d output)
    int n_clips, T_in, T_out;
    int F, S;              // input freq bins, S = F + 1
    int pool_t, pool_f;
    int tiles_per_clip;    // time tiles per clip
    int rows_out;          // pre-pool output rows per tile (multiple of pool_t)
    int Tt;                // input rows per tile = rows_out + 4 (block 1: + 4 as well: two time convs)
    int n_mt, P_alloc;     // M tiles per layer, allocated positions per chunk
    int n_layers;
    int ch_alloc;          // chunks allocated per activation buffer
    int w_buf_bytes;       // bytes of one weight buffer
    TcLayer layers[kMaxTcLayers];
};

struct TcSmemHeader {
    uint64_t tmem_full[kSlots];
    uint64_t tmem_empty[kSlots];
    uint64_t wbar[2];
    uint32_t tmem_base;
    uint32_t pad[3];
    float bias[kMaxTcLayers * 96];
    float l0[3 * 24 + 24 + 8];
};

__device__ __forceinline__ float leaky(float v) { return v > 0.f ? v : kLeaky * v; }

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <bool kFirstBlock>
__global__ void __launch_bounds__(kTcThreads, 1) tc_block_kernel(const TcBlockArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    TcSmemHeader& hdr = *reinterpret_cast<TcSmemHeader*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P_alloc = a.P_alloc;
    const uint32_t chunk_stride = (uint32_t)P_alloc * 16u;               // bytes between channel chunks
    unsigned char* wbuf0 = smem + ((sizeof(TcSmemHeader) + 127) & ~127);
    unsigned char* wbuf1 = wbuf0 + a.w_buf_bytes;
    unsigned char* act0 = wbuf1 + a.w_buf_bytes + 128;                    // +128: guard for the position -1 read
    unsigned char* act1 = act0 + (size_t)a.ch_alloc * chunk_stride + 128;
    float* mel_tile = reinterpret_cast<float*>(act1 + (size_t)a.ch_alloc * chunk_stride);  // block 1 only

    const int clip = blockIdx.x / a.tiles_per_clip;
    const int tile = blockIdx.x - clip * a.tiles_per_clip;
    const int row0_out = tile * a.rows_out;        // first pre-pool output row of this tile (block-local time)
    const int row0_in = row0_out;                  // convs are top aligned: output row r reads input rows r..r+4
    const int S = a.S, F = a.F, Tt = a.Tt;
    const int P = 1 + Tt * S;

    // ---- one-time setup ----------------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(&hdr.tmem_full[i], 1);
            mbar_init(&hdr.tmem_empty[i], 4);
        }
        mbar_init(&hdr.wbar[0], 1);
        mbar_init(&hdr.wbar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&hdr.tmem_base, kTmemCols);
    for (int i = tid; i < a.n_layers * 96; i += kTcThreads) {
        const int l = i / 96, c = i - l * 96;
        hdr.bias[i] = (c < a.layers[l].n) ? a.bias_packed[a.layers[l].bias_off + c] : 0.f;
    }
    if (kFirstBlock)
        for (int i = tid; i < 3 * 24 + 24; i += kTcThreads) hdr.l0[i] = a.l0_w[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr.tmem_base;

    // first layer's weights start streaming while the input tile is staged
    if (warp == 1 && lane == 0) {
        mbar_expect_tx(&hdr.wbar[0], (uint32_t)a.layers[0].w_bytes);
        bulk_g2s(wbuf0, a.w_packed + a.layers[0].w_off, (uint32_t)a.layers[0].w_bytes, &hdr.wbar[0]);
    }

    // ---- stage the input tile into act0 ------------------------------------------------------------------
    if (kFirstBlock) {
        // mel rows [row0_in, row0_in + Tt) -> shared (zero beyond the clip), then conv2d (Cin = 1) on CUDA cores
        const float* mel = reinterpret_cast<const float*>(a.in) + (int64_t)clip * a.T_in * kMels;
        for (int i = tid; i < Tt * kMels; i += kTcThreads) {
            const int r = i / kMels;
            mel_tile[i] = (row0_in + r < a.T_in) ? __ldg(mel + (int64_t)(row0_in + r) * kMels + (i - r * kMels)) : 0.f;
        }
        __syncthreads();
        for (int q = tid; q < P_alloc; q += kTcThreads) {
            const int r = (q - 1) / S, f = (q - 1) - r * S;
            const bool real = q >= 1 && q < P && f < F;
            float m0 = 0.f, m1 = 0.f, m2 = 0.f;
            if (real) {
                m1 = mel_tile[r * kMels + f];
                m0 = f > 0 ? mel_tile[r * kMels + f - 1] : 0.f;
                m2 = f < F - 1 ? mel_tile[r * kMels + f + 1] : 0.f;
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                uint4 pk = make_uint4(0, 0, 0, 0);
                if (real && ch < 3) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = ch * 8 + j;
                        float acc = fmaf(m0, hdr.l0[c], 0.f);
                        acc = fmaf(m1, hdr.l0[24 + c], acc);
                        acc = fmaf(m2, hdr.l0[48 + c], acc);
                        v[j] = leaky(acc + hdr.l0[72 + c]);
                    }
                    pk = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                }
                *reinterpret_cast<uint4*>(act0 + ch * chunk_stride + (size_t)q * 16) = pk;
            }
        }
    } else {
        const uint4* in = reinterpret_cast<const uint4*>(a.in);
        const int cin_chunks = a.layers[0].cin_chunks;
        for (int i = tid; i < cin_chunks * P_alloc; i += kTcThreads) {
            const int ch = i / P_alloc, q = i - ch * P_alloc;
            const int r = (q - 1) / S, f = (q - 1) - r * S;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (q >= 1 && q < P && f < F && row0_in + r < a.T_in)
                v = __ldg(in + (((int64_t)clip * cin_chunks + ch) * a.T_in + row0_in + r) * F + f);
            *reinterpret_cast<uint4*>(act0 + ch * chunk_stride + (size_t)q * 16) = v;
        }
    }
    if (a.dbg != nullptr && a.dbg_layer == 100) {
        __syncthreads();
        for (int i = tid; i < a.ch_alloc * P_alloc; i += kTcThreads)
            reinterpret_cast<uint4*>(a.dbg)[(int64_t)blockIdx.x * a.ch_alloc * P_alloc + i] =
                *reinterpret_cast<const uint4*>(act0 + (size_t)(i / P_alloc) * chunk_stride + (size_t)(i % P_alloc) * 16);
    }
    fence_proxy_async();   // generic-proxy stores above -> visible to the tensor core's async-proxy reads
    __syncthreads();

    // ---- the block's tensor-core layers ---------------------------------------------------------------------
    unsigned char* cur = act0;
    unsigned char* nxt = act1;
    int tile_counter = 0;   // accumulator-slot uses so far (same sequence on the MMA and epilogue sides)
    for (int l = 0; l < a.n_layers; ++l) {
        const TcLayer L = a.layers[l];
        unsigned char* wcur = (l & 1) ? wbuf1 : wbuf0;
        if (warp == 0) {
            // Warp-uniform issue loop: every lane computes the same descriptors (so they live in uniform registers)
            // and one elected lane issues the tcgen05 instructions.
            mbar_wait(&hdr.wbar[l & 1], (uint32_t)((l >> 1) & 1));
            tc_fence_after();
            const uint32_t idesc = make_idesc(128, L.n);
            const uint32_t a_base = smem_u32(cur), b_base = smem_u32(wcur);
            const int ksteps = L.cin_chunks / 2;
            const uint64_t a_hi = make_desc(0, chunk_stride, 128), b_hi = make_desc(0, (uint32_t)L.n * 16, 128);
            const uint32_t b_tap_stride = (uint32_t)(L.cin_chunks * L.n * 16);
            const uint32_t b_k_stride = (uint32_t)(2 * L.n * 16);
            for (int mt = 0; mt < a.n_mt; ++mt) {
                const int it = tile_counter + mt;
                const int slot = it % kSlots;
                if (it >= kSlots) mbar_wait(&hdr.tmem_empty[slot], (uint32_t)(((it / kSlots) - 1) & 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(slot * kSlotCols);
                uint32_t acc = 0;
                for (int tap = 0; tap < 3; ++tap) {
                    const int shift = L.is_time ? tap * S : tap - 1;
                    uint32_t a_addr = a_base + (uint32_t)((mt * 128 + shift) * 16);
                    uint32_t b_addr = b_base + (uint32_t)tap * b_tap_stride;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint64_t ad = a_hi | (uint64_t)((a_addr >> 4) & 0x3FFF);
                        const uint64_t bd = b_hi | (uint64_t)((b_addr >> 4) & 0x3FFF);
                        if (elect_one()) umma_f16(d_tmem, ad, bd, idesc, acc);
                        acc = 1;
                        a_addr += 2 * chunk_stride;
                        b_addr += b_k_stride;
                    }
                }
                if (elect_one()) umma_commit(&hdr.tmem_full[slot]);
            }
            __syncwarp();
        } else if (warp == 1) {
            // prefetch the next layer's weights into the other buffer (its last readers finished before this layer began)
            if (lane == 0 && l + 1 < a.n_layers) {
                const TcLayer Ln = a.layers[l + 1];
                unsigned char* wn = ((l + 1) & 1) ? wbuf1 : wbuf0;
                mbar_expect_tx(&hdr.wbar[(l + 1) & 1], (uint32_t)Ln.w_bytes);
                bulk_g2s(wn, a.w_packed + Ln.w_off, (uint32_t)Ln.w_bytes, &hdr.wbar[(l + 1) & 1]);
            }
            __syncwarp();
        } else {
            const int e = warp - 2;           // 0..7
            const int group = e >> 2;         // even / odd tiles
            const int quad = warp & 3;        // TMEM lane quadrant this warp may access
            const float* bias = hdr.bias + l * 96;
            for (int mt = group; mt < a.n_mt; mt += 2) {
                const int it = tile_counter + mt;
                const int slot = it % kSlots;
                mbar_wait(&hdr.tmem_full[slot], (uint32_t)((it / kSlots) & 1));
                tc_fence_after();
                const int q = mt * 128 + quad * 32 + lane;
                const int r = (q - 1) / S, f = (q - 1) - r * S;
                const bool is_pad = (q < 1) || (f >= F);
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * kSlotCols);
                unsigned char* dst = nxt + (size_t)q * 16;
                for (int c0 = 0; c0 < L.n; c0 += 16) {
                    float v[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float x0 = leaky(v[2 * j] + bias[c0 + 2 * j]);
                        float x1 = leaky(v[2 * j + 1] + bias[c0 + 2 * j + 1]);
                        if (is_pad) { x0 = 0.f; x1 = 0.f; }
                        pk[j] = pack_half2(x0, x1);
                    }
                    *reinterpret_cast<uint4*>(dst + (size_t)(c0 >> 3) * chunk_stride) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *reinterpret_cast<uint4*>(dst + (size_t)((c0 >> 3) + 1) * chunk_stride) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&hdr.tmem_empty[slot]);
            }
            fence_proxy_async();
        }
        tile_counter += a.n_mt;
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        unsigned char* t = cur; cur = nxt; nxt = t;
        if (a.dbg != nullptr && a.dbg_layer == l) {
            for (int i = tid; i < a.ch_alloc * P_alloc; i += kTcThreads)
                reinterpret_cast<uint4*>(a.dbg)[(int64_t)blockIdx.x * a.ch_alloc * P_alloc + i] =
                    *reinterpret_cast<const uint4*>(cur + (size_t)(i / P_alloc) * chunk_stride + (size_t)(i % P_alloc) * 16);
        }
    }

    // ---- max-pool + store (fp16 chunk-major [clip][chunk][T_out][F_out][8]) -----------------------------------
    {
        const int out_chunks = a.layers[a.n_layers - 1].n / 8;
        const int Fo = F / a.pool_f;
        const int rows_p = a.rows_out / a.pool_t;           // pooled rows this tile produces
        const int rowp0 = row0_out / a.pool_t;
        const int total = out_chunks * rows_p * Fo;
        for (int i = tid; i < total; i += kTcThreads) {
            const int ch = i / (rows_p * Fo);
            const int rem = i - ch * rows_p * Fo;
            const int rp = rem / Fo, fo = rem - rp * Fo;
            if (rowp0 + rp >= a.T_out) continue;
            __half2 m[4];
            bool first = true;
            for (int dt = 0; dt < a.pool_t; ++dt)
                for (int df = 0; df < a.pool_f; ++df) {
                    const int q = 1 + (rp * a.pool_t + dt) * S + fo * a.pool_f + df;
                    const uint4 v = *reinterpret_cast<const uint4*>(cur + (size_t)ch * chunk_stride + (size_t)q * 16);
                    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) m[j] = first ? h[j] : __hmax2_nan(m[j], h[j]);
                    first = false;
                }
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&m[0]);
            o.y = *reinterpret_cast<uint32_t*>(&m[1]);
            o.z = *reinterpret_cast<uint32_t*>(&m[2]);
            o.w = *reinterpret_cast<uint32_t*>(&m[3]);
            reinterpret_cast<uint4*>(a.out)[(((int64_t)clip * out_chunks + ch) * a.T_out + rowp0 + rp) * Fo + fo] = o;
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// fp16 chunk-major [clips][chunks][T][F][8] -> f32 NHWC [clips][T][F][C] (C real channels <= chunks*8)
__global__ void chunked_to_nhwc_kernel(const __half* __restrict__ in, float* __restrict__ out, int clips, int chunks, int T,
                                       int F, int C) {
    const int64_t total = (int64_t)clips * T * F * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t r = i / C;
        const int f = (int)(r % F);
        r /= F;
        const int t = (int)(r % T);
        const int b = (int)(r / T);
        out[i] = __half2float(in[((((int64_t)b * chunks + (c >> 3)) * T + t) * F + f) * 8 + (c & 7)]);
    }
}

// debug dump [tiles][chunks][P_alloc][8] -> f32 NHWC [clips][T][F][C] keeping rows < rows_valid of every tile
__global__ void dump_to_nhwc_kernel(const __half* __restrict__ dbg, float* __restrict__ out, int clips, int tiles_per_clip,
                                    int chunks, int P_alloc, int S, int F, int rows_out, int T, int C) {
    const int64_t total = (int64_t)clips * T * F * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t r = i / C;
        const int f = (int)(r % F);
        r /= F;
        const int t = (int)(r % T);
        const int b = (int)(r / T);
        int tile = t / rows_out;
        if (tile >= tiles_per_clip) tile = tiles_per_clip - 1;
        const int rl = t - tile * rows_out;
        const int q = 1 + rl * S + f;
        out[i] = __half2float(dbg[((((int64_t)(b * tiles_per_clip + tile)) * chunks + (c >> 3)) * P_alloc + q) * 8 + (c & 7)]);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct TcBlockPlan {
    int first_layer;      // conv index of the first tensor-core layer of the block
    int n_layers;
    int F;                // input freq bins
    int cin_pad;          // padded input channels of the first tc layer
    int c_pad;            // padded channels inside the block
    int c_real;           // real output channels
    int pool_t, pool_f;
    int rows_out_max;     // pre-pool output rows per tile (upper bound)
};
// block 1: conv2d (CUDA cores) + conv2d_1..3;  block 2: conv2d_4..7;  block 3: conv2d_8..11;  block 4: conv2d_12..15 (no pool here)
static const TcBlockPlan kPlans[4] = {
    {1, 3, 32, 32, 32, 24, 2, 2, 36},
    {4, 4, 16, 32, 48, 48, 1, 2, 32},
    {8, 4, 8, 48, 80, 72, 2, 2, 30},
    {12, 4, 4, 80, 96, 96, 1, 1, 32},
};

struct TcWeights {
    unsigned char* w_packed = nullptr;
    float* bias_packed = nullptr;
    float* l0 = nullptr;
    TcLayer layers[16];   // indexed by conv index 1..15
};

static inline int pad_c(int c) { return c == 24 ? 32 : (c == 72 ? 80 : c); }

int tc_prepare(hb_embed_model* m, const float* weights_host) {
    TcWeights* tw = new TcWeights();
    std::vector<__half> packed;
    std::vector<float> bias;
    int64_t off = 0;
    std::vector<int64_t> w_off(kNumConv), b_off(kNumConv);
    for (int i = 0; i < kNumConv; ++i) {
        w_off[i] = off;
        off += layer_weight_floats(kLayers[i]);
        b_off[i] = off;
        off += kLayers[i].cout;
    }
    for (int li = 1; li <= 15; ++li) {
        const ConvLayer& L = kLayers[li];
        const int cin_p = pad_c(L.cin), n_p = pad_c(L.cout);
        TcLayer t;
        t.cin_chunks = cin_p / 8;
        t.n = n_p;
        t.is_time = (L.kh == 3);
        t.w_bytes = 3 * cin_p * n_p * 2;
        t.w_off = (int64_t)packed.size() * 2;
        t.bias_off = (int)bias.size();
        // B operand, K-major no-swizzle canonical layout: [tap][k chunk][n][8]
        const float* w = weights_host + w_off[li];  // [kh][kw][cin][cout]; exactly one of kh,kw is 3
        for (int tap = 0; tap < 3; ++tap)
            for (int kc = 0; kc < cin_p / 8; ++kc)
                for (int n = 0; n < n_p; ++n)
                    for (int j = 0; j < 8; ++j) {
                        const int ci = kc * 8 + j;
                        float v = 0.f;
                        if (ci < L.cin && n < L.cout) v = w[((int64_t)tap * L.cin + ci) * L.cout + n];
                        packed.push_back(__float2half_rn(v));
                    }
        for (int n = 0; n < n_p; ++n) bias.push_back(n < L.cout ? weights_host[b_off[li] + n] : 0.f);
        tw->layers[li] = t;
    }
    HB_CUDA_OK(cudaMalloc(&tw->w_packed, packed.size() * 2));
    HB_CUDA_OK(cudaMemcpy(tw->w_packed, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMalloc(&tw->bias_packed, bias.size() * sizeof(float)));
    HB_CUDA_OK(cudaMemcpy(tw->bias_packed, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMalloc(&tw->l0, (3 * 24 + 24) * sizeof(float)));
    HB_CUDA_OK(cudaMemcpy(tw->l0, weights_host + w_off[0], (3 * 24 + 24) * sizeof(float), cudaMemcpyHostToDevice));
    m->tc = tw;
    return HB_OK;
}

void tc_release(hb_embed_model* m) {
    TcWeights* tw = reinterpret_cast<TcWeights*>(m->tc);
    if (!tw) return;
    cudaFree(tw->w_packed);
    cudaFree(tw->bias_packed);
    cudaFree(tw->l0);
    delete tw;
    m->tc = nullptr;
}

// geometry of one block launch for clips of T_in input rows
struct TcGeom {
    int T_in, T_pre, T_out;      // input rows, pre-pool output rows (T_in - 4), pooled rows
    int tiles_per_clip, rows_out, Tt, n_mt, P_alloc, ch_alloc, w_buf_bytes;
    size_t smem;
};

static TcGeom tc_geometry(int b, int T_in, const TcWeights* tw) {
    const TcBlockPlan& p = kPlans[b];
    TcGeom g;
    g.T_in = T_in;
    g.T_pre = T_in - 4;
    g.T_out = g.T_pre / p.pool_t;
    const int need = ceil_div(g.T_pre, p.pool_t) * p.pool_t;  // cover every pre-pool row (the activation hook returns them all)
    g.tiles_per_clip = std::max(1, ceil_div(need, p.rows_out_max));
    g.rows_out = ceil_div(ceil_div(need, g.tiles_per_clip), p.pool_t) * p.pool_t;
    g.Tt = g.rows_out + 4;
    const int S = p.F + 1;
    const int P = 1 + g.Tt * S;
    g.n_mt = ceil_div(P, 128);
    g.P_alloc = g.n_mt * 128 + 2 * S + 8;
    g.ch_alloc = std::max(p.cin_pad, p.c_pad) / 8;
    g.w_buf_bytes = 0;
    for (int l = 0; l < p.n_layers; ++l) g.w_buf_bytes = std::max(g.w_buf_bytes, tw->layers[p.first_layer + l].w_bytes);
    g.w_buf_bytes = (g.w_buf_bytes + 127) & ~127;
    g.smem = ((sizeof(TcSmemHeader) + 127) & ~127) + 2 * (size_t)g.w_buf_bytes + 2 * (128 + (size_t)g.ch_alloc * g.P_alloc * 16) +
             (b == 0 ? (size_t)g.Tt * kMels * sizeof(float) : 0) + 128;
    return g;
}

// per-block global activation sizes (fp16 elements per clip) for F input frames
static void tc_chain(int F, const TcWeights* tw, TcGeom g[4]) {
    int T = F;
    for (int b = 0; b < 4; ++b) {
        g[b] = tc_geometry(b, T, tw);
        T = g[b].T_out;
    }
}

static int64_t block_out_halves(int b, const TcGeom& g) {
    const TcBlockPlan& p = kPlans[b];
    return (int64_t)(p.c_pad / 8) * g.T_out * (p.F / p.pool_f) * 8;
}

int64_t tc_workspace_bytes(int B, int F) {
    // fp16 activations between blocks + f32 conv2d_15 output + tail scratch; bounded generously by the fp32 path's size
    return fp32_workspace_bytes(B, F);
}

static int launch_block(const hb_embed_model* m, int b, const TcGeom& g, const void* in, __half* out, int B, __half* dbg,
                        int dbg_layer, cudaStream_t st) {
    const TcWeights* tw = reinterpret_cast<const TcWeights*>(m->tc);
    const TcBlockPlan& p = kPlans[b];
    TcBlockArgs a;
    a.in = in;
    a.out = out;
    a.w_packed = tw->w_packed;
    a.bias_packed = tw->bias_packed;
    a.l0_w = tw->l0;
    a.dbg = dbg;
    a.dbg_layer = dbg_layer;
    a.n_clips = B;
    a.T_in = g.T_in;
    a.T_out = g.T_out;
    a.F = p.F;
    a.S = p.F + 1;
    a.pool_t = p.pool_t;
    a.pool_f = p.pool_f;
    a.tiles_per_clip = g.tiles_per_clip;
    a.rows_out = g.rows_out;
    a.Tt = g.Tt;
    a.n_mt = g.n_mt;
    a.P_alloc = g.P_alloc;
    a.n_layers = p.n_layers;
    a.ch_alloc = g.ch_alloc;
    a.w_buf_bytes = g.w_buf_bytes;
    for (int l = 0; l < p.n_layers; ++l) a.layers[l] = tw->layers[p.first_layer + l];
    HB_REQUIRE(g.smem <= 227 * 1024, "tc block %d needs %zu bytes of shared memory", b, g.smem);
    const int grid = B * g.tiles_per_clip;
    if (b == 0) {
        HB_CUDA_OK(cudaFuncSetAttribute(tc_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        tc_block_kernel<true><<<grid, kTcThreads, g.smem, st>>>(a);
    } else {
        HB_CUDA_OK(cudaFuncSetAttribute(tc_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
        tc_block_kernel<false><<<grid, kTcThreads, g.smem, st>>>(a);
    }
    HB_LAUNCHED();
    return HB_OK;
}

static int check_timeout() {
    unsigned int flag = 0;
    HB_CUDA_OK(cudaMemcpyFromSymbol(&flag, g_tc_timeout, sizeof(flag)));
    HB_REQUIRE(flag == 0, "tcgen05 embed kernel: an mbarrier wait timed out (pipeline bug)");
    return HB_OK;
}

int tc_embed_clips(const hb_embed_model* m, const float* mel, int B, int F, const int32_t* slot_offsets_host, int n_slots,
                   float* out, void* ws, int64_t ws_bytes, cudaStream_t st) {
    const TcWeights* tw = reinterpret_cast<const TcWeights*>(m->tc);
    HB_REQUIRE(tw != nullptr, "tc weights missing");
    HB_REQUIRE(ws_bytes >= tc_workspace_bytes(B, F), "hb_embed: workspace too small");
    TcGeom g[4];
    tc_chain(F, tw, g);
    // workspace carve-up: two fp16 ping-pong activation buffers, then f32 pre-pool conv2d_15 output, then tail scratch
    int64_t max_halves = 0;
    for (int b = 0; b < 4; ++b) max_halves = std::max(max_halves, block_out_halves(b, g[b]));
    const int64_t act_bytes = ((int64_t)B * max_halves * 2 + 255) & ~255ll;
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    __half* hA = reinterpret_cast<__half*>(base);
    __half* hB = reinterpret_cast<__half*>(base + act_bytes);
    float* l15 = reinterpret_cast<float*>(base + 2 * act_bytes);
    const int T15 = g[3].T_out;
    const int64_t l15_floats = (int64_t)B * T15 * 4 * kEmbDim;
    float* scratch = l15 + l15_floats;
    const int64_t scratch_floats = (ws_bytes - 2 * act_bytes) / 4 - l15_floats;
    HB_REQUIRE(scratch_floats > 0, "hb_embed: workspace too small for the tail");

    int rc;
    if ((rc = launch_block(m, 0, g[0], mel, hA, B, nullptr, -1, st))) return rc;
    if ((rc = launch_block(m, 1, g[1], hA, hB, B, nullptr, -1, st))) return rc;
    if ((rc = launch_block(m, 2, g[2], hB, hA, B, nullptr, -1, st))) return rc;
    if ((rc = launch_block(m, 3, g[3], hA, hB, B, nullptr, -1, st))) return rc;
    {
        const int64_t total = l15_floats;
        const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), 148 * 8);
        chunked_to_nhwc_kernel<<<blocks, 256, 0, st>>>(hB, l15, B, 12, T15, 4, kEmbDim);
        HB_LAUNCHED();
    }
    return fp32_tail_from_l15(m, l15, B, T15, slot_offsets_host, n_slots, out, scratch, scratch_floats, st);
}

int64_t tc_activation(const hb_embed_model* m, const float* mel, int B, int F, int layer, float* out, int64_t cap,
                      void* ws, int64_t ws_bytes, cudaStream_t st) {
    const TcWeights* tw = reinterpret_cast<const TcWeights*>(m->tc);
    if (!tw || ws_bytes < tc_workspace_bytes(B, F)) {
        set_error("hb_embed_activation(f16): bad workspace");
        return HB_ERR_INVALID;
    }
    if (layer > 15) {
        set_error("hb_embed_activation(f16): layers 16..19 run on the fp32 tail; query them in fp32 mode");
        return HB_ERR_UNSUPPORTED;
    }
    TcGeom g[4];
    tc_chain(F, tw, g);
    int64_t max_halves = 0;
    for (int b = 0; b < 4; ++b) max_halves = std::max(max_halves, block_out_halves(b, g[b]));
    const int64_t act_bytes = ((int64_t)B * max_halves * 2 + 255) & ~255ll;
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    __half* bufs[2] = {reinterpret_cast<__half*>(base), reinterpret_cast<__half*>(base + act_bytes)};
    __half* dbg = reinterpret_cast<__half*>(base + 2 * act_bytes);
    const int target_block = layer < 4 ? 0 : (layer < 8 ? 1 : (layer < 12 ? 2 : 3));
    const void* in = mel;
    int which = 0;
    for (int b = 0; b <= target_block; ++b) {
        const TcBlockPlan& p = kPlans[b];
        const bool last = (b == target_block);
        const int last_layer_of_block = p.first_layer + p.n_layers - 1;
        int dbg_layer = -1;
        if (last && layer != last_layer_of_block) dbg_layer = (layer == 0) ? 100 : layer - p.first_layer;
        if (dbg_layer >= 0) {
            const int64_t need = (int64_t)B * g[b].tiles_per_clip * g[b].ch_alloc * g[b].P_alloc * 16;
            if (2 * act_bytes + need > ws_bytes) {
                set_error("hb_embed_activation(f16): workspace too small for the debug dump");
                return HB_ERR_INVALID;
            }
        }
        int rc = launch_block(m, b, g[b], in, bufs[which], B, dbg_layer >= 0 ? dbg : nullptr, dbg_layer, st);
        if (rc) return rc;
        if (last) {
            int T, Fq, C;
            if (dbg_layer >= 0) {
                // rows valid after this layer: time convs done so far inside the block shrink T
                int t_convs = 0;
                for (int li = (b == 0 ? 0 : p.first_layer); li <= layer; ++li) t_convs += (kLayers[li].kh == 3);
                T = g[b].T_in - 2 * t_convs;
                Fq = p.F;
                C = kLayers[layer].cout;
                const int64_t n = (int64_t)B * T * Fq * C;
                if (n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
                const int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), 148 * 8);
                dump_to_nhwc_kernel<<<blocks, 256, 0, st>>>(dbg, out, B, g[b].tiles_per_clip, g[b].ch_alloc, g[b].P_alloc, p.F + 1,
                                                            Fq, g[b].rows_out, T, C);
                if (cudaGetLastError() != cudaSuccess) { set_error("dump kernel launch failed"); return HB_ERR_CUDA; }
                if (cudaStreamSynchronize(st) != cudaSuccess || check_timeout() != HB_OK) return HB_ERR_CUDA;
                return n;
            }
            T = g[b].T_out;
            Fq = p.F / p.pool_f;
            C = p.c_real;
            if (b == 3) {
                // conv2d_15 "after its pool" = 2x2 pool phase 0 of the pre-pool output
                const int64_t n_pre = (int64_t)B * T * 4 * kEmbDim;
                float* pre = reinterpret_cast<float*>(dbg);
                if (2 * act_bytes + n_pre * 4 > ws_bytes) { set_error("workspace too small"); return HB_ERR_INVALID; }
                const int blocks = (int)std::min<int64_t>(ceil_div64(n_pre, 256), 148 * 8);
                chunked_to_nhwc_kernel<<<blocks, 256, 0, st>>>(bufs[which], pre, B, 12, T, 4, kEmbDim);
                const int To = T / 2;
                const int64_t n = (int64_t)B * To * 2 * kEmbDim;
                if (n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
                int rc2 = fp32_pool_public(pre, out, B, T, 4, kEmbDim, 2, 2, 0, st);
                if (rc2) return rc2;
                if (cudaStreamSynchronize(st) != cudaSuccess || check_timeout() != HB_OK) return HB_ERR_CUDA;
                return n;
            }
            const int64_t n = (int64_t)B * T * Fq * C;
            if (n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
            const int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), 148 * 8);
            chunked_to_nhwc_kernel<<<blocks, 256, 0, st>>>(bufs[which], out, B, p.c_pad / 8, T, Fq, C);
            if (cudaGetLastError() != cudaSuccess) { set_error("convert kernel launch failed"); return HB_ERR_CUDA; }
            if (cudaStreamSynchronize(st) != cudaSuccess || check_timeout() != HB_OK) return HB_ERR_CUDA;
            return n;
        }
        in = bufs[which];
        which ^= 1;
    }
    return HB_ERR_INVALID;
}

}  // namespace hb
