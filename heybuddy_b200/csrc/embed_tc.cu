// K7 (product mode): speech-embedding conv stack as tcgen05 implicit GEMMs with TMEM accumulators.
//
// Replaces the ORT run of speech-embedding.onnx behind SpeechEmbeddingModel.__call__ (reference
// src/python/heybuddy/embeddings.py:32-42).  Layer table: embed_common.cuh / spec.py.
//
// Design (DESIGN.md "embedding conv stack"):
//  * Every conv is 1xk (freq) or kx1 (time).  Activations live in shared memory position-major --
//    position q = 1 + row*(F+1) + f with one shared zero pad column per row -- as
//    [channel chunk of 8][position][8 x fp16], which is the K-major, no-swizzle canonical UMMA
//    layout (core matrix = 8 positions x 16 B, SBO = 128 B, LBO = P_alloc*16 B).  A conv tap is a
//    pure shift of the position index, i.e. a shift of the descriptor start address: no im2col.
//  * Orientation: D^T[cout, position] = sum_tap W_tap^T[cout, cin] X^T[cin, position + shift].
//    A = weights (M = 128 rows, cout zero padded / aliased), B = activations with N = 256 positions
//    per tcgen05.mma (kind::f16, fp16 operands, fp32 accumulation in TMEM).  Measured on B200
//    (scripts/micro/umma_swizzle.cu): an SS tcgen05.mma with K = 16 costs N / 2 cycles whatever M is,
//    so the small dimension (cout = 24..96) sits in M and the long one (positions) in N, and N is the
//    tile's real position count rounded up to 16.  Blocks 1-3 go further (embed_tcg.cu: several
//    positions per column); this file runs block 4 and the tail.
//  * One CTA owns a tile (a time slice of one clip, or 6 whole clips for the tail) and runs the block's
//    convs back to back, activations ping-ponging between two shared-memory buffers, the next layer's
//    pre-packed weights arriving by cp.async.bulk (mbarrier tx-count) while the current layer computes.
//  * Warp roles: warp 0 = MMA issuer (warp-uniform loop, one elected lane) + TMEM allocator, warp 1 =
//    weight loader, warps 2-9 = epilogue.  The accumulator is transposed (lane = channel, column =
//    position): tcgen05.ld.16x256b hands thread t channel t/4 (+8) x positions 2(t%4), +1, which is
//    exactly the fragment stmatrix.m8n8.trans wants -> +bias, LeakyReLU, fp16, and one stmatrix writes
//    8 positions x 8 channels (16 B rows) straight into the next layer's operand buffer.  Two 256-column
//    TMEM slots overlap the MMAs of tile i+1 with the epilogue of tile i.
//  * conv2d (Cin = 1) is CUDA-core work fused into block 1's prologue; the last 2x2 max-pool (both
//    phases) is fused into the tail block's loader; block 5 runs on the same kernel, 6 clips per CTA.
#include "tc_ptx.cuh"

#include <cstdlib>
#include <vector>

namespace hb {

// ------------------------------------------------------------------------------------------------
// Block configuration
// ------------------------------------------------------------------------------------------------
// Two launch shapes: "wide"  = 576 threads (16 epilogue warps), 1 CTA/SM, 2 accumulator slots, double-buffered weights;
//                    "twin"  = 320 threads (8 epilogue warps), 2 CTAs/SM (each 256 TMEM columns, 1 slot, single weight
//                              buffer) -- the co-resident CTA fills the other's staging / epilogue / store phases.
constexpr int kMaxSlots = 2;      // TMEM accumulator slots
constexpr int kTmemCols = 512;
constexpr int kMaxTcLayers = 4;

struct TcLayer {
    int cin_chunks;   // padded Cin / 8
    int n_out;        // padded Cout
    int ntaps;
    int tap_rows[3];  // shift in rows (multiples of S)
    int tap_cols[3];  // shift in positions within the row
    int leaky;
    int w_rows;       // rows of one (tap, k chunk) region of the packed A operand (128 = explicit, < 128 = aliased)
    int w_bytes;      // packed A operand bytes
    int64_t w_off;    // byte offset in the packed weight buffer
    int bias_off;     // float offset in the packed bias buffer
    signed char chunk_of[4][4];  // output-channel chunk held by (TMEM lane quadrant, octet); -1 = none
};

struct TcBlockArgs {
    const void* in;        // in_mode 0: mel f32 [clips][in_T][32]; 1/2: fp16 chunk-major [clips][in_chunks][in_T][in_F][8]
    void* out;             // out_mode 0: fp16 chunk-major [clips][out_chunks][T_out][F_out][8]; 1: f32 [clips][rows_valid][96]
    const unsigned char* w_packed;
    const float* bias_packed;
    const float* l0_w;     // in_mode 0: conv2d kernel f32 [3][24] + bias [24]
    __half* dbg;           // optional activation dump [ctas][chunks][P_alloc][8]
    int dbg_layer;         // layer index within the block to dump (-1: none; 100: the staged input)
    int in_mode, out_mode;
    int n_clips, in_T, in_F, in_chunks;
    int T_out;             // out_mode 0: pooled rows per clip; out_mode 1: valid rows per clip
    int F, S;              // freq bins of the block's activations, S = F + 1
    int pool_t, pool_f, pool_phase;
    int tiles_per_clip, rows_out, Tt, segs;
    int n_nt, P_alloc, n_layers, ch_alloc, w_buf_bytes;
    int n_slots, tmem_cols, w_double;   // launch shape (see kMaxSlots)
    int in_place;          // 1: one activation buffer rewritten in place (only with a single accumulator tile per layer)
    TcLayer layers[kMaxTcLayers];
};

struct TcSmemHeader {
    uint64_t tmem_full[kMaxSlots];
    uint64_t tmem_empty[kMaxSlots];
    uint64_t wbar[2];
    uint32_t tmem_base;
    uint32_t pad[3];
    float bias[kMaxTcLayers * 96];
    float l0[3 * 24 + 24 + 8];
};


// kTileN = positions per tcgen05.mma = accumulator slot columns (256, or 128 for the small late blocks)
template <int kTcThreads, int kMinBlocks, int kTileN>
__global__ void __launch_bounds__(kTcThreads, kMinBlocks) tc_block_kernel(const TcBlockArgs a) {
    constexpr int kEpiWarps = kTcThreads / 32 - 2;   // 8 or 16: kEpiWarps / 4 warps per TMEM lane quadrant
    constexpr int kColsPerWarp = kTileN / (kEpiWarps / 4);
    const int kSlots = a.n_slots;
    const uint32_t kTmemCols = (uint32_t)a.tmem_cols;
    extern __shared__ __align__(128) unsigned char smem[];
    TcSmemHeader& hdr = *reinterpret_cast<TcSmemHeader*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P_alloc = a.P_alloc;
    const uint32_t chunk_stride = (uint32_t)P_alloc * 16u;               // bytes between channel chunks
    // per-position table: bit 0 = zero this position (pad column / position 0), bits 1.. = output row + 1 (out_mode 1)
    int16_t* pos_tab = reinterpret_cast<int16_t*>(smem + ((sizeof(TcSmemHeader) + 15) & ~15));
    unsigned char* wbuf0 = smem + ((sizeof(TcSmemHeader) + 2 * P_alloc + 127 + 16) & ~127);
    unsigned char* wbuf1 = a.w_double ? wbuf0 + a.w_buf_bytes : wbuf0;
    unsigned char* act0 = wbuf1 + a.w_buf_bytes + 128;                    // +128: guard for the position -1 read
    // in place: a layer with one accumulator tile has completed all its MMAs (tcgen05.commit) before the first epilogue store
    unsigned char* act1 = a.in_place ? act0 : act0 + (size_t)a.ch_alloc * chunk_stride + 128;
    float* mel_tile = reinterpret_cast<float*>(act1 + (size_t)a.ch_alloc * chunk_stride);  // in_mode 0 only

    int clip0, tile;
    if (a.segs > 1) { clip0 = blockIdx.x * a.segs; tile = 0; }
    else { clip0 = blockIdx.x / a.tiles_per_clip; tile = blockIdx.x - clip0 * a.tiles_per_clip; }
    const int row0 = tile * a.rows_out;            // convs are top aligned: output row r reads input rows r..r+4
    const int S = a.S, F = a.F, Tt = a.Tt;
    const int seg_pos = Tt * S;
    const int P = 1 + a.segs * seg_pos;

    TC_STAMP(0);
    // ---- one-time setup ----------------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kSlots; ++i) {
            mbar_init(&hdr.tmem_full[i], 1);
            mbar_init(&hdr.tmem_empty[i], kEpiWarps);
        }
        mbar_init(&hdr.wbar[0], 1);
        mbar_init(&hdr.wbar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&hdr.tmem_base, kTmemCols);
    for (int i = tid; i < a.n_layers * 96; i += kTcThreads) {
        const int l = i / 96, c = i - l * 96;
        hdr.bias[i] = (c < a.layers[l].n_out) ? a.bias_packed[a.layers[l].bias_off + c] : 0.f;
    }
    if (a.in_mode == 0)
        for (int i = tid; i < 3 * 24 + 24; i += kTcThreads) hdr.l0[i] = a.l0_w[i];
    if (a.out_mode == 1) {
        for (int q = tid; q < P_alloc; q += kTcThreads) {
            int v = 1;
            if (q >= 1 && q < P) {
                const int rem = (q - 1) % seg_pos, seg = (q - 1) / seg_pos;
                const int r = rem / S, f = rem - r * S;
                v = (f >= F) ? 1 : 0;
                if (f == 0 && r < a.rows_out && row0 + r < a.T_out && clip0 + seg < a.n_clips) v |= (seg * 64 + r + 1) << 1;
            }
            pos_tab[q] = (int16_t)v;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr.tmem_base;
    TC_STAMP(1);

    // first layer's weights start streaming while the input tile is staged
    if (warp == 1 && lane == 0) {
        mbar_expect_tx(&hdr.wbar[0], (uint32_t)a.layers[0].w_bytes);
        bulk_g2s(wbuf0, a.w_packed + a.layers[0].w_off, (uint32_t)a.layers[0].w_bytes, &hdr.wbar[0]);
    }

    // ---- stage the input tile into act0 ------------------------------------------------------------------
    if (a.in_mode == 0) {
        // mel rows [row0, row0 + Tt) -> shared (zero beyond the clip), then conv2d (Cin = 1) on CUDA cores
        const float* mel = reinterpret_cast<const float*>(a.in) + (int64_t)clip0 * a.in_T * kMels;
        for (int i = tid; i < Tt * kMels; i += kTcThreads) {
            const int r = i / kMels;
            mel_tile[i] = (row0 + r < a.in_T) ? __ldg(mel + (int64_t)(row0 + r) * kMels + (i - r * kMels)) : 0.f;
        }
        __syncthreads();
        // lane -> (8-channel chunk = lane & 3, its 24 weights + 8 biases in registers; freq bins (lane >> 2) + 8 j);
        // warp -> rows.  F == 32 here, so one warp row-pass covers the 32 bins with no index division.
        {
            const int ch = lane & 3;
            float w0[8], w1[8], w2[8], bb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = ch * 8 + j;
                w0[j] = ch < 3 ? hdr.l0[c] : 0.f;
                w1[j] = ch < 3 ? hdr.l0[24 + c] : 0.f;
                w2[j] = ch < 3 ? hdr.l0[48 + c] : 0.f;
                bb[j] = ch < 3 ? hdr.l0[72 + c] : 0.f;
            }
            unsigned char* dst_ch = act0 + ch * chunk_stride;
            for (int r = warp; r < Tt; r += kTcThreads / 32) {
                const float* mrow = mel_tile + r * kMels;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int f = (lane >> 2) + 8 * j4;
                    uint4 pk = make_uint4(0, 0, 0, 0);
                    if (ch < 3) {
                        const float m1 = mrow[f];
                        const float m0 = f > 0 ? mrow[f - 1] : 0.f;
                        const float m2 = f < kMels - 1 ? mrow[f + 1] : 0.f;
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float acc = fmaf(m0, w0[j], 0.f);
                            acc = fmaf(m1, w1[j], acc);
                            acc = fmaf(m2, w2[j], acc);
                            v[j] = leaky(acc + bb[j]);
                        }
                        pk = make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                    }
                    *reinterpret_cast<uint4*>(dst_ch + (size_t)(1 + r * S + f) * 16) = pk;
                }
            }
            // position 0, the pad column of every row, and the positions past the tile: zeros
            for (int i = tid; i < 4 * (Tt + 1 + (P_alloc - P)); i += kTcThreads) {
                const int c = i & 3, k = i >> 2;
                const int q = k == 0 ? 0 : (k <= Tt ? k * S : P + (k - Tt - 1));
                *reinterpret_cast<uint4*>(act0 + c * chunk_stride + (size_t)q * 16) = make_uint4(0, 0, 0, 0);
            }
        }
    } else if (a.in_mode == 1) {
        // plain chunk-major copy: every 16-byte position record goes global -> shared with cp.async (no register
        // round trip, all copies of the tile in flight at once); pads / rows beyond the clip are zero-filled
        const uint4* in = reinterpret_cast<const uint4*>(a.in) + (int64_t)clip0 * a.in_chunks * a.in_T * a.in_F;
        const int total = a.in_chunks * P_alloc;
        int ch = tid / P_alloc, q = tid - ch * P_alloc;           // running (chunk, position) of this thread
        const int dch = kTcThreads / P_alloc, dq = kTcThreads - dch * P_alloc;
        for (int i = tid; i < total; i += kTcThreads) {
            const int r = (q - 1) / S, f = (q - 1) - r * S;
            const bool real = q >= 1 && q < P && f < F && row0 + r < a.in_T;
            const uint4* src = real ? in + ((int64_t)ch * a.in_T + row0 + r) * a.in_F + f : in;
            cp_async16(act0 + ch * chunk_stride + (size_t)q * 16, src, real ? 16u : 0u);
            q += dq; ch += dch;
            if (q >= P_alloc) { q -= P_alloc; ++ch; }
        }
        cp_async_wait_all();
    } else {
        // tail: 2x2 max-pool with a time phase while loading (the two pool phases feed window offsets 0 / 4 mod 8)
        const uint4* in = reinterpret_cast<const uint4*>(a.in);
        const int in_chunks = a.in_chunks;
        for (int i = tid; i < in_chunks * P_alloc; i += kTcThreads) {
            const int ch = i / P_alloc, q = i - ch * P_alloc;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (q >= 1 && q < P) {
                const int seg = (q - 1) / seg_pos, rem = (q - 1) - seg * seg_pos;
                const int r = rem / S, f = rem - r * S;
                const int clip = clip0 + seg;
                const int tr = 2 * (row0 + r) + a.pool_phase;
                if (f < F && clip < a.n_clips && tr + 1 < a.in_T) {
                    const uint4* base = in + ((int64_t)clip * in_chunks + ch) * a.in_T * a.in_F;
                    const uint4 x0 = __ldg(base + (int64_t)tr * a.in_F + 2 * f), x1 = __ldg(base + (int64_t)tr * a.in_F + 2 * f + 1);
                    const uint4 x2 = __ldg(base + (int64_t)(tr + 1) * a.in_F + 2 * f), x3 = __ldg(base + (int64_t)(tr + 1) * a.in_F + 2 * f + 1);
                    const __half2* h0 = reinterpret_cast<const __half2*>(&x0);
                    const __half2* h1 = reinterpret_cast<const __half2*>(&x1);
                    const __half2* h2 = reinterpret_cast<const __half2*>(&x2);
                    const __half2* h3 = reinterpret_cast<const __half2*>(&x3);
                    __half2 m[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) m[j] = __hmax2_nan(__hmax2_nan(h0[j], h1[j]), __hmax2_nan(h2[j], h3[j]));
                    v = *reinterpret_cast<uint4*>(m);
                }
            }
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
    TC_STAMP(2);

    // ---- the block's tensor-core layers ---------------------------------------------------------------------
    unsigned char* cur = act0;
    unsigned char* nxt = act1;
    int tile_counter = 0;   // accumulator-slot uses so far (same sequence on the MMA and epilogue sides)
    for (int l = 0; l < a.n_layers; ++l) {
        const TcLayer& L = a.layers[l];
        unsigned char* wcur = (l & 1) ? wbuf1 : wbuf0;
        const int wb = a.w_double ? (l & 1) : 0;                        // barrier of this layer's weight buffer
        const uint32_t wpar = a.w_double ? (uint32_t)((l >> 1) & 1) : (uint32_t)(l & 1);
        const bool last_f32 = (a.out_mode == 1 && l == a.n_layers - 1);
        if (warp == 0) {
            // Warp-uniform issue loop: every lane computes the same descriptors (uniform registers), one elected lane issues.
            TC_FINE(0, 0);
            mbar_wait(&hdr.wbar[wb], wpar);
            tc_fence_after();
            TC_FINE(0, 1);
            // One elected lane issues the whole layer (a lean loop: the issue path must not be slower than the MMAs it feeds:
            // an N = 256 MMA takes 128 cycles, scripts/micro/umma_swizzle.cu).
            if (elect_one()) {
                const int ksteps = L.cin_chunks / 2;
                const uint32_t w_region16 = (uint32_t)L.w_rows;                // (bytes of one (tap, k chunk) region = LBO of A) >> 4
                const uint64_t a_hi = make_desc(0, w_region16 * 16u, 128), b_hi = make_desc(0, chunk_stride, 128);
                const uint32_t w_base16 = smem_u32(wcur) >> 4, x_base16 = smem_u32(cur) >> 4, chunk16 = chunk_stride >> 4;
                int shift[3];
                for (int tap = 0; tap < 3; ++tap) shift[tap] = L.tap_rows[tap] * S + L.tap_cols[tap];
                const int ntaps = L.ntaps, cin_chunks = L.cin_chunks;
                for (int nt = 0; nt < a.n_nt; ++nt) {
                    const int it = tile_counter + nt;
                    const int slot = it % kSlots;
                    if (it >= kSlots) mbar_wait(&hdr.tmem_empty[slot], (uint32_t)(((it / kSlots) - 1) & 1));
                    tc_fence_after();
                    TC_FINE(0, 2 + 2 * nt);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(slot * kTileN);
                    // N = the tile's real positions rounded up to 16 (an MMA costs N / 2 cycles): the last tile is usually short
                    const uint32_t idesc = make_idesc(128, min(kTileN, (P - nt * kTileN + 15) & ~15));
                    uint32_t acc = 0;
#pragma unroll
                    for (int tap = 0; tap < 3; ++tap) {
                        if (tap >= ntaps) break;
                        uint32_t w16 = w_base16 + (uint32_t)(tap * cin_chunks) * w_region16;
                        uint32_t x16 = x_base16 + (uint32_t)(nt * kTileN + shift[tap]);
                        for (int ks = 0; ks < ksteps; ++ks) {
                            umma_f16(d_tmem, a_hi | (uint64_t)w16, b_hi | (uint64_t)x16, idesc, acc);
                            acc = 1;
                            w16 += 2 * w_region16;
                            x16 += 2 * chunk16;
                        }
                    }
                    umma_commit(&hdr.tmem_full[slot]);
                    TC_FINE(0, 3 + 2 * nt);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // prefetch the next layer's weights into the other buffer (its last readers finished before this layer began)
            if (lane == 0 && l + 1 < a.n_layers && a.w_double) {
                const TcLayer& Ln = a.layers[l + 1];
                unsigned char* wn = ((l + 1) & 1) ? wbuf1 : wbuf0;
                mbar_expect_tx(&hdr.wbar[(l + 1) & 1], (uint32_t)Ln.w_bytes);
                bulk_g2s(wn, a.w_packed + Ln.w_off, (uint32_t)Ln.w_bytes, &hdr.wbar[(l + 1) & 1]);
            }
            if (!a.w_double && l + 1 < a.n_layers) {
                // single weight buffer: refill it for the next layer as soon as this layer's last MMAs have completed, i.e.
                // under this layer's last epilogue instead of after the layer barrier
                const int it = tile_counter + a.n_nt - 1;
                mbar_wait(&hdr.tmem_full[it % kSlots], (uint32_t)((it / kSlots) & 1));
                if (lane == 0) {
                    const TcLayer& Ln = a.layers[l + 1];
                    mbar_expect_tx(&hdr.wbar[0], (uint32_t)Ln.w_bytes);
                    bulk_g2s(wbuf0, a.w_packed + Ln.w_off, (uint32_t)Ln.w_bytes, &hdr.wbar[0]);
                }
            }
            __syncwarp();
        } else {
            const int e = warp - 2;            // 0..kEpiWarps-1
            const int quad = warp & 3;         // TMEM lane quadrant this warp may access
            const int part = e >> 2;           // which kColsPerWarp columns of every 256-column tile
            const float* bias = hdr.bias + l * 96;
            const uint32_t nxt_base = smem_u32(nxt);
            // per-layer invariants of this warp: chunks held by its two 16-lane halves, their biases, stmatrix row addresses
            int c0[2], c1[2];
            float bb0[2], bb1[2];
            uint32_t rows[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                c0[h] = L.chunk_of[quad][2 * h];
                c1[h] = L.chunk_of[quad][2 * h + 1];
                bb0[h] = c0[h] >= 0 ? bias[c0[h] * 8 + (lane >> 2)] : 0.f;
                bb1[h] = c1[h] >= 0 ? bias[c1[h] * 8 + (lane >> 2)] : 0.f;
                // Two chunks: x4 = [c0 | c1 | c0 (+8 positions) | c1 (+8 positions)], lanes 8m..8m+7 address matrix m.
                // One chunk: x2 = [c0 | c0 (+8 positions)].
                const int m = lane >> 3;
                rows[h] = c1[h] >= 0
                    ? nxt_base + (uint32_t)((m & 1) ? c1[h] : c0[h]) * chunk_stride + (uint32_t)(((m >> 1) * 8 + (lane & 7)) * 16)
                    : nxt_base + (uint32_t)(c0[h] < 0 ? 0 : c0[h]) * chunk_stride + (uint32_t)(((m & 1) * 8 + (lane & 7)) * 16);
            }
            for (int nt = 0; nt < a.n_nt; ++nt) {
                const int it = tile_counter + nt;
                const int slot = it % kSlots;
                if (warp == 2) TC_FINE(16, 3 * nt); else if (warp == 9) TC_FINE(32, 3 * nt);
                mbar_wait(&hdr.tmem_full[slot], (uint32_t)((it / kSlots) & 1));
                tc_fence_after();
                if (warp == 2) TC_FINE(16, 3 * nt + 1); else if (warp == 9) TC_FINE(32, 3 * nt + 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (c0[h] < 0) continue;   // octets are filled in order: c1 >= 0 implies c0 >= 0
                    for (int sub = 0; sub < kColsPerWarp / 64; ++sub) {
                        const int col = part * kColsPerWarp + sub * 64;
                        if (nt * kTileN + col >= P) continue;   // columns past the tile's positions were never written by the MMAs
                        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32 + h * 16) << 16) + (uint32_t)(slot * kTileN + col);
                        const int pos0 = nt * kTileN + col;
                        if (!last_f32) {
                            const uint32_t ra = rows[h] + (uint32_t)(pos0 * 16);
                            if (c1[h] >= 0) {
                                if (L.leaky) epilogue_sub<true, true>(taddr, ra, bb0[h], bb1[h]);
                                else epilogue_sub<true, false>(taddr, ra, bb0[h], bb1[h]);
                            } else {
                                if (L.leaky) epilogue_sub<false, true>(taddr, ra, bb0[h], bb1[h]);
                                else epilogue_sub<false, false>(taddr, ra, bb0[h], bb1[h]);
                            }
                        } else {
                            // final layer of the tail: f32 rows [clip][row][96] of the valid positions (no activation)
                            float v[32];
                            tmem_ld_16x256b_64cols(taddr, v);
                            float* o = reinterpret_cast<float*>(a.out);
                            const int ch_a = c0[h] * 8 + (lane >> 2), ch_b = c1[h] * 8 + (lane >> 2);
#pragma unroll
                            for (int g = 0; g < 8; ++g) {
                                const int p = pos0 + g * 8 + 2 * (lane & 3);
                                const int t0 = pos_tab[p], t1 = pos_tab[p + 1];
                                if (t0 >> 1) {
                                    const int sr = (t0 >> 1) - 1;   // seg * 64 + local row
                                    const int64_t row = (int64_t)(clip0 + (sr >> 6)) * a.T_out + row0 + (sr & 63);
                                    o[row * kEmbDim + ch_a] = v[4 * g + 0] + bb0[h];
                                    if (c1[h] >= 0) o[row * kEmbDim + ch_b] = v[4 * g + 2] + bb1[h];
                                }
                                if (t1 >> 1) {
                                    const int sr = (t1 >> 1) - 1;
                                    const int64_t row = (int64_t)(clip0 + (sr >> 6)) * a.T_out + row0 + (sr & 63);
                                    o[row * kEmbDim + ch_a] = v[4 * g + 1] + bb0[h];
                                    if (c1[h] >= 0) o[row * kEmbDim + ch_b] = v[4 * g + 3] + bb1[h];
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (warp == 2) TC_FINE(16, 3 * nt + 2); else if (warp == 9) TC_FINE(32, 3 * nt + 2);
                if (lane == 0) mbar_arrive(&hdr.tmem_empty[slot]);
            }
            fence_proxy_async();
        }
        tile_counter += a.n_nt;
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        TC_STAMP(3 + 2 * l);
        unsigned char* t = cur; cur = nxt; nxt = t;
        // The epilogue does not mask the pad column (and position 0): their values only ever feed pad outputs, except
        // through a conv with column taps (freq SAME padding) -- zero them right before such a layer.
        if (l + 1 < a.n_layers && (a.layers[l + 1].tap_cols[0] != 0 || a.layers[l + 1].tap_cols[1] != 0 || a.layers[l + 1].tap_cols[2] != 0)) {
            const int n_pad = a.segs * Tt + 1, chunks = L.n_out / 8;
            for (int i = tid; i < n_pad * chunks; i += kTcThreads) {
                const int ch = i / n_pad, k = i - ch * n_pad;
                const int q = k == 0 ? 0 : k * S;   // position 0, then the pad column 1 + (k-1)*S + F = k*S of every row
                *reinterpret_cast<uint4*>(cur + (size_t)ch * chunk_stride + (size_t)q * 16) = make_uint4(0, 0, 0, 0);
            }
            fence_proxy_async();
            __syncthreads();
        }
        if (a.dbg != nullptr && a.dbg_layer == l) {
            for (int i = tid; i < a.ch_alloc * P_alloc; i += kTcThreads)
                reinterpret_cast<uint4*>(a.dbg)[(int64_t)blockIdx.x * a.ch_alloc * P_alloc + i] =
                    *reinterpret_cast<const uint4*>(cur + (size_t)(i / P_alloc) * chunk_stride + (size_t)(i % P_alloc) * 16);
        }
        TC_STAMP(4 + 2 * l);
    }

    // ---- max-pool + store (fp16 chunk-major [clip][chunk][T_out][F_out][8]) -----------------------------------
    if (a.out_mode == 0) {
        const int out_chunks = a.layers[a.n_layers - 1].n_out / 8;
        const int Fo = F / a.pool_f;
        const int rows_p = a.rows_out / a.pool_t;           // pooled rows this tile produces
        const int rowp0 = row0 / a.pool_t;
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
            reinterpret_cast<uint4*>(a.out)[(((int64_t)clip0 * out_chunks + ch) * a.T_out + rowp0 + rp) * Fo + fo] = o;
        }
    }
    __syncthreads();
    TC_STAMP(11);
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
    TC_STAMP(12);
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

// debug dump [ctas][chunks][P_alloc][8] -> f32 NHWC [clips][T][F][C] keeping rows < rows_out of every tile
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
    int F;                // freq bins of the block's activations
    int cin_pad;          // padded input channels of the first tc layer
    int c_pad;            // padded channels inside the block
    int c_real;           // real output channels
    int pool_t, pool_f;
    int rows_out_max;     // pre-pool output rows per tile (upper bound, sized for shared memory)
    int spread;           // 1: output chunks spread over the four TMEM lane quadrants (explicit 128-row A operand)
    int twin;             // 1: "twin" launch shape (2 CTAs/SM), 0: "wide" (1 CTA/SM)
    int tile_n;           // positions per MMA / accumulator slot (256 or 128)
};
// block 1: conv2d (CUDA cores) + conv2d_1..3;  2: conv2d_4..7;  3: conv2d_8..11;  4: conv2d_12..15 (pool deferred);  5: conv2d_16..19
static const TcBlockPlan kPlans[5] = {
    {1, 3, 32, 32, 32, 24, 2, 2, 10, 1, 1, 256},
    {4, 4, 16, 32, 48, 48, 1, 2, 11, 1, 1, 256},
    {8, 4, 8, 48, 80, 72, 2, 2, 10, 0, 1, 128},
    {12, 4, 4, 80, 96, 96, 1, 1, 26, 0, 1, 256},
    {16, 4, 2, 96, 96, 96, 1, 1, 0, 0, 1, 256},
};

struct TcWeights {
    unsigned char* w_packed = nullptr;
    float* bias_packed = nullptr;
    float* l0 = nullptr;
    TcLayer layers[kNumConv];   // indexed by conv index 1..19
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
    for (int b = 0; b < 5; ++b) {
        const TcBlockPlan& p = kPlans[b];
        for (int li = p.first_layer; li < p.first_layer + p.n_layers; ++li) {
            const ConvLayer& L = kLayers[li];
            const int cin_p = pad_c(L.cin), n_p = pad_c(L.cout);
            TcLayer t;
            t.cin_chunks = cin_p / 8;
            t.n_out = n_p;
            t.ntaps = L.kh * L.kw;
            for (int tap = 0; tap < 3; ++tap) { t.tap_rows[tap] = 0; t.tap_cols[tap] = 0; }
            for (int tap = 0; tap < t.ntaps; ++tap) {
                if (L.kh > 1) t.tap_rows[tap] = tap;                    // time conv, VALID, top aligned
                else t.tap_cols[tap] = L.same ? tap - L.kw / 2 : tap;   // freq conv: SAME is centred, VALID starts at 0
            }
            t.leaky = L.leaky;
            const int n_chunks = n_p / 8;
            for (int q = 0; q < 4; ++q)
                for (int o = 0; o < 4; ++o) t.chunk_of[q][o] = -1;
            for (int j = 0; j < n_chunks; ++j) {
                if (p.spread) t.chunk_of[j % 4][j / 4] = (signed char)j;
                else t.chunk_of[j / 4][j % 4] = (signed char)j;
            }
            t.w_rows = p.spread ? 128 : n_p;
            while ((packed.size() * 2) % 128) packed.push_back(__float2half_rn(0.f));
            t.w_off = (int64_t)packed.size() * 2;
            t.bias_off = (int)bias.size();
            // A operand, K-major no-swizzle canonical layout: [tap][k chunk][row][8 k values]
            const float* w = weights_host + w_off[li];  // [kh][kw][cin][cout]; tap = kh*kw index
            for (int tap = 0; tap < t.ntaps; ++tap)
                for (int kc = 0; kc < cin_p / 8; ++kc)
                    for (int r = 0; r < t.w_rows; ++r) {
                        const int chunk = t.chunk_of[r / 32][(r % 32) / 8];
                        const int co = chunk < 0 ? -1 : chunk * 8 + r % 8;
                        for (int j = 0; j < 8; ++j) {
                            const int ci = kc * 8 + j;
                            float v = 0.f;
                            if (co >= 0 && co < L.cout && ci < L.cin) v = w[((int64_t)tap * L.cin + ci) * L.cout + co];
                            packed.push_back(__float2half_rn(v));
                        }
                    }
            t.w_bytes = (int)((int64_t)packed.size() * 2 - t.w_off);
            for (int n = 0; n < n_p; ++n) bias.push_back(n < L.cout ? weights_host[b_off[li] + n] : 0.f);
            tw->layers[li] = t;
        }
    }
    for (int i = 0; i < 1024; ++i) packed.push_back(__float2half_rn(0.f));  // slack for aliased row reads
    HB_CUDA_OK(cudaMalloc(&tw->w_packed, packed.size() * 2));
    HB_CUDA_OK(cudaMemcpy(tw->w_packed, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMalloc(&tw->bias_packed, bias.size() * sizeof(float)));
    HB_CUDA_OK(cudaMemcpy(tw->bias_packed, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMalloc(&tw->l0, (3 * 24 + 24) * sizeof(float)));
    HB_CUDA_OK(cudaMemcpy(tw->l0, weights_host + w_off[0], (3 * 24 + 24) * sizeof(float), cudaMemcpyHostToDevice));
    m->tc = tw;
    return tcg_prepare(m, weights_host);
}

void tc_release(hb_embed_model* m) {
    tcg_release(m);
    TcWeights* tw = reinterpret_cast<TcWeights*>(m->tc);
    if (!tw) return;
    cudaFree(tw->w_packed);
    cudaFree(tw->bias_packed);
    cudaFree(tw->l0);
    delete tw;
    m->tc = nullptr;
}

// geometry of one block launch
struct TcGeom {
    int T_in, T_pre, T_out;      // input rows, pre-pool output rows, rows the block writes per clip
    int tiles_per_clip, rows_out, Tt, segs, n_nt, P_alloc, ch_alloc, w_buf_bytes, grid, twin, tile_n;
    size_t smem;
};

static size_t tc_smem_bytes(const TcGeom& g, bool first) {
    const int bufs = g.n_nt == 1 ? 1 : 2;   // a single accumulator tile per layer -> in-place activations
    return ((sizeof(TcSmemHeader) + 2 * (size_t)g.P_alloc + 127 + 16) & ~(size_t)127) + (g.twin ? 1 : 2) * (size_t)g.w_buf_bytes +
           bufs * (128 + (size_t)g.ch_alloc * g.P_alloc * 16) + 128 + (first ? (size_t)g.Tt * kMels * sizeof(float) : 0) + 128;
}

static int tc_w_buf_bytes(int b, const TcWeights* tw) {
    const TcBlockPlan& p = kPlans[b];
    int w = 0;
    for (int l = 0; l < p.n_layers; ++l) w = std::max(w, tw->layers[p.first_layer + l].w_bytes);
    return (w + 127) & ~127;
}

// trunk blocks 0..3 for clips of T_in input rows
static TcGeom tc_geometry(int b, int T_in, int B, const TcWeights* tw) {
    const TcBlockPlan& p = kPlans[b];
    TcGeom g;
    g.T_in = T_in;
    g.T_pre = T_in - 4;
    g.T_out = g.T_pre / p.pool_t;
    const int need = ceil_div(g.T_pre, p.pool_t) * p.pool_t;  // cover every pre-pool row (the activation hook returns them all)
    g.tiles_per_clip = std::max(1, ceil_div(need, p.rows_out_max));
    g.rows_out = ceil_div(ceil_div(need, g.tiles_per_clip), p.pool_t) * p.pool_t;
    g.Tt = g.rows_out + 4;
    g.segs = 1;
    const int S = p.F + 1;
    const int P = 1 + g.Tt * S;
    g.tile_n = p.tile_n;
    g.n_nt = ceil_div(P, g.tile_n);
    g.P_alloc = g.n_nt * g.tile_n + 2 * S + 8;
    g.ch_alloc = std::max(p.cin_pad, p.c_pad) / 8;
    g.w_buf_bytes = tc_w_buf_bytes(b, tw);
    g.grid = B * g.tiles_per_clip;
    g.twin = p.twin;
    g.smem = tc_smem_bytes(g, b == 0);
    return g;
}

// tail block: pooled rows per clip = T15 / 2; several whole clips per CTA when they fit one tile, else time tiles
static TcGeom tc_tail_geometry(int T15, int B, const TcWeights* tw, bool one_clip_per_cta = false) {
    const TcBlockPlan& p = kPlans[4];
    TcGeom g;
    g.T_in = T15;
    const int rows = T15 / 2;              // pooled rows per clip (phase 0; phase 1 may have one fewer valid)
    g.T_pre = rows - 4;
    g.T_out = g.T_pre;                     // valid output rows per clip (two VALID time convs)
    const int S = p.F + 1;
    g.tile_n = p.tile_n;
    const int max_rows = std::min(63, (g.tile_n - 1) / S);    // rows that fit one accumulator tile (and the 6-bit row code)
    if (rows <= max_rows) {
        g.tiles_per_clip = 1;
        g.rows_out = g.T_out;
        g.Tt = rows;
        g.segs = one_clip_per_cta ? 1 : std::max(1, (g.tile_n - 1) / (g.Tt * S));   // the parity hook dumps one clip per CTA
        g.grid = ceil_div(B, g.segs);
    } else {
        g.segs = 1;
        const int per = max_rows - 4;
        g.tiles_per_clip = ceil_div(g.T_out, per);
        g.rows_out = ceil_div(g.T_out, g.tiles_per_clip);
        g.Tt = g.rows_out + 4;
        g.grid = B * g.tiles_per_clip;
    }
    const int P = 1 + g.segs * g.Tt * S;
    g.n_nt = ceil_div(P, g.tile_n);
    g.P_alloc = g.n_nt * g.tile_n + 2 * S + 8;
    g.ch_alloc = p.c_pad / 8;
    g.w_buf_bytes = tc_w_buf_bytes(4, tw);
    g.twin = p.twin;
    g.smem = tc_smem_bytes(g, false);
    return g;
}

static void tc_chain(int F, int B, const TcWeights* tw, TcGeom g[4]) {
    int T = F;
    for (int b = 0; b < 4; ++b) {
        g[b] = tc_geometry(b, T, B, tw);
        T = g[b].T_out;
    }
}

static int64_t block_out_halves(int b, const TcGeom& g) {
    const TcBlockPlan& p = kPlans[b];
    return (int64_t)(p.c_pad / 8) * g.T_out * (p.F / p.pool_f) * 8;
}

static int64_t tc_act_bytes(int B, const TcGeom g[4]) {
    int64_t max_halves = 0;
    for (int b = 0; b < 4; ++b) max_halves = std::max(max_halves, block_out_halves(b, g[b]));
    return ((int64_t)B * max_halves * 2 + 255) & ~255ll;
}

int64_t tc_workspace_bytes(int B, int F) {
    // two fp16 ping-pong activation buffers + two f32 tail outputs + slot table
    if (F < kEmbWindow || B <= 0) return 4096;
    TcWeights dummy;   // geometry only needs the per-layer byte counts; recompute them from the layer table
    for (int b = 0; b < 5; ++b)
        for (int li = kPlans[b].first_layer; li < kPlans[b].first_layer + kPlans[b].n_layers; ++li) {
            const ConvLayer& L = kLayers[li];
            dummy.layers[li].w_bytes = L.kh * L.kw * pad_c(L.cin) * (kPlans[b].spread ? 128 : pad_c(L.cout)) * 2;
        }
    TcGeom g[4];
    tc_chain(F, B, &dummy, g);
    const int64_t tail_rows = std::max(1, g[3].T_out / 2);
    return 2 * tc_act_bytes(B, g) + 2 * (((int64_t)B * tail_rows * kEmbDim * 4 + 255) & ~255ll) + 8192;
}

static int launch_block(const hb_embed_model* m, int b, const TcGeom& g, const void* in, void* out, int B, int in_chunks, int in_F,
                        int pool_phase, __half* dbg, int dbg_layer, cudaStream_t st) {
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
    a.in_mode = b == 0 ? 0 : (b == 4 ? 2 : 1);
    a.out_mode = b == 4 ? 1 : 0;
    a.n_clips = B;
    a.in_T = g.T_in;
    a.in_F = in_F;
    a.in_chunks = in_chunks;
    a.T_out = g.T_out;
    a.F = p.F;
    a.S = p.F + 1;
    a.pool_t = p.pool_t;
    a.pool_f = p.pool_f;
    a.pool_phase = pool_phase;
    a.tiles_per_clip = g.tiles_per_clip;
    a.rows_out = g.rows_out;
    a.Tt = g.Tt;
    a.segs = g.segs;
    a.n_nt = g.n_nt;
    a.P_alloc = g.P_alloc;
    a.n_layers = p.n_layers;
    a.ch_alloc = g.ch_alloc;
    a.w_buf_bytes = g.w_buf_bytes;
    a.tmem_cols = g.twin ? 256 : 512;
    a.n_slots = a.tmem_cols / g.tile_n;
    a.w_double = g.twin ? 0 : 1;
    a.in_place = g.n_nt == 1 ? 1 : 0;
    for (int l = 0; l < p.n_layers; ++l) a.layers[l] = tw->layers[p.first_layer + l];
    HB_REQUIRE(g.smem <= (size_t)(g.twin ? 113 : 227) * 1024, "tc block %d needs %zu bytes of shared memory", b, g.smem);
    HB_REQUIRE(g.P_alloc < 16383 && g.n_nt * g.tile_n < 32000, "tc block %d: tile too large", b);
    HB_REQUIRE(!(g.tile_n == 128 && !g.twin), "tc block %d: 128-position tiles are built for the twin shape only", b);
    static bool configured = false;
    if (!configured) {
        HB_CUDA_OK(cudaFuncSetAttribute(tc_block_kernel<576, 1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        HB_CUDA_OK(cudaFuncSetAttribute(tc_block_kernel<320, 2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
        HB_CUDA_OK(cudaFuncSetAttribute(tc_block_kernel<320, 2, 256>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        HB_CUDA_OK(cudaFuncSetAttribute(tc_block_kernel<320, 2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
        HB_CUDA_OK(cudaFuncSetAttribute(tc_block_kernel<320, 2, 128>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    if (g.twin && g.tile_n == 128) tc_block_kernel<320, 2, 128><<<g.grid, 320, g.smem, st>>>(a);
    else if (g.twin) tc_block_kernel<320, 2, 256><<<g.grid, 320, g.smem, st>>>(a);
    else tc_block_kernel<576, 1, 256><<<g.grid, 576, g.smem, st>>>(a);
    HB_LAUNCHED();
    return HB_OK;
}

static int check_timeout() {
    unsigned int flag = 0;
    HB_CUDA_OK(cudaMemcpyFromSymbol(&flag, g_tc_timeout, sizeof(flag)));
    HB_REQUIRE(flag == 0, "tcgen05 embed kernel: an mbarrier wait timed out (pipeline bug)");
    return tcg_check_timeout();
}

int tc_embed_clips(const hb_embed_model* m, const float* mel, int B, int F, const int32_t* slot_offsets_host, int n_slots,
                   float* out, void* ws, int64_t ws_bytes, cudaStream_t st) {
    const TcWeights* tw = reinterpret_cast<const TcWeights*>(m->tc);
    HB_REQUIRE(tw != nullptr, "tc weights missing");
    HB_REQUIRE(ws_bytes >= tc_workspace_bytes(B, F), "hb_embed: workspace too small");
    TcGeom g[4];
    tc_chain(F, B, tw, g);
    const int64_t act_bytes = tc_act_bytes(B, g);
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    __half* hA = reinterpret_cast<__half*>(base);
    __half* hB = reinterpret_cast<__half*>(base + act_bytes);
    const int T15 = g[3].T_out;
    const TcGeom gt = tc_tail_geometry(T15, B, tw);
    const int64_t tail_bytes = (((int64_t)B * std::max(1, T15 / 2) * kEmbDim * 4) + 255) & ~255ll;
    float* tmp[2] = {reinterpret_cast<float*>(base + 2 * act_bytes), reinterpret_cast<float*>(base + 2 * act_bytes + tail_bytes)};
    int* slot_m_dev = reinterpret_cast<int*>(base + 2 * act_bytes + 2 * tail_bytes);

    bool need_phase[2] = {false, false};
    std::vector<int> slot_m(n_slots);
    HB_REQUIRE(n_slots <= 1024, "hb_embed_clips: too many slots (%d)", n_slots);
    const int J[2] = {T15 / 2 - 4, (T15 - 1) / 2 - 4};
    for (int s = 0; s < n_slots; ++s) {
        slot_m[s] = slot_offsets_host[s] / 4;
        const int p = slot_m[s] & 1, j = slot_m[s] >> 1;
        need_phase[p] = true;
        HB_REQUIRE(j < J[p], "hb_embed_clips: slot %d (offset %d) beyond the strip (phase %d has %d outputs)", s,
                   slot_offsets_host[s], p, J[p]);
    }
    HB_CUDA_OK(cudaMemcpyAsync(slot_m_dev, slot_m.data(), n_slots * sizeof(int), cudaMemcpyHostToDevice, st));

    int rc;
    if ((rc = tcg_block1(m, mel, hA, B, F, nullptr, -1, st))) return rc;
    if ((rc = tcg_block2(m, hA, hB, B, g[1].T_in, nullptr, -1, st))) return rc;
    if ((rc = tcg_block3(m, hB, hA, B, g[2].T_in, nullptr, -1, st))) return rc;
    if ((rc = tcg_block4(m, hA, hB, B, g[3].T_in, nullptr, -1, st))) return rc;
    for (int p = 0; p < 2; ++p)
        if (need_phase[p] && (rc = launch_block(m, 4, gt, hB, tmp[p], B, 12, 4, p, nullptr, -1, st))) return rc;
    return fp32_gather_slots(tmp[0], tmp[1], gt.T_out, gt.T_out, slot_m_dev, n_slots, out, B, st);
}

int64_t tc_activation(const hb_embed_model* m, const float* mel, int B, int F, int layer, float* out, int64_t cap,
                      void* ws, int64_t ws_bytes, cudaStream_t st) {
    const TcWeights* tw = reinterpret_cast<const TcWeights*>(m->tc);
    if (!tw || ws_bytes < tc_workspace_bytes(B, F)) {
        set_error("hb_embed_activation(f16): bad workspace");
        return HB_ERR_INVALID;
    }
    TcGeom g[4];
    tc_chain(F, B, tw, g);
    const int64_t act_bytes = tc_act_bytes(B, g);
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    __half* bufs[2] = {reinterpret_cast<__half*>(base), reinterpret_cast<__half*>(base + act_bytes)};
    if (layer > 15) {
        // conv2d_16..19 = the tail block on pool phase 0 of conv2d_15's output (what spec.embedding_layer_shapes describes), one clip per
        // CTA so that the block's own activation dump maps back to [clip][row][96]
        int rc;
        if ((rc = tcg_block1(m, mel, bufs[0], B, F, nullptr, -1, st))) return rc;
        if ((rc = tcg_block2(m, bufs[0], bufs[1], B, g[1].T_in, nullptr, -1, st))) return rc;
        if ((rc = tcg_block3(m, bufs[1], bufs[0], B, g[2].T_in, nullptr, -1, st))) return rc;
        if ((rc = tcg_block4(m, bufs[0], bufs[1], B, g[3].T_in, nullptr, -1, st))) return rc;
        const int T15 = g[3].T_out;
        const TcGeom gt = tc_tail_geometry(T15, B, tw, true);
        if (gt.tiles_per_clip != 1) { set_error("hb_embed_activation(f16): conv2d_16..19 need a strip whose pooled rows fit one tile"); return HB_ERR_UNSUPPORTED; }
        float* tail_out = reinterpret_cast<float*>(base + 2 * act_bytes);
        const int rows = T15 / 2;
        const int T = layer == 16 ? rows : (layer == 19 ? rows - 4 : rows - 2);
        const int64_t n = (int64_t)B * T * kEmbDim;
        if (T <= 0 || n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
        void* dbg_mem = nullptr;
        if (layer < 19) {
            const int64_t need = (int64_t)gt.grid * gt.ch_alloc * gt.P_alloc * 16;
            if (cudaMalloc(&dbg_mem, need) != cudaSuccess) { set_error("hb_embed_activation: debug allocation failed"); return HB_ERR_CUDA; }
        }
        rc = launch_block(m, 4, gt, bufs[1], tail_out, B, 12, 4, 0, reinterpret_cast<__half*>(dbg_mem), layer < 19 ? layer - 16 : -1, st);
        if (rc) { if (dbg_mem) cudaFree(dbg_mem); return rc; }
        if (layer < 19) {
            const int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), 148 * 8);
            dump_to_nhwc_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<__half*>(dbg_mem), out, B, 1, gt.ch_alloc, gt.P_alloc, kPlans[4].F + 1, 1, gt.Tt, T, kEmbDim);
        } else if (cudaMemcpyAsync(out, tail_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
            set_error("hb_embed_activation: copy failed");
            return HB_ERR_CUDA;
        }
        const bool ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
        if (dbg_mem) cudaFree(dbg_mem);
        if (!ok) { set_error("hb_embed_activation: kernel failed"); return HB_ERR_CUDA; }
        if (check_timeout() != HB_OK) return HB_ERR_CUDA;
        return n;
    }
    const int target_block = layer < 4 ? 0 : (layer < 8 ? 1 : (layer < 12 ? 2 : 3));
    static const int in_chunks[4] = {0, 4, 6, 10}, in_F[4] = {kMels, 16, 8, 4};
    const void* in = mel;
    int which = 0;
    int64_t result = HB_ERR_INVALID;
    void* dbg_mem = nullptr;
    for (int b = 0; b <= target_block; ++b) {
        const TcBlockPlan& p = kPlans[b];
        const bool last = (b == target_block);
        const int last_layer_of_block = p.first_layer + p.n_layers - 1;
        int dbg_layer = -1;
        if (last && layer != last_layer_of_block) dbg_layer = (layer == 0) ? 100 : layer - p.first_layer;
        if (b <= 3) {
            // blocks 1-4 run on the grouped kernels (embed_tcg.cu), which dump f32 NHWC directly
            const __half* in_h = reinterpret_cast<const __half*>(in);
            if (dbg_layer >= 0) {
                int t_convs = 0;
                for (int li = (b == 0 ? 0 : p.first_layer); li <= layer; ++li) t_convs += (kLayers[li].kh == 3);
                const int T = g[b].T_in - 2 * t_convs;
                const int64_t n = (int64_t)B * T * p.F * kLayers[layer].cout;
                if (n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
                int rc = b == 0 ? tcg_block1(m, mel, bufs[which], B, F, out, dbg_layer, st)
                       : b == 1 ? tcg_block2(m, in_h, bufs[which], B, g[1].T_in, out, dbg_layer, st)
                       : b == 2 ? tcg_block3(m, in_h, bufs[which], B, g[2].T_in, out, dbg_layer, st)
                                : tcg_block4(m, in_h, bufs[which], B, g[3].T_in, out, dbg_layer, st);
                if (rc) return rc;
                if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("hb_embed_activation: kernel failed"); return HB_ERR_CUDA; }
                if (check_timeout() != HB_OK) return HB_ERR_CUDA;
                return n;
            }
            int rc = b == 0 ? tcg_block1(m, mel, bufs[which], B, F, nullptr, -1, st)
                   : b == 1 ? tcg_block2(m, in_h, bufs[which], B, g[1].T_in, nullptr, -1, st)
                   : b == 2 ? tcg_block3(m, in_h, bufs[which], B, g[2].T_in, nullptr, -1, st)
                            : tcg_block4(m, in_h, bufs[which], B, g[3].T_in, nullptr, -1, st);
            if (rc) return rc;
        } else if (dbg_layer >= 0) {
            const int64_t need = (int64_t)g[b].grid * g[b].ch_alloc * g[b].P_alloc * 16;
            if (cudaMalloc(&dbg_mem, need) != cudaSuccess) { set_error("hb_embed_activation: debug allocation failed"); return HB_ERR_CUDA; }
        }
        if (b > 3) {
            int rc = launch_block(m, b, g[b], in, bufs[which], B, in_chunks[b], in_F[b], 0, reinterpret_cast<__half*>(dbg_mem), dbg_layer, st);
            if (rc) { if (dbg_mem) cudaFree(dbg_mem); return rc; }
        }
        if (last) {
            int T, Fq, C;
            if (dbg_layer >= 0) {
                int t_convs = 0;   // time convs done so far inside the block shrink the valid rows
                for (int li = (b == 0 ? 0 : p.first_layer); li <= layer; ++li) t_convs += (kLayers[li].kh == 3);
                T = g[b].T_in - 2 * t_convs;
                Fq = p.F;
                C = kLayers[layer].cout;
                const int64_t n = (int64_t)B * T * Fq * C;
                if (n > cap) { set_error("hb_embed_activation: output capacity too small"); cudaFree(dbg_mem); return HB_ERR_INVALID; }
                const int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), 148 * 8);
                dump_to_nhwc_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<__half*>(dbg_mem), out, B, g[b].tiles_per_clip, g[b].ch_alloc,
                                                            g[b].P_alloc, p.F + 1, Fq, g[b].rows_out, T, C);
                result = n;
            } else if (b == 3) {
                // conv2d_15 "after its pool" = 2x2 pool phase 0 of the pre-pool output
                T = g[b].T_out;
                const int64_t n_pre = (int64_t)B * T * 4 * kEmbDim;
                if (cudaMalloc(&dbg_mem, n_pre * 4) != cudaSuccess) { set_error("hb_embed_activation: allocation failed"); return HB_ERR_CUDA; }
                float* pre = reinterpret_cast<float*>(dbg_mem);
                const int blocks = (int)std::min<int64_t>(ceil_div64(n_pre, 256), 148 * 8);
                chunked_to_nhwc_kernel<<<blocks, 256, 0, st>>>(bufs[which], pre, B, 12, T, 4, kEmbDim);
                const int64_t n = (int64_t)B * (T / 2) * 2 * kEmbDim;
                if (n > cap) { set_error("hb_embed_activation: output capacity too small"); cudaFree(dbg_mem); return HB_ERR_INVALID; }
                int rc2 = fp32_pool_public(pre, out, B, T, 4, kEmbDim, 2, 2, 0, st);
                if (rc2) { cudaFree(dbg_mem); return rc2; }
                result = n;
            } else {
                T = g[b].T_out;
                Fq = p.F / p.pool_f;
                C = p.c_real;
                const int64_t n = (int64_t)B * T * Fq * C;
                if (n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
                const int blocks = (int)std::min<int64_t>(ceil_div64(n, 256), 148 * 8);
                chunked_to_nhwc_kernel<<<blocks, 256, 0, st>>>(bufs[which], out, B, p.c_pad / 8, T, Fq, C);
                result = n;
            }
            const bool ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
            if (dbg_mem) cudaFree(dbg_mem);
            if (!ok) { set_error("hb_embed_activation: kernel failed"); return HB_ERR_CUDA; }
            if (check_timeout() != HB_OK) return HB_ERR_CUDA;
            return result;
        }
        in = bufs[which];
        which ^= 1;
    }
    return result;
}

}  // namespace hb

// profiling aid (not part of the public header): phase timestamps of the first 8 CTAs of the last tc block launch
#ifdef HB_TC_FINE
extern "C" int hb_debug_tc_fine(long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, hb::g_tc_fine, sizeof(long long) * 8 * 48) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int hb_debug_tcg_times(long long* out_host) { return hb::tcg_debug_times(out_host); }

extern "C" int hb_debug_tc_times(long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, hb::g_tc_times, sizeof(long long) * 8 * 16) == cudaSuccess ? 0 : -2;
}
