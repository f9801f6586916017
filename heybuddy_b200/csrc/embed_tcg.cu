// K7, blocks 1-4 of the speech-embedding conv stack (conv2d .. conv2d_15, reference embeddings.py:32-42 = one ONNX session
// run): tcgen05 implicit GEMMs with G output positions per accumulator column.
//
// An SS tcgen05.mma with K = 16 costs N / 2 cycles whatever M is (scripts/micro/umma_swizzle.cu), so the short dimension
// (24 .. 96 output channels) sits in M and positions in N.  With one position per column a 24-channel layer would fill 24 of
// the 128 M rows; here M carries (sub-position i = 0..G-1, cout) = up to 96 rows: a column is a group of G adjacent positions
// ALONG THE CONV AXIS, K runs over the G + 2 input positions the group touches (block 1: 6 x 24 = 144 = nine K = 16 steps)
// and the A operand is the banded Toeplitz matrix of the 3-tap kernel (zeros where a tap does not connect).  Block 1: 9 MMAs
// per 1024 positions instead of 24; block 2 (G = 2): 12 per 512 instead of 18; blocks 3 and 4 have G = 1.
//
// Layouts (all [plane][column][8 x fp16], 16-byte records, one plane per (phase, 8-channel chunk); FG = F / G):
//   T (input of a time conv):  plane (t mod G, chunk), column (t div G) * F + pi(f),  pi(f) = (f mod G) * FG + f div G
//   F (input of a freq conv):  plane (f mod G, chunk), column 1 + (FG + 1) t + f div G  (group FG of every row and column 0
//                              are the zero SAME padding)
//   P (pre-pool output):       [chunk][t * F + pi(f)]
// A K step is two K chunks = two plane addresses: the UMMA descriptor's (start, LBO) pair expresses any of them, so there is
// no im2col.  The tile's single activation buffer is rewritten in place: a layer has one accumulator tile (N <= 256 columns),
// so all of its MMAs have completed (tcgen05.commit) before the first epilogue store.
//
// conv2d (Cin = 1) is layer 0 of block 1: the f32 mel value enters as one input chunk whose channels 0 / 1 are its fp16
// hi / lo halves, the A operand carries conv2d's weight in both rows.  The bias of every layer rides the MMA as an extra
// "ones" K chunk (fp16 hi + lo pair).
//
// Tile = 28 / 26 / 24 / 32 input rows of one clip (blocks 1-4); 320 threads, 2 persistent CTAs per SM, 256 TMEM columns each.
// Warp 0 issues MMAs, warp 1 streams the next layer's weights as soon as the current MMAs have completed and zeroes the SAME
// padding, warps 2-9 run the epilogue (TMEM lane quadrant q holds row chunks 3 q .. 3 q + 2 of the 12 (sub-position,
// 8-channel chunk) row chunks): tcgen05.ld -> LeakyReLU on the packed fp16 pair -> stmatrix.trans through a host-built
// scatter table straight into the next layer's operand layout.
//
// NaN note: the zero Toeplitz entries multiply neighbouring positions of the same tile, so a non-finite activation reaches
// up to G + 1 more positions of its own clip than in the reference arithmetic (0 * inf); finite data is unaffected.
#include "tc_ptx.cuh"

#include <vector>

namespace hb {

namespace {

#ifndef HB_TCG_PLANE_SKEW
#define HB_TCG_PLANE_SKEW 0    // experiments: 32 = the old bank skew between planes
#endif
#ifndef HB_TCG_INTERLEAVE
#define HB_TCG_INTERLEAVE 1   // narrow-plane configuration (block 4): interleave the epilogue warps' column blocks
#endif
#ifndef HB_TCG_WIDE_CTAS
#define HB_TCG_WIDE_CTAS 2      // CTAs per SM of the 16-epilogue-warp shape (1 lifts its 56-register cap; experiments only)
#endif
#ifndef HB_TCG_THREADS23
#define HB_TCG_THREADS23 320   // launch shape of blocks 2 and 3 (320 = 8 epilogue warps, 576 = 16)
#endif
constexpr int kGPlanes = 12;        // planes allocated (G * channel chunks <= 12)
constexpr int kGTmemCols = 256;
constexpr int kGMaxLayers = 4;
// Epilogue scatter table of a layer: entry (n & 15) * kGTabRow + (n >> 4) = byte offset >> 4 of accumulator column n in the layer's
// target layout.  The four columns n, n + 16, n + 32, n + 48 one lane stores from a 64-column fragment are one 8-byte load; a row
// pitch of 20 entries (40 B) keeps the 16 rows a warp reads in different banks.
constexpr int kGTabRow = 20;

// G positions per column, CC 8-channel chunks (G * CC = 12 row chunks = 96 M rows), F freq bins (F / G = 8), TT input rows
// per tile (multiple of G), NL tensor-core layers alternating freq / time, CIN0 input chunks of the first layer.
template <int G_, int CC_, int F_, int TT_, int NL_, bool FIRST_FREQ_, int CIN0_, int POOL_T_, bool MEL_IN_, int OUT_CH_, int CONV0_, int THREADS_,
          int POOL_F_ = 2, int PCOLS_ = 256>
struct GCfg {
    // launch shape: warp 0 MMA issuer, warp 1 weight loader, then 8 or 16 epilogue warps (2 or 4 per TMEM lane quadrant, each
    // owning 128 or 64 accumulator columns = two fragments of SUBC columns per 16-lane half)
    static constexpr int THREADS = THREADS_, EPI_WARPS = THREADS_ / 32 - 2, PARTS = EPI_WARPS / 4, SUBC = 128 / PARTS;
    static_assert(THREADS_ % 32 == 0 && (EPI_WARPS == 8 || EPI_WARPS == 16), "launch shape");
    static constexpr int G = G_, CC = CC_, F = F_, TT = TT_, NL = NL_, CIN0 = CIN0_, POOL_T = POOL_T_, OUT_CH = OUT_CH_;
    static constexpr int POOL_F = POOL_F_, PCOLS = PCOLS_;
    static constexpr int FG = F_ / G_;                         // column groups per row (a freq conv's row has FG + 1 columns: + the zero pad)
    // bytes per plane: PCOLS columns.  With FG = 8 the pitch is a multiple of 128 B: consecutive rows of layout T then start in
    // the same bank, which is what makes the freq layers' row-straddling stmatrix stores conflict-free (tcg_tables).
    static constexpr int PLANE = PCOLS_ * 16 + HB_TCG_PLANE_SKEW;
    static constexpr int ACT_BYTES = kGPlanes * PLANE;
    static constexpr bool FIRST_FREQ = FIRST_FREQ_, MEL_IN = MEL_IN_;
    static constexpr int CONV0 = CONV0_;                       // conv index (layer table) of the first tensor-core layer
    static constexpr int C = CC * 8;
    static constexpr int N_TIME = FIRST_FREQ ? NL / 2 : (NL + 1) / 2;
    static constexpr int ROWS_OUT = TT - 2 * N_TIME;
    static constexpr int W_MAX = (((G + 2) * CC + 2) & ~1) * 2048;   // largest A operand: [K chunk][row 128][8]
    static constexpr int PLAIN = TT * F * 16;                  // layout P bytes per chunk
    static constexpr int RC = G * CC;                          // row chunks (sub-position, channel chunk); rows = 8 RC <= 96
    static_assert(RC <= kGPlanes && F % G == 0 && TT % G == 0 && (TT / G) * F <= 256 && TT * (FG + 1) <= 256, "tile shape");
    static_assert(TT * (FG + 1) + 2 <= PCOLS_ && (TT / G) * F + (G > 1 ? F : 2 * F) <= PCOLS_, "plane too small for the MMAs' reads");
    static_assert(CC * PLAIN <= ACT_BYTES && ROWS_OUT % POOL_T == 0, "tile shape");
    __host__ __device__ static constexpr bool is_freq(int l) { return ((l & 1) == 0) == FIRST_FREQ; }
    __host__ __device__ static constexpr int cin(int l) { return l == 0 ? CIN0 : CC; }
    __host__ __device__ static constexpr int kreal(int l) { return (G + 2) * cin(l); }            // K chunks of the conv itself
    // + one "ones" chunk that carries the bias (A row = [bias_hi, bias_lo, 0..], B record = [1, 1, 0..]), padded to whole K = 16 steps
    __host__ __device__ static constexpr int kchunks(int l) { return (kreal(l) + 2) & ~1; }
    // the constant ones plane: a spare plane when the block uses fewer than 12, else right after the buffer's guard + dump slots
    static constexpr int ONES_OFF = RC < kGPlanes ? RC * PLANE : ACT_BYTES + 512 + 128;
    static constexpr int EXTRA_SMEM = RC < kGPlanes ? 0 : PLANE + 64;
    __host__ __device__ static constexpr int w_bytes(int l) { return kchunks(l) * 2048; }
    __host__ __device__ static constexpr int w_off(int l) { return l == 0 ? 0 : w_off(l - 1) + w_bytes(l - 1); }
    __host__ __device__ static constexpr int pi(int f) { return (f % G) * FG + f / G; }
    // accumulator columns a layer really has (a multiple of 16): an MMA costs N / 2 cycles, so N is not rounded up to 256
    __host__ __device__ static constexpr int ncols(int l) { return ((is_freq(l) ? TT * (FG + 1) : (TT / G) * F) + 15) & ~15; }
    // K chunk kk = c * (G + 2) + ord (channel chunk major).  ord -> input offset d within the group: time d = dt = ord;
    // freq d = df: 0 .. G, -1 (G >= 2) or -1, 0, 1 (G = 1) -- orders that keep every K-step pair address-ordered.
    __host__ __device__ static constexpr int kd(int l, int kk) {
        const int ord = kk % (G + 2);
        return !is_freq(l) ? ord : (G == 1 ? ord - 1 : (ord <= G ? ord : -1));
    }
    // byte offset (within the activation buffer, relative to column n = 0) of K chunk kk of layer l
    __host__ __device__ static constexpr uint32_t koff(int l, int kk) {
        if (kk >= kreal(l)) return (uint32_t)(ONES_OFF + (kk - kreal(l)) * 16);   // ones chunk, then a zero-weight pad chunk
        const int ci = cin(l), c = kk / (G + 2), d = kd(l, kk);
        if (is_freq(l)) {
            const int fm = ((d % G) + G) % G;
            const int sh = d < 0 ? 0 : (d >= G ? 2 : 1);        // 1 + column shift (column 0 is the guard)
            return (uint32_t)((fm * ci + c) * PLANE + sh * 16);
        }
        return (uint32_t)(((d % G) * ci + c) * PLANE + (d / G) * F * 16);
    }
};
using Cfg1 = GCfg<4, 3, 32, 28, 4, true, 1, 2, true, 4, 0, 320>;     // conv2d (mel as an fp16 hi + lo pair) .. conv2d_3, pool 2x2
using Cfg2 = GCfg<2, 6, 16, 26, 4, true, 3, 1, false, 6, 4, HB_TCG_THREADS23>;    // conv2d_4..7, pool 1x2
using Cfg3 = GCfg<1, 9, 8, 24, 4, true, 6, 2, false, 10, 8, HB_TCG_THREADS23>;
// conv2d_12..15, no pool (the tail pools with a time phase): the whole clip (<= 32 rows) is one tile; 162-column planes keep two CTAs per SM
using Cfg4 = GCfg<1, 12, 4, 32, 4, true, 9, 1, false, 12, 12, 320, 1, 162>;    // conv2d_8..11, pool 2x2 (one position per column, N = 256)

struct GArgs {
    const void* in;               // MEL_IN: mel f32 [clips][in_T][32]; else fp16 chunk-major [clips][in_chunks][in_T][F][8]
    __half* out;                  // fp16 chunk-major [clips][OUT_CH][T_out][F / 2][8]
    const unsigned char* w;       // packed A operands of the NL layers
    const float* bias;            // NL x C
    const float* l0_w;            // MEL_IN: conv2d kernel f32 [3][24] + bias [24]
    const uint16_t* tab;          // [NL][16 * kGTabRow] epilogue scatter tables (tcg_tables)
    __half* pool2_out;            // block 4 only, optional: the tail's input instead of `out` -- conv2d_15's 2x2 pool for BOTH time
                                  // phases, fp16 [(clip, phase, row)][bin 0..1][12 chunks][8] (embed_tail.cu); one tile per clip
    float* dbg;                   // optional f32 NHWC [clips][dbg_T][F][C] activation dump
    int dbg_layer;                // -1 none, 100 = staged input (MEL_IN: conv2d output), l = output of tensor-core layer l
    int dbg_T;
    int n_clips, in_T, in_chunks, T_out, tiles_per_clip;
};

template <class Cfg>
struct GSmemHeader {
    uint64_t tmem_full, tmem_empty, wbar;
    uint32_t tmem_base;
    uint32_t pad[1];
    alignas(16) uint16_t tab[kGMaxLayers][kGTabRow * 16];   // epilogue scatter tables (tcg_tables), per layer
};

// LeakyReLU + fp16 + stmatrix of one 16-lane x (8 NR)-column fragment (already in registers; the bias came through the MMA).
// tb = the lane's scatter-table entries (columns n, n + 16, ..; 16 bits each); 16-column groups at or past n_valid
// (warp-uniform) were never written by the layer's MMAs and are skipped.
template <bool kTwo, int NR>
__device__ __forceinline__ void g_epilogue(const uint32_t (&r)[NR], uint2 tb, uint32_t base_lane, int col0, int n_valid) {
#pragma unroll
    for (int gg = 0; gg < NR / 8; ++gg) {
        if (col0 + 16 * gg >= n_valid) break;
        const uint32_t w = gg < 2 ? tb.x : tb.y;
        const uint32_t e = (gg & 1) ? (w >> 16) : (w & 0xffffu);
        const uint32_t addr = base_lane + (e << 4);
        const uint32_t ra = leaky_half2(__uint_as_float(r[8 * gg + 0]), __uint_as_float(r[8 * gg + 1]));
        const uint32_t rc = leaky_half2(__uint_as_float(r[8 * gg + 4]), __uint_as_float(r[8 * gg + 5]));
        if (kTwo) {
            const uint32_t rb = leaky_half2(__uint_as_float(r[8 * gg + 2]), __uint_as_float(r[8 * gg + 3]));
            const uint32_t rd = leaky_half2(__uint_as_float(r[8 * gg + 6]), __uint_as_float(r[8 * gg + 7]));
            stmatrix_x4_trans(addr, ra, rb, rc, rd);
        } else {
            stmatrix_x2_trans(addr, ra, rc);
        }
    }
}

// All MMAs of layer l: descriptor = base + compile-time immediates (start address >> 4 in bits [0,14), LBO >> 4 in [16,30)).
template <class Cfg, int L, int J>
__device__ __forceinline__ void issue_steps(uint32_t d_tmem, uint64_t a_base, uint64_t b_base, uint32_t idesc) {
    if constexpr (J < Cfg::kchunks(L) / 2) {
        constexpr uint32_t idesc_l = make_idesc(128, Cfg::ncols(L));
        constexpr uint32_t o0 = Cfg::koff(L, 2 * J), o1 = Cfg::koff(L, 2 * J + 1);
        static_assert(o1 > o0 && ((o1 - o0) >> 4) < 0x4000, "K chunk pair must be address-ordered");
        constexpr uint64_t a_inc = (uint64_t)((2 * J * 2048) >> 4);
        constexpr uint64_t b_inc = (uint64_t)(o0 >> 4) | ((uint64_t)((o1 - o0) >> 4) << 16);
        umma_f16(d_tmem, a_base + a_inc, b_base + b_inc, idesc_l, J > 0 ? 1u : 0u);
        issue_steps<Cfg, L, J + 1>(d_tmem, a_base, b_base, idesc);
    }
}
template <class Cfg>
__device__ __forceinline__ void issue_layer(int l, uint32_t d_tmem, uint64_t a_base, uint64_t b_base, uint32_t idesc) {
    if (l == 0) issue_steps<Cfg, 0, 0>(d_tmem, a_base, b_base, idesc);
    else if (l == 1) issue_steps<Cfg, 1, 0>(d_tmem, a_base, b_base, idesc);
    else if (l == 2) issue_steps<Cfg, 2, 0>(d_tmem, a_base, b_base, idesc);
    else if (Cfg::NL > 3) issue_steps<Cfg, (Cfg::NL > 3 ? 3 : 0), 0>(d_tmem, a_base, b_base, idesc);
}

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::THREADS > 320 ? HB_TCG_WIDE_CTAS : 2) tcg_block_kernel(const GArgs a) {
    constexpr int kGThreads = Cfg::THREADS, kGEpiWarps = Cfg::EPI_WARPS;
    constexpr int G = Cfg::G, CC = Cfg::CC, F = Cfg::F, TT = Cfg::TT, NL = Cfg::NL, C = Cfg::C, FG = Cfg::FG;
    constexpr int kGPlane = Cfg::PLANE, kGActBytes = Cfg::ACT_BYTES;
    extern __shared__ __align__(128) unsigned char smem[];
    GSmemHeader<Cfg>& hdr = *reinterpret_cast<GSmemHeader<Cfg>*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* wbuf = smem + ((sizeof(GSmemHeader<Cfg>) + 127) & ~127);
    unsigned char* act = wbuf + Cfg::W_MAX;
    unsigned char* dump = act + kGActBytes + 512;             // 512 B finite guard for reads past the last plane

    // Persistent CTA: tiles blockIdx.x, blockIdx.x + gridDim.x, ... -- TMEM, barriers, tables and the ones plane are set up
    // once; mbarrier phases simply keep counting across tiles.
    const int n_tiles = a.n_clips * a.tiles_per_clip;

    TC_STAMP(0);
    // MEL_IN: a tile's mel rows start their trip from global memory one tile ahead and wait in registers
    constexpr int kMelPerThread = (TT * kMels + kGThreads - 1) / kGThreads;
    float mel_reg[Cfg::MEL_IN ? kMelPerThread : 1];
    auto mel_bin = [](int i) { return ((i & 7) << 2) | ((i >> 3) & 3); };
    auto prefetch_mel = [&](int tile_id) {
        if (!Cfg::MEL_IN || tile_id >= n_tiles) return;
        const int pc = tile_id / a.tiles_per_clip, pr0 = (tile_id - pc * a.tiles_per_clip) * Cfg::ROWS_OUT;
        const float* mel = reinterpret_cast<const float*>(a.in) + (int64_t)pc * a.in_T * kMels;
#pragma unroll
        for (int k = 0; k < kMelPerThread; ++k) {
            // element (row r, bin mel_bin(i)): a warp still reads one 128-byte row, but its 8-lane groups hold bins 4 j + const,
            // i.e. ONE plane of layout F and 8 consecutive columns -- conflict-free 16-byte staging stores
            const int i = tid + k * kGThreads, r = i / kMels;
            mel_reg[k] = (i < TT * kMels && pr0 + r < a.in_T) ? __ldg(mel + (int64_t)(pr0 + r) * kMels + mel_bin(i)) : 0.f;
        }
    };
    prefetch_mel(blockIdx.x);
    // ---- one-time setup ----------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(&hdr.tmem_full, 1);
        mbar_init(&hdr.tmem_empty, kGEpiWarps + 1);   // the 8 epilogue warps + warp 1 (padding zeroing)
        mbar_init(&hdr.wbar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&hdr.tmem_base, kGTmemCols);
    for (int i = tid; i < NL * kGTabRow * 16 / 8; i += kGThreads)   // epilogue scatter tables, precomputed on the host (tcg_tables)
        reinterpret_cast<uint4*>(&hdr.tab[0][0])[i] = __ldg(reinterpret_cast<const uint4*>(a.tab) + i);
    // everything the MMAs may read must be finite: clear the activation buffer and its guard
    for (int i = tid; i < (kGActBytes + 512) / 16; i += kGThreads) reinterpret_cast<uint4*>(act)[i] = make_uint4(0, 0, 0, 0);
    if (Cfg::RC < kGPlanes) __syncthreads();                  // the ones plane is one of the planes just cleared
    for (int i = tid; i < Cfg::PCOLS + 2; i += kGThreads)     // [1, 1, 0, 0, 0, 0, 0, 0] per column: the bias K chunk's B operand
        *reinterpret_cast<uint4*>(act + Cfg::ONES_OFF + i * 16) = make_uint4(0x3c003c00u, 0, 0, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr.tmem_base;

    if (warp == 1 && lane == 0) {
        mbar_expect_tx(&hdr.wbar, (uint32_t)Cfg::w_bytes(0));
        bulk_g2s(wbuf, a.w, (uint32_t)Cfg::w_bytes(0), &hdr.wbar);
    }
    const uint32_t act_u32 = smem_u32(act), w_u32 = smem_u32(wbuf);

    uint32_t ph = 0;   // layers processed so far by this CTA = phase index of tmem_full / wbar (tmem_empty lags by one)
    for (int tile_id = blockIdx.x; tile_id < n_tiles; tile_id += gridDim.x) {
    const bool first_tile = (tile_id == (int)blockIdx.x);
    // profiling stamps describe the CTA's SECOND tile (steady state: warm caches, the co-resident CTA out of phase); stamp 1 = its start
    const bool stamp_tile = (tile_id == (int)blockIdx.x + (int)gridDim.x);
    if (stamp_tile) TC_STAMP(1);
    const int clip = tile_id / a.tiles_per_clip, tile = tile_id - clip * a.tiles_per_clip;
    const int row0 = tile * Cfg::ROWS_OUT;

    // ---- stage the input tile -----------------------------------------------------------------------------
    static_assert(Cfg::FIRST_FREQ, "blocks start with a freq conv");
    if (!first_tile) {
        // the buffer holds the previous tile's layout P: restore layout F's zero padding (column 0, group 8 of every row)
        for (int i = tid; i < G * Cfg::CIN0 * (TT + 1); i += kGThreads) {
            const int pl = i / (TT + 1), k = i - pl * (TT + 1);
            const int col = k == 0 ? 0 : 1 + (FG + 1) * (k - 1) + FG;
            *reinterpret_cast<uint4*>(act + pl * kGPlane + col * 16) = make_uint4(0, 0, 0, 0);
        }
    }
    if (Cfg::MEL_IN) {
        // mel rows -> layout F with ONE input chunk whose channels 0 / 1 are the fp16 hi / lo halves of the f32 mel value (exact to
        // 2^-22; the A operand carries conv2d's weight in both rows): conv2d (Cin = 1) runs on the tensor core like every other layer
        static_assert(!Cfg::MEL_IN || Cfg::CIN0 == 1, "mel input = one chunk");
#pragma unroll
        for (int k = 0; k < kMelPerThread; ++k) {
            const int i = tid + k * kGThreads;
            if (i < TT * kMels) {
                const int t = i / kMels, f = mel_bin(i);
                const __half hi = __float2half_rn(mel_reg[k]);
                const __half lo = __float2half_rn(mel_reg[k] - __half2float(hi));
                *reinterpret_cast<uint4*>(act + (f % G) * kGPlane + (1 + (FG + 1) * t + f / G) * 16) =
                    make_uint4((uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16), 0, 0, 0);
            }
        }
        prefetch_mel(tile_id + (int)gridDim.x);
    } else {
        // fp16 chunk-major rows [row0, row0 + TT) -> layout F (the first layer is a freq conv), 16-byte cp.async records
        const uint4* in = reinterpret_cast<const uint4*>(a.in) + (int64_t)clip * a.in_chunks * a.in_T * F;
#ifndef HB_TCG_NO_PREFETCH
        // the next tile's rows (written to HBM by the previous block's kernel) start their trip to L2 now: its staging then pays an
        // L2 latency instead of an HBM one
        if (tid < Cfg::CIN0 && tile_id + (int)gridDim.x < n_tiles) {
            const int nt = tile_id + (int)gridDim.x, nc = nt / a.tiles_per_clip, nr0 = (nt - nc * a.tiles_per_clip) * Cfg::ROWS_OUT;
            const int rows = min(TT, a.in_T - nr0);
            if (rows > 0)
                bulk_prefetch_l2(reinterpret_cast<const uint4*>(a.in) + (((int64_t)nc * a.in_chunks + tid) * a.in_T + nr0) * F,
                                 (uint32_t)(rows * F * 16));
        }
#endif
        // Through registers, not cp.async: a 16-byte LDGSTS is one shared-memory wavefront per THREAD (1248 per tile in block 2,
        // a quarter of the kernel's LSU wavefronts), a 16-byte STS is one per 8 lanes.  All loads of a thread are in flight before
        // its first store; with G = 2 an 8-lane group takes the even (or odd) bins of a row = 8 consecutive columns of one plane.
        constexpr int kRec = Cfg::CIN0 * TT * F, kPer = (kRec + kGThreads - 1) / kGThreads;
        uint4 rec[kPer];
        auto place = [&](int i, int& c, int& t, int& f) {
            c = i / (TT * F);
            const int rem = i - c * (TT * F);
            t = rem / F;
            const int fi = rem - t * F;
            f = (G == 2 && F == 16) ? (((fi & 7) << 1) | (fi >> 3)) : fi;
        };
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kGThreads;
            int c, t, f;
            place(i, c, t, f);
            // rows past the clip are zero-filled: they only feed outputs that are dropped, but must be finite
            rec[k] = (i < kRec && row0 + t < a.in_T) ? __ldg(in + ((int64_t)c * a.in_T + row0 + t) * F + f) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int i = tid + k * kGThreads;
            int c, t, f;
            place(i, c, t, f);
            if (i < kRec) *reinterpret_cast<uint4*>(act + ((f % G) * Cfg::CIN0 + c) * kGPlane + (1 + (FG + 1) * t + f / G) * 16) = rec[k];
        }
    }
    fence_proxy_async();   // generic-proxy stores above -> visible to the tensor core's async-proxy reads
    __syncthreads();
    if (stamp_tile) TC_STAMP(2);

    // f32 NHWC dump of the tile's rows from layout T / F (parity hook only)
    auto dump_act = [&](bool layout_f, int cc_planes) {
        for (int i = tid; i < Cfg::ROWS_OUT * F * C; i += kGThreads) {
            const int c = i % C, f = (i / C) % F, t = i / (C * F);
            if (row0 + t >= a.dbg_T || (c >> 3) >= cc_planes) continue;
            const unsigned char* p = layout_f ? act + ((f % G) * cc_planes + (c >> 3)) * kGPlane + (1 + (FG + 1) * t + f / G) * 16
                                              : act + ((t % G) * cc_planes + (c >> 3)) * kGPlane + ((t / G) * F + Cfg::pi(f)) * 16;
            a.dbg[(((int64_t)clip * a.dbg_T + row0 + t) * F + f) * C + c] = __half2float(reinterpret_cast<const __half*>(p)[c & 7]);
        }
    };
    if (a.dbg != nullptr && a.dbg_layer == 100) dump_act(Cfg::FIRST_FREQ, Cfg::CIN0);

    // ---- the block's tensor-core layers ------------------------------------------------------------------------
#pragma unroll
    for (int l = 0; l < NL; ++l, ++ph) {
        constexpr int kDummy = 0; (void)kDummy;
        const bool freq = Cfg::is_freq(l);
        const bool last = (l == NL - 1);
#define TCG_FINE(i) do { if (l == 1 && stamp_tile && lane == 0 && blockIdx.x < 8) g_tc_times[blockIdx.x][i] = clock64(); } while (0)
        if (warp == 0) {
            TCG_FINE(9);
            mbar_wait(&hdr.wbar, ph & 1u, 1u);
            TCG_FINE(10);
            if (ph > 0) mbar_wait(&hdr.tmem_empty, (ph - 1) & 1u, 2u);
            tc_fence_after();
            // One elected lane issues the whole layer; l and j are compile-time, so every descriptor is an immediate
            // (an issue loop that computes offsets at run time is slower than the MMAs it feeds).
            if (elect_one()) {
                constexpr uint32_t idesc = make_idesc(128, 256);
                const uint64_t a_base = make_desc(w_u32, 2048u, 128u);
                const uint64_t b_base = make_desc(act_u32, 0u, 128u);
                issue_layer<Cfg>(l, tmem_base, a_base, b_base, idesc);
                umma_commit(&hdr.tmem_full);   // same thread: the commit tracks the MMAs this thread issued
            }
            TCG_FINE(11);
            __syncwarp();
        } else if (warp == 1) {
            // all MMAs of the layer have completed: the weight buffer and the activation buffer are free
            mbar_wait(&hdr.tmem_full, ph & 1u, 3u);
#ifndef HB_TCG_POLL_ALL
            // this warp is the only one that polls the mbarrier: the eight epilogue warps wait at a hardware barrier it releases
            named_bar_sync(1, 32 * (kGEpiWarps + 1));
#endif
            if (lane == 0 && !last) {
                mbar_expect_tx(&hdr.wbar, (uint32_t)Cfg::w_bytes(l + 1));
                bulk_g2s(wbuf, a.w + Cfg::w_off(l + 1), (uint32_t)Cfg::w_bytes(l + 1), &hdr.wbar);
            } else if (lane == 0 && tile_id + (int)gridDim.x < n_tiles) {
                // the next tile's first layer: its weights arrive under this tile's last epilogue, store and next staging
                mbar_expect_tx(&hdr.wbar, (uint32_t)Cfg::w_bytes(0));
                bulk_g2s(wbuf, a.w, (uint32_t)Cfg::w_bytes(0), &hdr.wbar);
            }
            if (!freq && !last) {
                // the epilogue is writing layout F: zero its SAME padding (column 0 and group 8 of every row, all planes)
                for (int i = lane; i < Cfg::RC * (TT + 1); i += 32) {
                    const int pl = i / (TT + 1), k = i - pl * (TT + 1);
                    const int col = k == 0 ? 0 : 1 + (FG + 1) * (k - 1) + FG;
                    *reinterpret_cast<uint4*>(act + pl * kGPlane + col * 16) = make_uint4(0, 0, 0, 0);
                }
            }
            // stores above -> visible to the tensor core's async-proxy reads, then release the buffer to the MMA issuer
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr.tmem_empty);
        } else {
            // columns per fragment; a warp's two fragments are SUBC-column blocks 2 part, 2 part + 1 -- or part, part + PARTS
            // (interleaved) in the narrow-plane configuration, so that its 128 / 160-column layers spread over all the warps
            constexpr int SUBC = Cfg::SUBC, S_PART = (Cfg::PCOLS < 256 && HB_TCG_INTERLEAVE) ? 1 : 2, S_SUB = (Cfg::PCOLS < 256 && HB_TCG_INTERLEAVE) ? Cfg::PARTS : 1;
            const int e = warp - 2, quad = warp & 3, part = e >> 2;
            const int m = lane >> 3;
            // (sub-position i, chunk cc) -> byte offset in the layer's target layout
            const uint32_t i_unit = freq ? (uint32_t)(FG * 16) : (last ? (uint32_t)(F * 16) : (uint32_t)((FG + 1) * 16));
            const uint32_t cc_unit = last ? (uint32_t)Cfg::PLAIN : (uint32_t)kGPlane;
            // four fragments per warp: (16-lane half h, SUBC-column sub).  Half 0 holds row chunks 3 quad, 3 quad + 1 (stmatrix.x4:
            // matrix m = row chunk m & 1, column group m >> 1), half 1 row chunk 3 quad + 2 (stmatrix.x2: column group m & 1).
            const int rc_a = 3 * quad + (m & 1), rc_b = 3 * quad + 2;
            const uint32_t base_a = act_u32 + (uint32_t)(rc_a / CC) * i_unit + (uint32_t)(rc_a % CC) * cc_unit;
            const uint32_t base_b = act_u32 + (uint32_t)(rc_b / CC) * i_unit + (uint32_t)(rc_b % CC) * cc_unit;
            const uint16_t* tab_a = hdr.tab[l] + (8 * (m >> 1) + (lane & 7)) * kGTabRow + part * (S_PART * SUBC / 16);
            const uint16_t* tab_b = hdr.tab[l] + (8 * (m & 1) + (lane & 7)) * kGTabRow + part * (S_PART * SUBC / 16);
            uint32_t frag[2][SUBC / 2];
            const int n_half = (3 * quad + 2 < Cfg::RC) ? 2 : ((3 * quad < Cfg::RC) ? 1 : 0);   // rows of half 1 / half 0 exist?
            // columns at or past ncols were never written
            const int n_sub = (part * S_PART * SUBC >= Cfg::ncols(l)) ? 0 : (((part * S_PART + S_SUB) * SUBC < Cfg::ncols(l)) ? 2 : 1);
            const int n_frag = n_half * n_sub;
            auto frag_addr = [&](int k) {
                const int h = k / n_sub, sub = k - h * n_sub;
                return tmem_base + ((uint32_t)(quad * 32 + h * 16) << 16) + (uint32_t)((part * S_PART + sub * S_SUB) * SUBC);
            };
            auto frag_tab = [&](int k) {
                const int h = k / n_sub, sub = k - h * n_sub;
                const uint16_t* t = (h == 0 ? tab_a : tab_b) + sub * (S_SUB * SUBC / 16);
                if constexpr (SUBC == 64) return *reinterpret_cast<const uint2*>(t);
                else return make_uint2(*reinterpret_cast<const uint32_t*>(t), 0u);
            };
            // the table entries do not depend on the MMAs: they are in registers before the accumulator is ready
            uint2 tb = n_frag > 0 ? frag_tab(0) : make_uint2(0, 0);
#ifndef HB_TCG_POLL_ALL
            named_bar_sync(1, 32 * (kGEpiWarps + 1));   // released by warp 1 once tmem_full has completed
#else
            mbar_wait(&hdr.tmem_full, ph & 1u);
#endif
            tc_fence_after();
            if (warp == 2) TCG_FINE(12);
            // the TMEM load and the table entries of fragment k + 1 are in flight while fragment k is converted and stored
            if (n_frag > 0) tmem_ld_frag_issue(frag_addr(0), frag[0]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k >= n_frag) break;
                const int h = k / n_sub, sub = k - h * n_sub;
                const uint2 tb_k = tb;
                tmem_ld_frag_wait(frag[k & 1]);
                if (k + 1 < n_frag) {
                    tmem_ld_frag_issue(frag_addr(k + 1), frag[(k + 1) & 1]);
                    tb = frag_tab(k + 1);
                }
                const int col0 = (part * S_PART + sub * S_SUB) * SUBC;
                if (h == 0) g_epilogue<true, SUBC / 2>(frag[k & 1], tb_k, base_a, col0, Cfg::ncols(l));
                else g_epilogue<false, SUBC / 2>(frag[k & 1], tb_k, base_b, col0, Cfg::ncols(l));
            }
            fence_proxy_async();          // this warp's stmatrix stores -> visible to the async proxy before the buffer is released
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr.tmem_empty);
            if (warp == 2) TCG_FINE(13); else if (warp == kGThreads / 32 - 1) TCG_FINE(14);
        }
        // No CTA-wide barrier between layers: the MMA issuer waits on tmem_empty (9 arrivals: every epilogue warp and warp 1 have
        // stored and fenced), the epilogue warps and warp 1 wait on tmem_full.  Only the block's last layer (the store below reads
        // the buffer with all threads) and the parity hook need everyone.
        if (last || a.dbg != nullptr) {
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
        }
        if (stamp_tile) TC_STAMP(3 + l);
        if (a.dbg != nullptr && a.dbg_layer == l && !last) dump_act(!freq, CC);
    }

    // ---- block 4 feeding the tail: conv2d_15's 2x2 max-pool for both time phases, straight from the tile in shared memory ------
    bool stored = false;
    if constexpr (G == 1 && F == 4 && Cfg::POOL_T == 1 && Cfg::POOL_F == 1) {
        if (a.pool2_out != nullptr) {
            // item (chunk, pooled row, phase, pooled bin), bin fastest: rows 2 rp + phase, + 1 and bins 2 fo, 2 fo + 1 of layout P.
            // The four (phase, bin) items of a pooled row start in the 16-byte bank groups 0, 2, 4, 6 of layout P's 128-byte window, so the
            // lanes of odd pooled rows read their two bins in the other order (max is commutative): eight lanes, eight bank groups.
            const int R = a.T_out / 2;
            uint4* o = reinterpret_cast<uint4*>(a.pool2_out) + (int64_t)clip * 2 * R * 2 * CC;
            for (int i = tid; i < CC * 2 * R * 2; i += kGThreads) {
                const int fo = i & 1, ph = (i >> 1) & 1;
                const int r = i >> 2;
                const int rp = r % R, ch = r / R;
                const int rr = 2 * rp + ph;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (rr + 1 < a.T_out) {
                    const unsigned char* base = act + ch * Cfg::PLAIN + (rr * F + 2 * fo) * 16;
                    const int sw = (rp & 1) * 16;
                    const uint4 x0 = *reinterpret_cast<const uint4*>(base + sw), x1 = *reinterpret_cast<const uint4*>(base + 16 - sw);
                    const uint4 x2 = *reinterpret_cast<const uint4*>(base + F * 16 + sw), x3 = *reinterpret_cast<const uint4*>(base + F * 16 + 16 - sw);
                    const __half2 *h0 = reinterpret_cast<const __half2*>(&x0), *h1 = reinterpret_cast<const __half2*>(&x1);
                    const __half2 *h2 = reinterpret_cast<const __half2*>(&x2), *h3 = reinterpret_cast<const __half2*>(&x3);
                    __half2 mx[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) mx[j] = __hmax2_nan(__hmax2_nan(h0[j], h1[j]), __hmax2_nan(h2[j], h3[j]));
                    v = *reinterpret_cast<uint4*>(mx);
                }
                o[((int64_t)(ph * R + rp) * 2 + fo) * CC + ch] = v;
            }
            stored = true;
        }
    }
    // ---- max-pool + store (fp16 chunk-major [clip][OUT_CH][T_out][F / 2][8]) -----------------------------------
    if (!stored) {
        constexpr int PT = Cfg::POOL_T, PF = Cfg::POOL_F, Fo = F / PF;
        constexpr int rows_p = Cfg::ROWS_OUT / PT;
        static_assert(PT == 1 || PF == 2, "a time pool comes with a freq pool");
        const int rowp0 = tile * rows_p;
        uint4* out = reinterpret_cast<uint4*>(a.out);
        for (int i = tid; i < Cfg::OUT_CH * rows_p * Fo; i += kGThreads) {
            const int ch = i / (rows_p * Fo);
            const int rem = i - ch * rows_p * Fo;
            const int rp = rem / Fo;
            int fo = rem - rp * Fo;
            // G = 4: pi(2 fo) = 16 (fo % 2) + fo / 2, so consecutive fo pairs share a 16-byte bank group: an 8-lane group takes the
            // even fo (groups 0..7), the next one the odd fo
            if constexpr (G == 4 && Fo == 16) fo = ((fo & 7) << 1) | (fo >> 3);
            if (rowp0 + rp >= a.T_out) continue;
            uint4 o = make_uint4(0, 0, 0, 0);
            if (ch < CC) {
                constexpr int dpi = (Cfg::pi(1) - Cfg::pi(0)) * 16;   // pi(2 fo + 1) - pi(2 fo), in bytes
                const unsigned char* base = act + ch * Cfg::PLAIN + (PT * rp) * F * 16 + Cfg::pi(PF * fo) * 16;
                // G = 1 with a freq pool: the four fo of a row sit in groups 0, 2, 4, 6 and the next pooled row starts in group 0
                // again: odd rows read their two columns in the other order (max is commutative)
                if constexpr (G == 1 && PF == 2) base += (rp & 1) * dpi;
                const int dpi_l = (G == 1 && PF == 2 && (rp & 1)) ? -dpi : dpi;
                const uint4 x0 = *reinterpret_cast<const uint4*>(base);
                const __half2* h0 = reinterpret_cast<const __half2*>(&x0);
                __half2 mx[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) mx[j] = h0[j];
                if (PF == 2) {
                    const uint4 x1 = *reinterpret_cast<const uint4*>(base + dpi_l);
                    const __half2* h1 = reinterpret_cast<const __half2*>(&x1);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mx[j] = __hmax2_nan(mx[j], h1[j]);
                }
                if (PT == 2) {
                    const uint4 x2 = *reinterpret_cast<const uint4*>(base + F * 16), x3 = *reinterpret_cast<const uint4*>(base + F * 16 + dpi_l);
                    const __half2* h2 = reinterpret_cast<const __half2*>(&x2);
                    const __half2* h3 = reinterpret_cast<const __half2*>(&x3);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mx[j] = __hmax2_nan(mx[j], __hmax2_nan(h2[j], h3[j]));
                }
                o = *reinterpret_cast<uint4*>(mx);
            }
            out[(((int64_t)clip * Cfg::OUT_CH + ch) * a.T_out + rowp0 + rp) * Fo + fo] = o;
        }
    }
    __syncthreads();   // the store has read the buffer: the next tile may stage into it
    if (stamp_tile) TC_STAMP(7);
    }                  // tile loop
    if (warp == 0) tmem_dealloc(tmem_base, kGTmemCols);
    TC_STAMP(8);
}

struct GWeights {
    unsigned char* w[4] = {nullptr, nullptr, nullptr, nullptr};
    float* bias[4] = {nullptr, nullptr, nullptr, nullptr};
    uint16_t* tab[4] = {nullptr, nullptr, nullptr, nullptr};
    float* l0 = nullptr;
};

template <class Cfg>
size_t tcg_smem_bytes() {
    return ((sizeof(GSmemHeader<Cfg>) + 127) & ~(size_t)127) + Cfg::W_MAX + Cfg::ACT_BYTES + 512 + 128 + Cfg::EXTRA_SMEM + 128;
}

// Epilogue scatter table of layer l: accumulator column n -> byte offset >> 4 of its first output record in the layer's target
// layout (the per-row-chunk part is added by the epilogue), stored at (n & 15) * kGTabRow + (n >> 4).  Padding columns of a freq
// layer (group 8 of every row, rows >= TT) go to columns (TT / G) * F + 8 i of their own plane: layout T does not use them, and
// the time conv that follows reads them only as rows >= TT, i.e. into output rows that are dropped anyway (finite garbage instead
// of zeros).  A time layer has no padding columns below ncols(l); the epilogue skips the 16-column groups at or past it.
template <class Cfg>
void tcg_tables(std::vector<uint16_t>& tab) {
    constexpr int G = Cfg::G, CC = Cfg::CC, F = Cfg::F, TT = Cfg::TT, NL = Cfg::NL, FG = Cfg::FG, kGPlane = Cfg::PLANE;
    static_assert((TT / G) * F + (FG > 8 ? FG : 8) * G <= Cfg::PCOLS && (TT / G) * F % 16 == 0, "no free columns for the freq layers' padding outputs");
    tab.assign((size_t)NL * kGTabRow * 16, 0);
    for (int l = 0; l < NL; ++l) {
        int off_of[256];
        bool valid_of[256];
        for (int n = 0; n < 256; ++n) {
            int off = 0;
            bool valid = true;
            if (Cfg::is_freq(l)) {
                // column n = (FG + 1) t + fg -> layout T: plane (t mod G, .), column (t div G) * F + (i * FG +) fg
                const int t = n / (FG + 1), fg = n - t * (FG + 1);
                valid = fg < FG && t < TT;
                off = valid ? (t % G) * CC * kGPlane + ((t / G) * F + fg) * 16 : (TT / G) * F * 16;
            } else {
                const int tq = n / F, pf = n - tq * F;
                valid = tq < TT / G;                                                                  // >= ncols(l): skipped
                if (!valid) off = 0;
                else if (l < NL - 1) off = (pf / FG) * CC * kGPlane + (1 + (FG + 1) * G * tq + (pf % FG)) * 16;   // -> layout F
                else off = (G * tq * F + pf) * 16;                                                    // -> layout P
            }
            off_of[n] = off;
            valid_of[n] = valid;
        }
        // One stmatrix matrix = 8 consecutive columns = 8 row addresses.  A freq layer's matrix straddles a row boundary (FG + 1 = 9
        // columns per row): with a plane pitch that is a multiple of 128 B the records of the two rows use different 16-byte bank
        // groups, and the padding column(s) in between are dumped to the group(s) the matrix leaves free -- one wavefront per matrix.
        if (Cfg::is_freq(l) && (G == 1 || kGPlane % 128 == 0))
            for (int n0 = 0; n0 < 256; n0 += 8) {
                bool used[8] = {false, false, false, false, false, false, false, false};
                for (int n = n0; n < n0 + 8; ++n)
                    if (valid_of[n]) used[(off_of[n] >> 4) & 7] = true;
                int u = 0;
                for (int n = n0; n < n0 + 8; ++n) {
                    if (valid_of[n]) continue;
                    while (u < 8 && used[u]) ++u;
                    if (u < 8) {
                        off_of[n] = (TT / G) * F * 16 + u * 16;
                        used[u] = true;
                    }
                }
            }
        for (int n = 0; n < 256; ++n)
            if (valid_of[n] || Cfg::is_freq(l))
                tab[(size_t)l * kGTabRow * 16 + (size_t)(n & 15) * kGTabRow + (n >> 4)] = (uint16_t)(off_of[n] >> 4);
    }
}

// Banded Toeplitz A operands of the block's layers: [layer][K chunk][row 128][8 cin]; row = 32 q + 8 o + channel holds row
// chunk rc = 3 q + o (o < 3) = (sub-position rc / CC, channel chunk rc % CC).
template <class Cfg>
int tcg_pack(const float* weights_host, const std::vector<int64_t>& w_off, const std::vector<int64_t>& b_off, unsigned char** w_dev,
             float** bias_dev, uint16_t** tab_dev) {
    constexpr int G = Cfg::G, CC = Cfg::CC;
    std::vector<__half> packed((size_t)Cfg::w_off(Cfg::NL) / 2, __float2half_rn(0.f));
    std::vector<float> bias((size_t)Cfg::NL * Cfg::C);
    for (int l = 0; l < Cfg::NL; ++l) {
        const int li = Cfg::CONV0 + l;
        const ConvLayer& L = kLayers[li];
        const bool freq = Cfg::is_freq(l);
        const bool mel_layer = Cfg::MEL_IN && l == 0;         // conv2d: Cin = 1, fed as the (hi, lo) fp16 pair of the mel value
        HB_REQUIRE((mel_layer ? L.cin == 1 && Cfg::cin(l) == 1 : L.cin == Cfg::cin(l) * 8) && L.cout == Cfg::C && L.kh * L.kw == 3 &&
                       L.leaky && (L.kw == 3) == freq,
                   "tcg: unexpected layer table entry for conv2d_%d", li);
        const float* w = weights_host + w_off[li];            // [tap][cin][cout]
        for (int kk = 0; kk < Cfg::kchunks(l); ++kk) {
            const int c = kk / (G + 2), d = Cfg::kd(l, kk);    // channel chunk, input offset within the group (df or dt)
            if (kk & 1) HB_REQUIRE(Cfg::koff(l, kk) > Cfg::koff(l, kk - 1), "tcg: K chunk pair %d of layer %d is not address-ordered", kk / 2, l);
            if (kk == Cfg::kreal(l)) {
                // ones chunk: bias as an fp16 hi + lo pair (the B record is [1, 1, 0..]), exact to ~2^-22
                for (int row = 0; row < 128; ++row) {
                    const int q = row >> 5, o = (row >> 3) & 3, r = row & 7;
                    const int rc = 3 * q + o, cc = rc % CC;
                    if (o >= 3 || rc >= Cfg::RC) continue;
                    const float b = weights_host[b_off[li] + cc * 8 + r];
                    const __half hi = __float2half_rn(b);
                    const size_t at = (size_t)Cfg::w_off(l) / 2 + ((size_t)kk * 128 + row) * 8;
                    packed[at] = hi;
                    packed[at + 1] = __float2half_rn(b - __half2float(hi));
                }
                continue;
            }
            if (kk > Cfg::kreal(l)) continue;                   // zero pad chunk
            for (int row = 0; row < 128; ++row) {
                const int q = row >> 5, o = (row >> 3) & 3, r = row & 7;
                const int rc = 3 * q + o, i = rc / CC, cc = rc % CC;
                if (o >= 3 || rc >= Cfg::RC) continue;
                const int tap = freq ? d - i + 1 : d - i;
                if (tap < 0 || tap > 2) continue;
                for (int e = 0; e < (mel_layer ? 2 : 8); ++e)
                    packed[(size_t)Cfg::w_off(l) / 2 + ((size_t)kk * 128 + row) * 8 + e] =
                        __float2half_rn(w[((int64_t)tap * L.cin + (mel_layer ? 0 : c * 8 + e)) * L.cout + cc * 8 + r]);
            }
        }
        for (int n = 0; n < Cfg::C; ++n) bias[(size_t)l * Cfg::C + n] = weights_host[b_off[li] + n];
    }
    HB_CUDA_OK(cudaMalloc(w_dev, packed.size() * 2));
    HB_CUDA_OK(cudaMemcpy(*w_dev, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMalloc(bias_dev, bias.size() * sizeof(float)));
    HB_CUDA_OK(cudaMemcpy(*bias_dev, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::vector<uint16_t> tab;
    tcg_tables<Cfg>(tab);
    HB_CUDA_OK(cudaMalloc(tab_dev, tab.size() * sizeof(uint16_t)));
    HB_CUDA_OK(cudaMemcpy(*tab_dev, tab.data(), tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    return HB_OK;
}

template <class Cfg>
int tcg_launch(const GWeights* gw, int which, const void* in, int in_chunks, __half* out, int B, int in_T, float* dbg, int dbg_layer,
               cudaStream_t st, __half* pool2_out = nullptr) {
    GArgs a;
    a.in = in;
    a.out = out;
    a.pool2_out = pool2_out;
    a.w = gw->w[which];
    a.bias = gw->bias[which];
    a.l0_w = gw->l0;
    a.tab = gw->tab[which];
    a.dbg = dbg;
    a.dbg_layer = dbg ? dbg_layer : -1;
    int t_convs = 0;   // time convs up to and including the dumped layer shrink its valid rows
    if (dbg && dbg_layer != 100)
        for (int l = 0; l <= dbg_layer && l < Cfg::NL; ++l) t_convs += !Cfg::is_freq(l);
    a.dbg_T = in_T - 2 * t_convs;
    a.n_clips = B;
    a.in_T = in_T;
    a.in_chunks = in_chunks;
    a.T_out = (in_T - 2 * Cfg::N_TIME) / Cfg::POOL_T;
    const int rows_needed = dbg ? a.dbg_T : Cfg::POOL_T * a.T_out;
    a.tiles_per_clip = std::max(1, ceil_div(rows_needed, Cfg::ROWS_OUT));
    HB_REQUIRE((int64_t)B * a.tiles_per_clip < (1ll << 31), "tcg: grid too large");
    static bool configured = false;
    if (!configured) {
        HB_CUDA_OK(cudaFuncSetAttribute(tcg_block_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcg_smem_bytes<Cfg>()));
        HB_CUDA_OK(cudaFuncSetAttribute(tcg_block_kernel<Cfg>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        HB_CUDA_OK(cudaGetDevice(&dev));
        HB_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int64_t n_tiles = (int64_t)B * a.tiles_per_clip;
    const int grid = (int)std::min<int64_t>(n_tiles, (Cfg::THREADS > 320 ? HB_TCG_WIDE_CTAS : 2) * (int64_t)n_sm);   // persistent: two CTAs per SM stride over the tiles
    tcg_block_kernel<Cfg><<<grid, Cfg::THREADS, tcg_smem_bytes<Cfg>(), st>>>(a);
    HB_LAUNCHED();
    return HB_OK;
}

}  // namespace

int tcg_prepare(hb_embed_model* m, const float* weights_host) {
    GWeights* gw = new GWeights();
    std::vector<int64_t> w_off(kNumConv), b_off(kNumConv);
    int64_t off = 0;
    for (int i = 0; i < kNumConv; ++i) {
        w_off[i] = off;
        off += layer_weight_floats(kLayers[i]);
        b_off[i] = off;
        off += kLayers[i].cout;
    }
    m->tcg = gw;
    int rc;
    if ((rc = tcg_pack<Cfg1>(weights_host, w_off, b_off, &gw->w[0], &gw->bias[0], &gw->tab[0]))) return rc;
    if ((rc = tcg_pack<Cfg2>(weights_host, w_off, b_off, &gw->w[1], &gw->bias[1], &gw->tab[1]))) return rc;
    if ((rc = tcg_pack<Cfg3>(weights_host, w_off, b_off, &gw->w[2], &gw->bias[2], &gw->tab[2]))) return rc;
    if ((rc = tcg_pack<Cfg4>(weights_host, w_off, b_off, &gw->w[3], &gw->bias[3], &gw->tab[3]))) return rc;
    HB_CUDA_OK(cudaMalloc(&gw->l0, (3 * 24 + 24) * sizeof(float)));
    HB_CUDA_OK(cudaMemcpy(gw->l0, weights_host + w_off[0], (3 * 24 + 24) * sizeof(float), cudaMemcpyHostToDevice));
    return HB_OK;
}

void tcg_release(hb_embed_model* m) {
    GWeights* gw = reinterpret_cast<GWeights*>(m->tcg);
    if (!gw) return;
    for (int i = 0; i < 4; ++i) {
        cudaFree(gw->w[i]);
        cudaFree(gw->bias[i]);
        cudaFree(gw->tab[i]);
    }
    cudaFree(gw->l0);
    delete gw;
    m->tcg = nullptr;
}

// mel f32 [B][in_T][32] -> pooled conv2d_3 output, fp16 chunk-major [B][4][T_out][16][8], T_out = (in_T - 4) / 2.
// dbg != nullptr: also dump the activation after conv2d (dbg_layer 100) / conv2d_1 (0) / conv2d_2 (1) as f32 NHWC.
int tcg_block1(const hb_embed_model* m, const float* mel, __half* out, int B, int in_T, float* dbg, int dbg_layer,
               cudaStream_t st) {
    const GWeights* gw = reinterpret_cast<const GWeights*>(m->tcg);
    HB_REQUIRE(gw != nullptr, "tcg weights missing");
    // conv2d is tensor-core layer 0 of the block: the caller's 100 (conv2d) / 0 / 1 are layers 0 / 1 / 2
    return tcg_launch<Cfg1>(gw, 0, mel, 0, out, B, in_T, dbg, dbg_layer == 100 ? 0 : dbg_layer + 1, st);
}

// block 1 output fp16 [B][4][in_T][16][8] -> conv2d_7 output after its 1x2 pool, fp16 chunk-major [B][6][in_T - 4][8][8].
// dbg != nullptr: dump the activation after conv2d_4 / 5 / 6 (dbg_layer 0 / 1 / 2) as f32 NHWC [B][rows][16][48].
int tcg_block2(const hb_embed_model* m, const __half* in, __half* out, int B, int in_T, float* dbg, int dbg_layer,
               cudaStream_t st) {
    const GWeights* gw = reinterpret_cast<const GWeights*>(m->tcg);
    HB_REQUIRE(gw != nullptr, "tcg weights missing");
    return tcg_launch<Cfg2>(gw, 1, in, 4, out, B, in_T, dbg, dbg_layer, st);
}

// block 2 output fp16 [B][6][in_T][8][8] -> conv2d_11 output after its 2x2 pool, fp16 chunk-major [B][10][(in_T - 4) / 2][4][8]
// (chunk 9 = zero padding to K = 80).  dbg: activation after conv2d_8 / 9 / 10 (dbg_layer 0 / 1 / 2), f32 NHWC [B][rows][8][72].
int tcg_block3(const hb_embed_model* m, const __half* in, __half* out, int B, int in_T, float* dbg, int dbg_layer,
               cudaStream_t st) {
    const GWeights* gw = reinterpret_cast<const GWeights*>(m->tcg);
    HB_REQUIRE(gw != nullptr, "tcg weights missing");
    return tcg_launch<Cfg3>(gw, 2, in, 6, out, B, in_T, dbg, dbg_layer, st);
}

// block 3 output fp16 [B][10][in_T][4][8] (chunk 9 = padding, not read) -> conv2d_15 output before its pool, fp16 chunk-major
// [B][12][in_T - 4][4][8] (a 1.44 s clip has in_T = 30: one tile).  dbg: activation after conv2d_12 / 13 / 14, f32 NHWC [B][rows][4][96].
// pool2_out != nullptr (and in_T - 4 <= 28 rows: one tile per clip): instead of `out`, write conv2d_15's 2x2 pool for both time phases
// in the tail's operand format, fp16 [B][2 phases][(in_T - 4) / 2 rows][2 bins][12 chunks][8].
int tcg_block4(const hb_embed_model* m, const __half* in, __half* out, int B, int in_T, float* dbg, int dbg_layer,
               cudaStream_t st, __half* pool2_out) {
    const GWeights* gw = reinterpret_cast<const GWeights*>(m->tcg);
    HB_REQUIRE(gw != nullptr, "tcg weights missing");
    HB_REQUIRE(pool2_out == nullptr || (dbg == nullptr && in_T - 4 <= Cfg4::ROWS_OUT), "tcg_block4: the pooled store needs one tile per clip");
    return tcg_launch<Cfg4>(gw, 3, in, 10, out, B, in_T, dbg, dbg_layer, st, pool2_out);
}
int tcg_block4_rows_per_tile() { return Cfg4::ROWS_OUT; }
int tcg_block4_max_rows() { return Cfg4::TT; }

// profiling aid: phase timestamps (start, setup, staged, layers.., stored, dealloc) of the first 8 CTAs of the last launch
int tcg_debug_times(long long* out_host) {
    return cudaMemcpyFromSymbol(out_host, g_tc_times, sizeof(long long) * 8 * 16) == cudaSuccess ? 0 : -2;
}

int tcg_check_timeout() {
    unsigned int flag = 0;
    HB_CUDA_OK(cudaMemcpyFromSymbol(&flag, g_tc_timeout, sizeof(flag)));
    HB_REQUIRE(flag == 0, "tcgen05 embed kernel (blocks 1-4): an mbarrier wait timed out (pipeline bug; barrier code %u: 2 = weights, 3 = accumulator free, 4 = accumulator ready)", flag);
    return HB_OK;
}

}  // namespace hb
