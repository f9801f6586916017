// K7, block 1 (conv2d .. conv2d_3 + 2x2 max-pool): tcgen05 implicit GEMMs with FOUR output positions per column.
//
// Why a second kernel: an SS tcgen05.mma costs max(141, N/2 + 43) cycles whatever M is (scripts/micro/umma_bench.cu),
// and block 1 has only 24 channels, so the one-position-per-column form of embed_tc.cu fills 24 of the 128 M rows
// and pays 6 MMAs (3 taps x K = 32) per 256 positions.  Here M carries (sub-position i = 0..3, cout = 24) = 96
// rows: a column is a group of four adjacent positions ALONG THE CONV AXIS, K runs over the six input positions the
// group touches (6 x 24 = 144 = nine K = 16 steps) and the A operand is the banded Toeplitz matrix of the 3-tap
// kernel (zeros where a tap does not connect).  9 MMAs per 1024 positions instead of 24: block 1 needs 162 MMAs
// per 1.44 s clip instead of 504, and a third of the epilogue calls.
//
// Layouts (all [plane][column][8 x fp16], 16-byte records, one plane per (phase, 8-channel chunk), 4096 B planes):
//   T (input of a time conv):  plane (t mod 4, chunk), column (t div 4) * 32 + pi(f),  pi(f) = (f mod 4) * 8 + f div 4
//   F (input of the freq conv): plane (f mod 4, chunk), column 1 + 9 t + f div 4  (group 8 of every row and column 0
//                               are the zero SAME padding)
//   P (pre-pool output):        [chunk][t * 32 + pi(f)]
// A K step is two K chunks = two plane addresses: the UMMA descriptor's (start, LBO) pair expresses any of them, so
// there is still no im2col.  The permutation pi makes every stmatrix of the epilogue (8 consecutive columns x 8
// channels) land on 128 contiguous bytes in either target layout.  The tile's single activation buffer is rewritten
// in place: a layer has one accumulator tile (N = 256 columns), so all of its MMAs have completed (tcgen05.commit)
// before the first epilogue store.
//
// Tile = 28 input rows of one clip -> 24 conv2d_3 rows -> 12 pooled rows; twin launch shape (320 threads, 2 CTAs/SM,
// 256 TMEM columns each) like embed_tc.cu.  Warp 0 issues MMAs, warp 1 streams the next layer's weights as soon as
// the current MMAs have completed and zeroes the SAME padding, warps 2-9 run the epilogue (TMEM lane quadrant q holds
// sub-position i = q: rows 32 q + 8 chunk + channel).
//
// NaN note: the zero Toeplitz entries multiply neighbouring positions of the same tile, so a non-finite activation
// reaches up to 3 more positions of its own clip than in the reference arithmetic (0 * inf); finite data is unaffected.
#include "tc_ptx.cuh"

#include <vector>

namespace hb {

namespace {

constexpr int kGThreads = 320;
constexpr int kGEpiWarps = 8;
constexpr int kGTt = 28;            // input rows per tile
constexpr int kGRowsOut = 24;       // conv2d_3 rows per tile (pre-pool)
constexpr int kGC = 24;             // channels
constexpr int kGPlane = 4096;       // bytes per plane
constexpr int kGPlanes = 12;
constexpr int kGActBytes = kGPlanes * kGPlane;
constexpr int kGPlainPlane = kGTt * 32 * 16;     // layout P bytes per chunk
constexpr int kGWBytes = 18 * 128 * 16;          // one layer's A operand: [K chunk 18][row 128][8]
constexpr int kGTmemCols = 256;

struct GArgs {
    const float* mel;             // f32 [clips][in_T][32]
    __half* out;                  // fp16 chunk-major [clips][4][T_out][16][8] (chunk 3 = zero padding to K = 32)
    const unsigned char* w;       // 3 layers x kGWBytes
    const float* bias;            // 3 x 24
    const float* l0_w;            // conv2d kernel f32 [3][24] + bias [24]
    float* dbg;                   // optional f32 NHWC [clips][dbg_T][32][24] activation dump
    int dbg_layer;                // -1 none, 100 = conv2d output, 0 / 1 = conv2d_1 / conv2d_2 output
    int dbg_T;
    int n_clips, in_T, T_out, tiles_per_clip;
};

struct GSmemHeader {
    uint64_t tmem_full, tmem_empty, wbar;
    uint32_t tmem_base;
    uint32_t pad[1];
    float bias[3 * kGC];
    float l0[3 * 24 + 24 + 8];
    uint16_t tab[3][256];         // column -> ((byte offset >> 4) << 1) | invalid, per layer
};

__host__ __device__ constexpr int pi_f(int f) { return (f & 3) * 8 + (f >> 2); }

// byte offset (within the activation buffer) of K chunk kk of a time / freq layer, relative to column n = 0
__device__ __forceinline__ uint32_t koff_time(int kk) {
    const int dt = kk / 3, c = kk - dt * 3;
    return (uint32_t)(((dt & 3) * 3 + c) * kGPlane + (dt >> 2) * 32 * 16);
}
__device__ __forceinline__ uint32_t koff_freq(int kk) {
    const int ord = kk / 3, c = kk - ord * 3;                 // df = 0, 1, 2, 3, 4, -1
    const int fm = ord < 4 ? ord : (ord == 4 ? 0 : 3);
    const int sh = ord < 4 ? 1 : (ord == 4 ? 2 : 0);          // 1 + column shift (column 0 is the guard)
    return (uint32_t)((fm * 3 + c) * kGPlane + sh * 16);
}

template <bool kTwo>
__device__ __forceinline__ void g_epilogue(uint32_t taddr, const uint16_t* __restrict__ tab, int n_lane, uint32_t base_lane,
                                           uint32_t dump_lane, float b0, float b1) {
    float v[32];
    tmem_ld_16x256b_64cols(taddr, v);
#pragma unroll
    for (int g = 0; g < 8; g += 2) {
        const uint32_t e = tab[n_lane + 8 * g];
        const uint32_t addr = (e & 1u) ? dump_lane : base_lane + ((e >> 1) << 4);
        const float x0 = leaky(v[4 * g + 0] + b0), x1 = leaky(v[4 * g + 1] + b0);
        const float x4 = leaky(v[4 * g + 4] + b0), x5 = leaky(v[4 * g + 5] + b0);
        const uint32_t ra = pack_half2(x0, x1), rc = pack_half2(x4, x5);
        if (kTwo) {
            const float x2 = leaky(v[4 * g + 2] + b1), x3 = leaky(v[4 * g + 3] + b1);
            const float x6 = leaky(v[4 * g + 6] + b1), x7 = leaky(v[4 * g + 7] + b1);
            stmatrix_x4_trans(addr, ra, pack_half2(x2, x3), rc, pack_half2(x6, x7));
        } else {
            stmatrix_x2_trans(addr, ra, rc);
        }
    }
}

__global__ void __launch_bounds__(kGThreads, 2) tcg_block1_kernel(const GArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    GSmemHeader& hdr = *reinterpret_cast<GSmemHeader*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* wbuf = smem + ((sizeof(GSmemHeader) + 127) & ~127);
    unsigned char* act = wbuf + kGWBytes;
    unsigned char* dump = act + kGActBytes + 512;             // 512 B finite guard for reads past the last plane
    float* mel_tile = reinterpret_cast<float*>(dump + 128);

    const int clip = blockIdx.x / a.tiles_per_clip, tile = blockIdx.x - clip * a.tiles_per_clip;
    const int row0 = tile * kGRowsOut;

    // ---- one-time setup ----------------------------------------------------------------------------------
    if (tid == 0) {
        mbar_init(&hdr.tmem_full, 1);
        mbar_init(&hdr.tmem_empty, kGEpiWarps);
        mbar_init(&hdr.wbar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&hdr.tmem_base, kGTmemCols);
    for (int i = tid; i < 3 * kGC; i += kGThreads) hdr.bias[i] = a.bias[i];
    for (int i = tid; i < 3 * 24 + 24; i += kGThreads) hdr.l0[i] = a.l0_w[i];
    for (int i = tid; i < 3 * 256; i += kGThreads) {
        const int l = i >> 8, n = i & 255;
        int off, valid;
        if (l == 1) {
            // freq layer, column n = 9 t + fg -> layout T: plane (t mod 4, .), column (t div 4) * 32 + (i * 8 +) fg
            const int t = n / 9, fg = n - t * 9;
            valid = fg < 8 && t < kGTt;
            off = (t & 3) * 3 * kGPlane + ((t >> 2) * 32 + fg) * 16;
        } else {
            const int tq = n >> 5, pf = n & 31;
            valid = tq < kGTt / 4;
            if (l == 0) off = (pf >> 3) * 3 * kGPlane + (1 + 36 * tq + (pf & 7)) * 16;   // -> layout F
            else off = (128 * tq + pf) * 16;                                              // -> layout P
        }
        hdr.tab[l][n] = (uint16_t)(valid ? ((off >> 4) << 1) : 1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = hdr.tmem_base;

    if (warp == 1 && lane == 0) {
        mbar_expect_tx(&hdr.wbar, (uint32_t)kGWBytes);
        bulk_g2s(wbuf, a.w, (uint32_t)kGWBytes, &hdr.wbar);
    }

    // ---- stage: mel rows -> conv2d (Cin = 1, CUDA cores) -> layout T ---------------------------------------
    {
        const float* mel = a.mel + (int64_t)clip * a.in_T * kMels;
        for (int i = tid; i < kGTt * kMels; i += kGThreads) {
            const int r = i / kMels;
            mel_tile[i] = (row0 + r < a.in_T) ? __ldg(mel + (int64_t)(row0 + r) * kMels + (i - r * kMels)) : 0.f;
        }
        // the rows past the tile (t = 28..31) only feed outputs that are dropped, but must be finite
        for (int i = tid; i < kGPlanes * 32 + 32; i += kGThreads) {
            unsigned char* p = i < kGPlanes * 32 ? act + (i >> 5) * kGPlane + ((kGTt / 4) * 32 + (i & 31)) * 16
                                                 : act + kGActBytes + (i - kGPlanes * 32) * 16;
            *reinterpret_cast<uint4*>(p) = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
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
        if (ch < 3) {
            for (int r = warp; r < kGTt; r += kGThreads / 32) {
                const float* mrow = mel_tile + r * kMels;
                unsigned char* dst = act + ((r & 3) * 3 + ch) * kGPlane + (r >> 2) * 32 * 16;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const int f = (lane >> 2) + 8 * j4;
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
                    *reinterpret_cast<uint4*>(dst + pi_f(f) * 16) =
                        make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                }
            }
        }
    }
    fence_proxy_async();   // generic-proxy stores above -> visible to the tensor core's async-proxy reads
    __syncthreads();

    auto dump_layout_t = [&](int t_layer) {
        // f32 NHWC dump of the tile's rows from layout T (profiling / parity hook only)
        for (int i = tid; i < kGRowsOut * 32 * kGC; i += kGThreads) {
            const int c = i % kGC, f = (i / kGC) & 31, t = i / (kGC * 32);
            if (row0 + t >= t_layer) continue;
            const __half* p = reinterpret_cast<const __half*>(act + ((t & 3) * 3 + (c >> 3)) * kGPlane + ((t >> 2) * 32 + pi_f(f)) * 16);
            a.dbg[(((int64_t)clip * t_layer + row0 + t) * 32 + f) * kGC + c] = __half2float(p[c & 7]);
        }
    };
    if (a.dbg != nullptr && a.dbg_layer == 100) dump_layout_t(a.dbg_T);

    // ---- conv2d_1 (time), conv2d_2 (freq), conv2d_3 (time) ---------------------------------------------------
    const uint32_t act_u32 = smem_u32(act), w_u32 = smem_u32(wbuf);
    for (int l = 0; l < 3; ++l) {
        const bool freq = (l == 1);
        if (warp == 0) {
            mbar_wait(&hdr.wbar, (uint32_t)(l & 1));
            if (l > 0) mbar_wait(&hdr.tmem_empty, (uint32_t)((l - 1) & 1));
            tc_fence_after();
            const uint32_t idesc = make_idesc(128, 256);
#pragma unroll 1
            for (int j = 0; j < 9; ++j) {
                const uint32_t o0 = freq ? koff_freq(2 * j) : koff_time(2 * j);
                const uint32_t o1 = freq ? koff_freq(2 * j + 1) : koff_time(2 * j + 1);
                const uint64_t ad = make_desc(w_u32 + (uint32_t)(2 * j) * 2048u, 2048u, 128u);
                const uint64_t bd = make_desc(act_u32 + o0, o1 - o0, 128u);
                if (elect_one()) umma_f16(tmem_base, ad, bd, idesc, j > 0 ? 1u : 0u);
            }
            if (elect_one()) umma_commit(&hdr.tmem_full);
            __syncwarp();
        } else if (warp == 1) {
            // all MMAs of the layer have completed: the weight buffer and the activation buffer are free
            mbar_wait(&hdr.tmem_full, (uint32_t)(l & 1));
            if (lane == 0 && l < 2) {
                mbar_expect_tx(&hdr.wbar, (uint32_t)kGWBytes);
                bulk_g2s(wbuf, a.w + (size_t)(l + 1) * kGWBytes, (uint32_t)kGWBytes, &hdr.wbar);
            }
            if (l == 0) {
                // layout F zero padding: column 0 and group 8 of every row, in all 12 planes
                for (int i = lane; i < kGPlanes * (kGTt + 1); i += 32) {
                    const int pl = i / (kGTt + 1), k = i - pl * (kGTt + 1);
                    const int col = k == 0 ? 0 : 1 + 9 * (k - 1) + 8;
                    *reinterpret_cast<uint4*>(act + pl * kGPlane + col * 16) = make_uint4(0, 0, 0, 0);
                }
            }
            fence_proxy_async();
            __syncwarp();
        } else {
            const int e = warp - 2, quad = warp & 3, part = e >> 2;   // quad = sub-position i, part = column half
            const float* bias = hdr.bias + l * kGC;
            const uint16_t* tab = hdr.tab[l];
            const int m = lane >> 3;
            // byte offset of (sub-position quad, chunk cc) in the layer's target layout
            const uint32_t g2_unit = l == 0 ? 9u * 16u : (l == 1 ? 128u : 512u);
            const uint32_t cc_stride = l == 2 ? (uint32_t)kGPlainPlane : (uint32_t)kGPlane;
            const uint32_t dump_lane = smem_u32(dump) + (uint32_t)((lane & 7) * 16);
            mbar_wait(&hdr.tmem_full, (uint32_t)(l & 1));
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cc_lane = h == 0 ? (m & 1) : 2;                       // chunk this lane's stmatrix rows belong to
                const int pg = h == 0 ? (m >> 1) : (m & 1);                     // +8 column group of this lane's matrix
                const uint32_t base_lane = act_u32 + (uint32_t)quad * g2_unit + (uint32_t)cc_lane * cc_stride;
                const float b0 = bias[(2 * h) * 8 + (lane >> 2)];
                const float b1 = h == 0 ? bias[8 + (lane >> 2)] : 0.f;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const int col = part * 128 + sub * 64;
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32 + h * 16) << 16) + (uint32_t)col;
                    const int n_lane = col + 8 * pg + (lane & 7);
                    if (h == 0) g_epilogue<true>(taddr, tab, n_lane, base_lane, dump_lane, b0, b1);
                    else g_epilogue<false>(taddr, tab, n_lane, base_lane, dump_lane, b0, b1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hdr.tmem_empty);
            fence_proxy_async();
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (a.dbg != nullptr && a.dbg_layer == l) {
            if (l == 0) {
                for (int i = tid; i < kGRowsOut * 32 * kGC; i += kGThreads) {
                    const int c = i % kGC, f = (i / kGC) & 31, t = i / (kGC * 32);
                    if (row0 + t >= a.dbg_T) continue;
                    const __half* p = reinterpret_cast<const __half*>(act + ((f & 3) * 3 + (c >> 3)) * kGPlane + (1 + 9 * t + (f >> 2)) * 16);
                    a.dbg[(((int64_t)clip * a.dbg_T + row0 + t) * 32 + f) * kGC + c] = __half2float(p[c & 7]);
                }
            } else if (l == 1) {
                dump_layout_t(a.dbg_T);
            }
        }
    }

    // ---- 2x2 max-pool + store (fp16 chunk-major [clip][4][T_out][16][8]) --------------------------------------
    {
        constexpr int rows_p = kGRowsOut / 2;
        const int rowp0 = tile * rows_p;
        uint4* out = reinterpret_cast<uint4*>(a.out);
        for (int i = tid; i < 4 * rows_p * 16; i += kGThreads) {
            const int ch = i / (rows_p * 16);
            const int rem = i - ch * rows_p * 16;
            const int rp = rem >> 4, fo = rem & 15;
            if (rowp0 + rp >= a.T_out) continue;
            uint4 o = make_uint4(0, 0, 0, 0);
            if (ch < 3) {
                const unsigned char* base = act + ch * kGPlainPlane + (2 * rp) * 32 * 16 + pi_f(2 * fo) * 16;
                const uint4 x0 = *reinterpret_cast<const uint4*>(base), x1 = *reinterpret_cast<const uint4*>(base + 8 * 16);
                const uint4 x2 = *reinterpret_cast<const uint4*>(base + 32 * 16), x3 = *reinterpret_cast<const uint4*>(base + 40 * 16);
                const __half2* h0 = reinterpret_cast<const __half2*>(&x0);
                const __half2* h1 = reinterpret_cast<const __half2*>(&x1);
                const __half2* h2 = reinterpret_cast<const __half2*>(&x2);
                const __half2* h3 = reinterpret_cast<const __half2*>(&x3);
                __half2 mx[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) mx[j] = __hmax2_nan(__hmax2_nan(h0[j], h1[j]), __hmax2_nan(h2[j], h3[j]));
                o = *reinterpret_cast<uint4*>(mx);
            }
            out[(((int64_t)clip * 4 + ch) * a.T_out + rowp0 + rp) * 16 + fo] = o;
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kGTmemCols);
}

struct GWeights {
    unsigned char* w = nullptr;
    float* bias = nullptr;
    float* l0 = nullptr;
};

size_t tcg_smem_bytes() {
    return ((sizeof(GSmemHeader) + 127) & ~(size_t)127) + kGWBytes + kGActBytes + 512 + 128 + kGTt * kMels * sizeof(float) + 128;
}

}  // namespace

// Pack conv2d_1..3 as banded Toeplitz A operands: [layer][K chunk 18][row 128][8 cin], row = 32 i + 8 chunk + channel.
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
    std::vector<__half> packed((size_t)3 * kGWBytes / 2, __float2half_rn(0.f));
    std::vector<float> bias(3 * kGC);
    for (int l = 0; l < 3; ++l) {
        const ConvLayer& L = kLayers[1 + l];
        HB_REQUIRE(L.cin == kGC && L.cout == kGC && L.kh * L.kw == 3 && L.leaky, "tcg: unexpected layer table for conv2d_%d", 1 + l);
        const bool freq = (L.kw == 3);
        HB_REQUIRE(freq == (l == 1), "tcg: conv2d_%d orientation", 1 + l);
        const float* w = weights_host + w_off[1 + l];            // [tap][cin][cout]
        for (int kk = 0; kk < 18; ++kk) {
            const int ord = kk / 3, c = kk % 3;
            const int d = freq ? (ord < 5 ? ord : -1) : ord;       // input offset within the group: df (freq) or dt (time)
            for (int row = 0; row < 128; ++row) {
                const int i = row >> 5, cc = (row >> 3) & 3, r = row & 7;
                if (cc >= 3) continue;
                const int tap = freq ? d - i + 1 : d - i;
                if (tap < 0 || tap > 2) continue;
                for (int e = 0; e < 8; ++e)
                    packed[(((size_t)l * 18 + kk) * 128 + row) * 8 + e] =
                        __float2half_rn(w[((int64_t)tap * kGC + c * 8 + e) * kGC + cc * 8 + r]);
            }
        }
        for (int n = 0; n < kGC; ++n) bias[l * kGC + n] = weights_host[b_off[1 + l] + n];
    }
    HB_CUDA_OK(cudaMalloc(&gw->w, packed.size() * 2));
    HB_CUDA_OK(cudaMemcpy(gw->w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMalloc(&gw->bias, bias.size() * sizeof(float)));
    HB_CUDA_OK(cudaMemcpy(gw->bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    HB_CUDA_OK(cudaMalloc(&gw->l0, (3 * 24 + 24) * sizeof(float)));
    HB_CUDA_OK(cudaMemcpy(gw->l0, weights_host + w_off[0], (3 * 24 + 24) * sizeof(float), cudaMemcpyHostToDevice));
    m->tcg = gw;
    return HB_OK;
}

void tcg_release(hb_embed_model* m) {
    GWeights* gw = reinterpret_cast<GWeights*>(m->tcg);
    if (!gw) return;
    cudaFree(gw->w);
    cudaFree(gw->bias);
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
    GArgs a;
    a.mel = mel;
    a.out = out;
    a.w = gw->w;
    a.bias = gw->bias;
    a.l0_w = gw->l0;
    a.dbg = dbg;
    a.dbg_layer = dbg ? dbg_layer : -1;
    a.dbg_T = dbg_layer == 100 ? in_T : (dbg_layer == 0 || dbg_layer == 1 ? in_T - 2 : in_T - 4);
    a.n_clips = B;
    a.in_T = in_T;
    a.T_out = (in_T - 4) / 2;
    // cover every row a dump may ask for (in_T for the conv2d output), i.e. ceil(in_T / 24) tiles when dumping
    const int rows_needed = dbg ? a.dbg_T : 2 * a.T_out;
    a.tiles_per_clip = std::max(1, ceil_div(rows_needed, kGRowsOut));
    HB_REQUIRE((int64_t)B * a.tiles_per_clip < (1ll << 31), "tcg: grid too large");
    static bool configured = false;
    if (!configured) {
        HB_CUDA_OK(cudaFuncSetAttribute(tcg_block1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcg_smem_bytes()));
        HB_CUDA_OK(cudaFuncSetAttribute(tcg_block1_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    tcg_block1_kernel<<<B * a.tiles_per_clip, kGThreads, tcg_smem_bytes(), st>>>(a);
    HB_LAUNCHED();
    return HB_OK;
}

int tcg_check_timeout() {
    unsigned int flag = 0;
    HB_CUDA_OK(cudaMemcpyFromSymbol(&flag, g_tc_timeout, sizeof(flag)));
    HB_REQUIRE(flag == 0, "tcgen05 embed kernel (block 1): an mbarrier wait timed out (pipeline bug)");
    return HB_OK;
}

}  // namespace hb
