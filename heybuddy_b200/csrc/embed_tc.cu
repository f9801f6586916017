// K7 (product mode): the speech-embedding conv stack on tcgen05 -- the driver of the five stages.
//
// Replaces the ORT run of speech-embedding.onnx behind SpeechEmbeddingModel.__call__ (reference
// src/python/heybuddy/embeddings.py:32-42).  Layer table: embed_common.cuh / spec.py.
//
//   blocks 1-4  conv2d .. conv2d_15   embed_tcg.cu: one persistent kernel per block, G positions per accumulator column,
//                                     weights as the A operand, activations resident in shared memory across the block's layers
//   tail        conv2d_16 .. 19       embed_tail.cu: one launch per layer over the (clip, pool phase, row) columns of both 2x2-pool
//                                     phases, positions as M, weights resident per CTA
//   slots                             gather of the per-slot rows of the two phases (embed_fp32.cu: gather_slots_kernel)
//
// Between blocks the activations are fp16 chunk-major [clip][C / 8][T][F][8] (16-byte position records) in two ping-pong buffers of
// the caller's workspace.  A clip that is a single block-4 tile (the 1.44 s clip: 26 rows) leaves block 4 already max-pooled for
// both time phases in the tail's operand format; longer strips (streaming) leave it un-pooled and conv2d_16's staging pools.
// (Round 1's generic run-time-parameter block kernel, which ran block 4 and then only the tail, is gone: round 2 moved block 4 to the
// grouped kernel and the tail to embed_tail.cu -- 0.33 -> 0.10 ms per 8192 clips.)
#include "tc_ptx.cuh"

#include <algorithm>
#include <vector>

namespace hb {

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

// what a block reads and writes (the kernels themselves are compile-time configurations in embed_tcg.cu)
struct BlockShape {
    int first_layer, n_layers;   // conv indices [first_layer, first_layer + n_layers)
    int F;                       // freq bins inside the block
    int out_chunks;              // 8-channel chunks of its output (padded)
    int c_real;                  // real output channels
    int pool_t, pool_f;          // max-pool fused into its store
};
static const BlockShape kBlocks[4] = {
    {0, 4, 32, 4, 24, 2, 2},     // conv2d .. conv2d_3
    {4, 4, 16, 6, 48, 1, 2},     // conv2d_4 .. 7
    {8, 4, 8, 10, 72, 2, 2},     // conv2d_8 .. 11
    {12, 4, 4, 12, 96, 1, 1},    // conv2d_12 .. 15: its 2x2 pool belongs to the tail (two time phases)
};

struct Chain {
    int T_in[4], T_out[4];       // rows per clip entering / leaving each block (two VALID time convs, then the pool)
    int64_t act_bytes;           // one ping-pong buffer
};

static Chain tc_chain(int F, int B) {
    Chain c;
    int T = F;
    int64_t max_halves = 0;
    for (int b = 0; b < 4; ++b) {
        c.T_in[b] = T;
        c.T_out[b] = (T - 4) / kBlocks[b].pool_t;
        max_halves = std::max<int64_t>(max_halves, (int64_t)kBlocks[b].out_chunks * c.T_out[b] * (kBlocks[b].F / kBlocks[b].pool_f) * 8);
        T = c.T_out[b];
    }
    c.act_bytes = ((int64_t)B * max_halves * 2 + 255) & ~255ll;
    return c;
}

static int64_t tail_out_bytes(int B, int T15) { return (((int64_t)B * std::max(1, T15 / 2) * kEmbDim * 4) + 255) & ~255ll; }

int tc_prepare(hb_embed_model* m, const float* weights_host) {
    int rc = tcg_prepare(m, weights_host);
    if (rc) return rc;
    return tail_prepare(m, weights_host);
}

void tc_release(hb_embed_model* m) {
    tcg_release(m);
    tail_release(m);
}

int64_t tc_workspace_bytes(int B, int F) {
    // two fp16 ping-pong activation buffers + two f32 tail outputs + slot table
    if (F < kEmbWindow || B <= 0) return 4096;
    const Chain c = tc_chain(F, B);
    return 2 * c.act_bytes + 2 * tail_out_bytes(B, c.T_out[3]) + 8192;
}

static int check_timeout() {
    int rc = tcg_check_timeout();
    return rc ? rc : tail_check_timeout();
}

static int run_block(const hb_embed_model* m, int b, const void* in, __half* out, int B, int T_in, float* dbg, int dbg_layer, cudaStream_t st,
                     __half* pool2_out = nullptr) {
    const __half* in_h = reinterpret_cast<const __half*>(in);
    return b == 0 ? tcg_block1(m, reinterpret_cast<const float*>(in), out, B, T_in, dbg, dbg_layer, st)
         : b == 1 ? tcg_block2(m, in_h, out, B, T_in, dbg, dbg_layer, st)
         : b == 2 ? tcg_block3(m, in_h, out, B, T_in, dbg, dbg_layer, st)
                  : tcg_block4(m, in_h, out, B, T_in, dbg, dbg_layer, st, pool2_out);
}

int tc_embed_clips(const hb_embed_model* m, const float* mel, int B, int F, const int32_t* slot_offsets_host, int n_slots,
                   float* out, void* ws, int64_t ws_bytes, cudaStream_t st) {
    HB_REQUIRE(m->tcg != nullptr && m->tail != nullptr, "tensor-core weights missing");
    HB_REQUIRE(ws_bytes >= tc_workspace_bytes(B, F), "hb_embed: workspace too small");
    const Chain c = tc_chain(F, B);
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    __half* hA = reinterpret_cast<__half*>(base);
    __half* hB = reinterpret_cast<__half*>(base + c.act_bytes);
    const int T15 = c.T_out[3];
    const int64_t tail_bytes = tail_out_bytes(B, T15);
    float* tmp[2] = {reinterpret_cast<float*>(base + 2 * c.act_bytes), reinterpret_cast<float*>(base + 2 * c.act_bytes + tail_bytes)};
    int* slot_m_dev = reinterpret_cast<int*>(base + 2 * c.act_bytes + 2 * tail_bytes);

    std::vector<int> slot_m(n_slots);
    HB_REQUIRE(n_slots <= 1024, "hb_embed_clips: too many slots (%d)", n_slots);
    const int J[2] = {T15 / 2 - 4, (T15 - 1) / 2 - 4};     // outputs of pool phase 0 / 1
    const int R = T15 / 2, dup = tail_max_dup();
    // slot map for conv2d_19's epilogue: the slots fed by (phase, row); it replaces the gather pass unless a row feeds more than `dup`
    // slots or the table does not fit the workspace's 8 KB (long strips)
    std::vector<int32_t> slot_map((size_t)2 * std::max(R, 1) * dup, -1);
    bool direct = (int64_t)slot_map.size() * 4 + (int64_t)n_slots * 4 <= 8192;
    for (int s = 0; s < n_slots; ++s) {
        slot_m[s] = slot_offsets_host[s] / 4;
        const int p = slot_m[s] & 1, j = slot_m[s] >> 1;
        HB_REQUIRE(j < J[p], "hb_embed_clips: slot %d (offset %d) beyond the strip (phase %d has %d outputs)", s,
                   slot_offsets_host[s], p, J[p]);
        int32_t* e = slot_map.data() + ((size_t)p * R + j) * dup;
        int d = 0;
        while (d < dup && e[d] >= 0) ++d;
        if (d < dup) e[d] = s; else direct = false;
    }
    int32_t* slot_map_dev = reinterpret_cast<int32_t*>(slot_m_dev + n_slots);
    HB_CUDA_OK(cudaMemcpyAsync(slot_m_dev, slot_m.data(), n_slots * sizeof(int), cudaMemcpyHostToDevice, st));
    if (direct) HB_CUDA_OK(cudaMemcpyAsync(slot_map_dev, slot_map.data(), slot_map.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));

    int rc;
    if ((rc = run_block(m, 0, mel, hA, B, F, nullptr, -1, st))) return rc;
    if ((rc = run_block(m, 1, hA, hB, B, c.T_in[1], nullptr, -1, st))) return rc;
    if ((rc = run_block(m, 2, hB, hA, B, c.T_in[2], nullptr, -1, st))) return rc;
    // a clip that is one block-4 tile (the 1.44 s clip: 26 rows) leaves block 4 already pooled for both time phases
    const bool pooled = T15 <= tcg_block4_rows_per_tile();
    if ((rc = run_block(m, 3, hA, hB, B, c.T_in[3], nullptr, -1, st, pooled ? hB : nullptr))) return rc;
    // the tail (pool phases 0 and 1, conv2d_16 .. 19): one launch per layer over all columns; hA is free once block 4 has read it
    HB_REQUIRE(c.act_bytes >= tail_scratch_bytes(B, T15), "hb_embed: activation buffer smaller than the tail's scratch");
    if (direct) return tail_run(m, hB, pooled, B, T15, tmp[0], tmp[1], hA, c.act_bytes, 19, nullptr, st, slot_map_dev, out, n_slots);
    if ((rc = tail_run(m, hB, pooled, B, T15, tmp[0], tmp[1], hA, c.act_bytes, 19, nullptr, st))) return rc;
    return fp32_gather_slots(tmp[0], tmp[1], T15 / 2 - 4, T15 / 2 - 4, slot_m_dev, n_slots, out, B, st);
}

// Parity hook: the activation after conv `layer` (after its pool, phase 0 for conv2d_15) as f32 NHWC -> element count or < 0.
int64_t tc_activation(const hb_embed_model* m, const float* mel, int B, int F, int layer, float* out, int64_t cap,
                      void* ws, int64_t ws_bytes, cudaStream_t st) {
    if (!m->tcg || !m->tail || ws_bytes < tc_workspace_bytes(B, F)) {
        set_error("hb_embed_activation(f16): bad workspace");
        return HB_ERR_INVALID;
    }
    const Chain c = tc_chain(F, B);
    unsigned char* base = reinterpret_cast<unsigned char*>(ws);
    __half* bufs[2] = {reinterpret_cast<__half*>(base), reinterpret_cast<__half*>(base + c.act_bytes)};
    auto finish = [&](int64_t n) -> int64_t {
        if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) { set_error("hb_embed_activation: kernel failed"); return HB_ERR_CUDA; }
        if (check_timeout() != HB_OK) return HB_ERR_CUDA;
        return n;
    };
    const int target_block = layer > 15 ? 3 : (layer < 4 ? 0 : (layer < 8 ? 1 : (layer < 12 ? 2 : 3)));
    const void* in = mel;
    int which = 0;
    for (int b = 0; b <= target_block; ++b) {
        const BlockShape& p = kBlocks[b];
        const bool last = b == target_block && layer <= 15;
        const int last_layer_of_block = p.first_layer + p.n_layers - 1;
        if (last && layer != last_layer_of_block) {
            // an inner layer: the grouped kernel dumps it as f32 NHWC itself (block 1 counts from conv2d_1; 100 = conv2d's output)
            const int dbg_layer = layer == 0 ? 100 : layer - (b == 0 ? 1 : p.first_layer);
            int t_convs = 0;
            for (int li = p.first_layer; li <= layer; ++li) t_convs += (kLayers[li].kh == 3);
            const int T = c.T_in[b] - 2 * t_convs;
            const int64_t n = (int64_t)B * T * p.F * kLayers[layer].cout;
            if (n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
            int rc = run_block(m, b, in, bufs[which], B, c.T_in[b], out, dbg_layer, st);
            if (rc) return rc;
            return finish(n);
        }
        int rc = run_block(m, b, in, bufs[which], B, c.T_in[b], nullptr, -1, st);
        if (rc) return rc;
        if (last) {
            const int T = c.T_out[b];
            const int blocks_max = 148 * 8;
            if (b == 3) {
                // conv2d_15 "after its pool" = 2x2 pool phase 0 of the un-pooled output
                const int64_t n_pre = (int64_t)B * T * 4 * kEmbDim;
                const int64_t n = (int64_t)B * (T / 2) * 2 * kEmbDim;
                if (n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
                void* pre_mem = nullptr;
                if (cudaMalloc(&pre_mem, n_pre * 4) != cudaSuccess) { set_error("hb_embed_activation: allocation failed"); return HB_ERR_CUDA; }
                float* pre = reinterpret_cast<float*>(pre_mem);
                chunked_to_nhwc_kernel<<<(int)std::min<int64_t>(ceil_div64(n_pre, 256), blocks_max), 256, 0, st>>>(bufs[which], pre, B, 12, T, 4, kEmbDim);
                int rc2 = fp32_pool_public(pre, out, B, T, 4, kEmbDim, 2, 2, 0, st);
                const int64_t res = rc2 ? rc2 : finish(n);
                cudaFree(pre_mem);
                return res;
            }
            const int Fq = p.F / p.pool_f;
            const int64_t n = (int64_t)B * T * Fq * p.c_real;
            if (n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
            chunked_to_nhwc_kernel<<<(int)std::min<int64_t>(ceil_div64(n, 256), blocks_max), 256, 0, st>>>(bufs[which], out, B, p.out_chunks, T, Fq, p.c_real);
            return finish(n);
        }
        in = bufs[which];
        which ^= 1;
    }
    // conv2d_16 .. 19: the production tail (embed_tail.cu) on block 4's un-pooled output, stopped after `layer`; its phase-0
    // activation is dumped as f32 [B][T][96] (what spec.embedding_layer_shapes describes)
    const __half* b4 = bufs[which ^ 1];
    __half* scratch = bufs[which];
    const int T15 = c.T_out[3];
    const int64_t tail_bytes = tail_out_bytes(B, T15);
    float* tail_out = reinterpret_cast<float*>(base + 2 * c.act_bytes);
    const int rows = T15 / 2;
    const int T = layer == 16 ? rows : (layer == 19 ? rows - 4 : rows - 2);
    const int64_t n = (int64_t)B * T * kEmbDim;
    if (T <= 0 || n > cap) { set_error("hb_embed_activation: output capacity too small"); return HB_ERR_INVALID; }
    if (c.act_bytes < tail_scratch_bytes(B, T15)) { set_error("hb_embed_activation: activation buffer smaller than the tail's scratch"); return HB_ERR_INVALID; }
    int rc = tail_run(m, b4, false, B, T15, tail_out, reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(tail_out) + tail_bytes), scratch,
                      c.act_bytes, layer, out, st);
    if (rc) return rc;
    if (layer == 19 && cudaMemcpyAsync(out, tail_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("hb_embed_activation: copy failed");
        return HB_ERR_CUDA;
    }
    return finish(n);
}

}  // namespace hb

// Every tcgen05 kernel bounds its mbarrier waits (~2 s of SM clock) and raises a flag instead of hanging the GPU; this returns an error
// naming the kernel family if any flag is up.  Synchronises the device.
extern "C" int hb_check_kernels(void) {
    HB_CUDA_OK(cudaDeviceSynchronize());
    int rc = hb::tcg_check_timeout();
    if (rc == HB_OK) rc = hb::tail_check_timeout();
    if (rc == HB_OK) rc = hb::gemm_tf32_check_timeout();
    if (rc == HB_OK) rc = hb::mlp_fused_check_timeout();
    return rc;
}

// profiling aids (not part of the public header): phase timestamps of the first 8 CTAs of the last block / tail launch
extern "C" int hb_debug_tcg_times(long long* out_host) { return hb::tcg_debug_times(out_host); }
extern "C" int hb_debug_tail_times(long long* out_host) { return hb::tail_debug_times(out_host); }
