// Microbenchmark behind DESIGN.md 4.4 "why the conv stack's activations are staged LDG.128 -> STS.128 and not by TMA".
//
// The grouped tcgen05 kernels (csrc/embed_tcg.cu) want a block's input tile in "layout F": one plane per (f mod G, 8-channel
// chunk), inside a plane the records of a row are (FG + 1) columns apart -- FG = F / G real groups + ONE zero column that is the
// SAME padding of the 1x3 freq conv, which is what lets a conv tap be a shifted view of the same buffer (no im2col).  So a tile is
// C * TT * G runs of FG 16-byte records (128 B for block 2), each landing at its own, non-uniformly spaced shared-memory address.
// A tiled TMA load (cp.async.bulk.tensor) writes its box densely, so the pad column forces one copy per run; the cheapest
// TMA form of "one copy per run" is the 1-D bulk copy below (the same TMA engine, UBLKCP in SASS, no tensor map needed).
//
//   A  320 threads: LDG.128 (all of a thread's loads in flight) -> STS.128            [what the kernels do]
//   B  one warp issues C * TT * G cp.async.bulk of 128 B each, completion on an mbarrier (expect_tx = tile bytes);
//      the source is pre-laid out so that every run is contiguous (the best case for TMA)
//   C  like B with the copies issued by ALL ten warps (32 lanes each take a share)
//
// Both read an L2-resident source (the tile's rows were prefetched by the previous iteration in the real kernel), two CTAs per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tma_stage scripts/micro/tma_stage.cu && scripts/micro/tma_stage
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int C = 3, TT = 26, F = 16, G = 2, FG = F / G, PLANE = 256 * 16, THREADS = 320;
constexpr int TILE_REC = C * TT * F;                 // 1248 records of 16 B = 19,968 B
constexpr int RUNS = C * TT * G;                     // 156 runs of FG = 8 records (128 B)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// layout F address (bytes) of record (chunk c, row t, freq f)
__device__ __forceinline__ int layout_f(int c, int t, int f) { return ((f % G) * C + c) * PLANE + (1 + (FG + 1) * t + f / G) * 16; }

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) stage_kernel(const uint4* __restrict__ in, int tiles_total, long long* cycles, uint32_t* sink) {
    extern __shared__ __align__(128) unsigned char act[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0, acc = 0;
    const long long t0 = clock64();
    for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
        const uint4* src = in + (int64_t)(tile % 512) * TILE_REC;          // 512 distinct tiles = 10 MB: L2 resident
        if (MODE == 0) {
            constexpr int PER = (TILE_REC + THREADS - 1) / THREADS;
            uint4 rec[PER];
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int i = tid + k * THREADS;
                rec[k] = i < TILE_REC ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int i = tid + k * THREADS;
                if (i < TILE_REC) {
                    const int c = i / (TT * F), rem = i - c * TT * F, t = rem / F, fi = rem - t * F;
                    const int f = ((fi & 7) << 1) | (fi >> 3);            // an 8-lane group takes the even (odd) bins of a row: one plane
                    *reinterpret_cast<uint4*>(act + layout_f(c, t, f)) = rec[k];
                }
            }
        } else {
            // source pre-laid out [c][t][f mod G][f div G]: run r = (c, t, g) is 128 contiguous bytes
            const int first = tid, step = MODE == 1 ? 32 : THREADS;   // MODE 1: the 32 lanes of warp 0 share the runs
            if (tid == 0) mbar_expect_tx(&bar, TILE_REC * 16);
            __syncthreads();
            if (MODE == 2 || tid < 32)
                for (int r = first; r < RUNS; r += step) {
                    const int c = r / (TT * G), rem = r - c * TT * G, t = rem / G, g = rem - t * G;
                    bulk_g2s(act + (g * C + c) * PLANE + (1 + (FG + 1) * t) * 16, src + r * FG, FG * 16, &bar);
                }
            mbar_wait(&bar, phase & 1u);
            ++phase;
        }
        __syncthreads();
        acc += *reinterpret_cast<uint32_t*>(act + layout_f(tid % C, tid % TT, tid % F));   // consume
        __syncthreads();
    }
    if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    const int tiles_total = 296 * 64;
    uint4* in;
    long long* cyc;
    uint32_t* sink;
    cudaMalloc(&in, (size_t)512 * TILE_REC * 16);
    cudaMemset(in, 1, (size_t)512 * TILE_REC * 16);
    cudaMalloc(&cyc, 296 * sizeof(long long));
    cudaMalloc(&sink, 4);
    const size_t smem = (size_t)G * C * PLANE + 1024;
    const char* names[3] = {"A  LDG.128 -> STS.128, 320 threads", "B  156 x cp.async.bulk(128 B), one warp issues", "C  156 x cp.async.bulk(128 B), all warps issue"};
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            if (mode == 0) stage_kernel<0><<<296, THREADS, smem>>>(in, tiles_total, cyc, sink);
            if (mode == 1) stage_kernel<1><<<296, THREADS, smem>>>(in, tiles_total, cyc, sink);
            if (mode == 2) stage_kernel<2><<<296, THREADS, smem>>>(in, tiles_total, cyc, sink);
        }
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(cudaGetLastError())); return 1; }
        std::vector<long long> h(296);
        cudaMemcpy(h.data(), cyc, 296 * sizeof(long long), cudaMemcpyDeviceToHost);
        double sum = 0;
        for (long long v : h) sum += (double)v;
        printf("%-52s %8.0f cycles per tile per CTA (2 CTAs / SM, %d tiles per CTA, %d B per tile)\n", names[mode], sum / 296 / 64, 64, TILE_REC * 16);
        fflush(stdout);
    }
    return 0;
}
