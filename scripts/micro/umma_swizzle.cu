// Microbenchmark: tcgen05.mma (SS, M=128, N=256, K=16, f16) issue rate vs shared-memory layout type of the operands
// (descriptor bits [61,64): 0 none/interleave, 6 = 32B swizzle, 4 = 64B, 2 = 128B).  Data values are irrelevant here.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
           ((uint64_t)layout << 61);
}
__global__ void __launch_bounds__(128, 1) bench(int layout, int n, int iters, uint32_t lbo, uint32_t sbo, uint32_t astep, uint32_t bstep, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 64 * 1024;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem_base + (uint32_t)((i & 1) * 256);
            const uint64_t ad = make_desc(a_addr + (i % 4) * astep, lbo, sbo, layout);
            const uint64_t bd = make_desc(b_addr + (i % 4) * bstep, lbo, sbo, layout);
            uint32_t pred;
            asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
            if (pred) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
        }
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
        if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)));
        long long t1 = clock64();
        if (lane == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}
int main() {
    long long* d; cudaMalloc(&d, 64); long long h[8];
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    const int iters = 4000;
    struct V { const char* name; int layout; uint32_t lbo, sbo, astep, bstep; };
    // K-major operands, 16 K elements per MMA.  none: core matrix 8 rows x 16 B, LBO = K-chunk stride, SBO = 128.
    // swizzled: row pitch = swizzle width, SBO = 8 rows; successive K steps advance the start address by 32 B inside the row.
    V vs[] = {{"none  (LBO 4096, SBO 128)", 0, 4096, 128, 8192, 8192},
              {"none  (LBO 2048+16 skew) ", 0, 2064, 128, 8192, 8192},
              {"none  B start +16 B      ", 0, 4096, 128, 8192, 8192 + 16},
              {"none  B start +64 B      ", 0, 4096, 128, 8192, 8192 + 64},
              {"none  LBO 9376           ", 0, 9376, 128, 8192, 8192},
              {"none  LBO 4096+32        ", 0, 4128, 128, 8192, 8192},
              {"none  A,B start +16 B    ", 0, 4096, 128, 8192 + 16, 8192 + 16},
              {"32B swizzle (SBO 256)    ", 6, 16, 256, 8192, 8192},
              {"64B swizzle (SBO 512)    ", 4, 16, 512, 32, 32},
              {"128B swizzle (SBO 1024)  ", 2, 16, 1024, 32, 32}};
    for (auto& v : vs)
        for (int n : {128, 256}) {
            cudaMemset(d, 0, 64);
            bench<<<1, 128, 210 * 1024>>>(v.layout, n, iters, v.lbo, v.sbo, v.astep, v.bstep, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s N=%d: error %s\n", v.name, n, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
            printf("%s N=%3d : %.1f cycles/mma\n", v.name, n, h[0] / (double)iters);
        }
    return 0;
}
