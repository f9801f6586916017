// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128) as a function of N and the smem layout of A/B.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__global__ void __launch_bounds__(128, 1) bench(int N, int layout, int iters, int same_acc, int a_tmem, int M, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
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
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 64 * 1024;
        // layout 0: no swizzle, K-major: LBO = 128*16 (next K chunk), SBO = 128;  layout 2: SW128: LBO unused(1), SBO = 1024
        const uint64_t a_hi = layout == 0 ? make_desc(0, 2048 * 16, 128, 0) : make_desc(0, 16, 1024, layout);
        const uint64_t b_hi = layout == 0 ? make_desc(0, (uint32_t)N * 16, 128, 0) : make_desc(0, 16, 1024, layout);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem_base + (same_acc ? 0 : (uint32_t)((i & 1) * 256));
            const uint64_t ad = a_hi | (uint64_t)(((a_addr + (i & 7) * 32) >> 4) & 0x3FFF);
            const uint64_t bd = b_hi | (uint64_t)((b_addr >> 4) & 0x3FFF);
            uint32_t pred;
            asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
            if (pred) {
                if (a_tmem)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(tmem_base + 320), "l"(bd), "r"(idesc), "r"(1u));
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
            }
        }
        long long t1 = clock64();
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
        if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)));
        long long t2 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}
int main() {
    long long* d; cudaMalloc(&d, 16); long long h[2];
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2000;
    for (int M : {64, 128})
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
    for (int layout : {0})
        for (int N : {64, 128, 160, 192, 224, 256})
            for (int same : {1}) {
                bench<<<1, 128, 200 * 1024>>>(N, layout, iters, same, a_tmem, M, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("M %d a_tmem %d layout %d N %3d same_acc %d : issue %.1f cyc/mma, total %.1f cyc/mma (math floor %.1f)\n", M, a_tmem, layout, N, same, h[0] / (double)iters, h[1] / (double)iters, N / 2.0 * (M == 64 ? 1.0 : 1.0));
            }
    return 0;
}
