// Microbenchmark: tcgen05.ld cost (cycles per load) by shape and number of concurrent warps, plus stmatrix.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(512, 1) bench(int mode, int nwarps, int iters, long long* out) {
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) unsigned char buf[16 * 1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < nwarps) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint32_t r[16];
            const uint32_t ta = tb + (uint32_t)((i & 7) * 32);
            if (mode == 0) {
                asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta) : "memory");
            } else if (mode == 1) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta) : "memory");
            } else {
                // stmatrix.x2.trans only
                for (int j = 0; j < 16; ++j) r[j] = i + j;
                const uint32_t addr = smem_u32(buf) + (uint32_t)(warp * 1024 + (lane & 15) * 16);
                asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1, %2};" ::"r"(addr), "r"(r[0]), "r"(r[1]) : "memory");
                asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1, %2};" ::"r"(addr + 256), "r"(r[2]), "r"(r[3]) : "memory");
                asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1, %2};" ::"r"(addr + 512), "r"(r[4]), "r"(r[5]) : "memory");
                asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1, %2};" ::"r"(addr + 768), "r"(r[6]), "r"(r[7]) : "memory");
            }
            for (int j = 0; j < 16; ++j) acc += r[j];
        }
        t1 = clock64();
    }
    if (lane == 0 && warp < nwarps) { out[warp * 2] = t1 - t0; out[warp * 2 + 1] = acc; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_s));
}
int main() {
    long long* d; cudaMalloc(&d, 16 * 2 * 8); long long h[32];
    const int iters = 1000;
    const char* names[3] = {"ld.16x256b.x4+wait", "ld.32x32b.x16+wait", "4x stmatrix.x2.trans"};
    for (int mode = 0; mode < 3; ++mode)
        for (int nw : {1, 4, 8, 16}) {
            bench<<<1, 512>>>(mode, nw, iters, d);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("%-22s warps %2d : %.1f cycles per iteration (warp 0), %.1f (last warp)\n", names[mode], nw, h[0] / (double)iters, h[(nw - 1) * 2] / (double)iters);
        }
    return 0;
}
