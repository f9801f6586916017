// Microbenchmark: tcgen05.mma (SS, M=128, N=256, K=16) rate while other warps run an epilogue-like load:
// mode bit 0: background tcgen05.ld (16x256b.x4) loops, bit 1: background stmatrix.x4.trans loops, bit 2: LDS/STS generic traffic
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(576, 1) bench(int mode, int nbg, int iters, int a_tmem, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 576) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        stop = 0;
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
        const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 32 * 1024;
        const uint64_t a_hi = make_desc(0, 2048, 128), b_hi = make_desc(0, 9376, 128);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem_base + (uint32_t)((i & 1) * 256);
            const uint64_t ad = a_hi | (uint64_t)(((a_addr + (i % 6) * 4096) >> 4) & 0x3FFF);
            const uint64_t bd = b_hi | (uint64_t)(((b_addr + (i % 5) * 4096 + (i % 3) * 16) >> 4) & 0x3FFF);
            uint32_t pred;
            asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
            if (pred) {
                if (a_tmem) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(tmem_base + 500), "l"(bd), "r"(idesc), "r"(1u));
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u));
            }
        }
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
        if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)));
        long long t1 = clock64();
        if (lane == 0) { out[0] = t1 - t0; stop = 1; }
    } else if (warp >= 2 && warp < 2 + nbg) {
        uint32_t acc = 0;
        long long n = 0;
        const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t sa = smem_u32(smem) + 128 * 1024 + (uint32_t)((warp - 2) * 4096) + (uint32_t)((lane & 7) * 16 + (lane >> 3) * 1024);
        while (!stop) {
            uint32_t r[16];
            for (int j = 0; j < 16; ++j) r[j] = acc + j;
            if (mode & 1) {
                asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tb + (uint32_t)((n & 7) * 32)) : "memory");
            }
            if (mode & 2) {
                asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
                asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(sa + 256), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
            }
            if (mode & 4) {
                for (int j = 0; j < 8; ++j) acc += reinterpret_cast<volatile uint32_t*>(smem)[40 * 1024 + ((threadIdx.x + j * 576) & 8191)];
            }
            for (int j = 0; j < 16; ++j) acc += r[j];
            ++n;
        }
        if (lane == 0) { out[1 + warp] = n; out[40] = acc; }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}
int main() {
    long long* d; cudaMalloc(&d, 64 * 8); long long h[64];
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    const int iters = 4000;
    for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
        for (int mode : {0, 1, 2, 3, 4, 7})
            for (int nbg : {8, 16}) {
                if (mode == 0 && nbg == 16) continue;
                cudaMemset(d, 0, 64 * 8);
                bench<<<1, 576, 210 * 1024>>>(mode, mode ? nbg : 0, iters, a_tmem, d);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                long long bg = 0; for (int w = 2; w < 18; ++w) bg += h[1 + w];
                printf("a_tmem %d bg mode %d (ld=1 stsm=2 lds=4) warps %2d : %.1f cycles/mma ; background iterations %lld (%.2f per 1000 cycles)\n", a_tmem, mode, mode ? nbg : 0, h[0] / (double)iters, bg, bg * 1000.0 / (double)h[0]);
            }
    return 0;
}
