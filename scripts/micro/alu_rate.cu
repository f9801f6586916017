// Microbenchmark: issue rate (thread-instructions per clock per SM) of the epilogue's ALU instructions on sm_100a:
// cvt.rn.f16x2.f32 (F2FP.F16.F32.PACK_AB), mul.f16x2, max.NaN.f16x2, fma.f32 -- 1024 threads on one SM, 8 independent chains.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(int iters, float seed, long long* cyc, uint32_t* sink) {
    float f[8]; uint32_t h[8];
    for (int j = 0; j < 8; ++j) { f[j] = seed + threadIdx.x * 0.001f + j; h[j] = 0x3c003c00u + j; }
    const uint32_t c02 = 0x32663266u;  // 0.2, 0.2
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(f[j]), "f"(f[j])); f[j] = __uint_as_float(h[j] | 0x30003000u); }
            else if (MODE == 1) asm volatile("mul.f16x2 %0, %0, %1;" : "+r"(h[j]) : "r"(c02));
            else if (MODE == 2) asm volatile("max.NaN.f16x2 %0, %0, %1;" : "+r"(h[j]) : "r"(c02));
            else if (MODE == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[j]) : "f"(seed));
            else if (MODE == 4) {  // the epilogue's leaky_half2: cvt + mul + max
                uint32_t t, m;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(f[j]), "f"(f[j]));
                asm volatile("mul.f16x2 %0, %1, %2;" : "=r"(m) : "r"(t), "r"(c02));
                asm volatile("max.NaN.f16x2 %0, %1, %2;" : "=r"(h[j]) : "r"(t), "r"(m));
                f[j] = __uint_as_float(h[j]);
            } else if (MODE == 5) {  // fp32 leaky then pack: mul + max (x2) + cvt
                float a = f[j], b = f[j], a2, b2;
                asm volatile("mul.f32 %0, %1, %2;" : "=f"(a2) : "f"(a), "f"(seed));
                asm volatile("mul.f32 %0, %1, %2;" : "=f"(b2) : "f"(b), "f"(seed));
                asm volatile("max.NaN.f32 %0, %0, %1;" : "+f"(a2) : "f"(a));
                asm volatile("max.NaN.f32 %0, %0, %1;" : "+f"(b2) : "f"(b));
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[j]) : "f"(a2), "f"(b2));
                f[j] = __uint_as_float(h[j]);
            } else if (MODE == 6) {  // fp32 scale, two packs, one packed max: 2 FMUL + 2 cvt + 1 max
                float a = f[j], a2; uint32_t t, m;
                asm volatile("mul.f32 %0, %1, %2;" : "=f"(a2) : "f"(a), "f"(seed));
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(t) : "f"(a), "f"(a));
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(m) : "f"(a2), "f"(a2));
                asm volatile("max.NaN.f16x2 %0, %1, %2;" : "=r"(h[j]) : "r"(t), "r"(m));
                f[j] = __uint_as_float(h[j]);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
    for (int j = 0; j < 8; ++j) acc ^= h[j] ^ __float_as_uint(f[j]);
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    sink[threadIdx.x] = acc;
}
int main() {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4096);
    const int iters = 4096;
    const char* names[] = {"cvt.rn.f16x2.f32", "mul.f16x2", "max.NaN.f16x2", "fma.f32", "leaky_half2 (cvt+mul+max)", "fp32 leaky + cvt (5 instr)", "fmul + 2 cvt + max (4 instr)"};
    const int per[] = {1, 1, 1, 1, 3, 5, 4};
    for (int mode = 0; mode < 7; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
                case 0: k<0><<<1, 1024>>>(iters, 1.5f, cyc, sink); break;
                case 1: k<1><<<1, 1024>>>(iters, 1.5f, cyc, sink); break;
                case 2: k<2><<<1, 1024>>>(iters, 1.5f, cyc, sink); break;
                case 3: k<3><<<1, 1024>>>(iters, 1.5f, cyc, sink); break;
                case 4: k<4><<<1, 1024>>>(iters, 1.5f, cyc, sink); break;
                case 5: k<5><<<1, 1024>>>(iters, 1.5f, cyc, sink); break;
                default: k<6><<<1, 1024>>>(iters, 1.5f, cyc, sink); break;
            }
            cudaDeviceSynchronize();
        }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double n = (double)iters * 8 * per[mode] * 1024;
        printf("%-28s %.1f thread-instr / clk / SM  (%.2f cycles per warp-instr per SMSP)\n", names[mode], n / c, c / (n / 32 / 4));
    }
    return 0;
}
