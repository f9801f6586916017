// Microbenchmark: shared-memory wavefronts of stmatrix.sync.aligned.m8n8.x4.trans by row-address pattern (run under
// ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int iters, uint32_t* sink) {
    __shared__ __align__(128) unsigned char buf[40 * 1024];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, m = lane >> 3, r = lane & 7;
    uint32_t off;
    if (MODE == 0) off = m * 4096 + r * 16;                       // 4 matrices, rows contiguous, 128-byte aligned
    else if (MODE == 1) off = m * 4096 + 16 + r * 16;             // contiguous, starts 16 B into a line
    else if (MODE == 2) off = m * 4096 + 80 + r * 16;             // contiguous, straddles a 128-byte line in the middle
    else if (MODE == 3) off = m * 4096 + (r < 3 ? 80 + r * 16 : (r == 3 ? 64 + 8192 : 12288 + (r - 4) * 16));  // rows 5,6,7 | dump | rows 0..3 of another plane, same bank base
    else if (MODE == 4) off = m * 4096 + (r < 3 ? 80 + r * 16 : (r == 3 ? 64 + 8192 : 12288 + 96 + (r - 4) * 16)); // same, other plane skewed by 96 B (the old pitch)
    else off = m * 4096 + r * 144;                                // rows 144 B apart (9 records): distinct bank groups
    const uint32_t addr = smem_u32(buf) + warp * 512 * 0 + off + (warp & 1) * 2048;
    uint32_t a = lane, b = lane + 1, c = lane + 2, d = lane + 3;
    for (int i = 0; i < iters; ++i) {
        asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
        a += i;
    }
    sink[threadIdx.x] = a + buf[threadIdx.x];
}
int main() {
    uint32_t* sink; cudaMalloc(&sink, 4096);
    k<0><<<1, 256>>>(1000, sink); k<1><<<1, 256>>>(1000, sink); k<2><<<1, 256>>>(1000, sink);
    k<3><<<1, 256>>>(1000, sink); k<4><<<1, 256>>>(1000, sink); k<5><<<1, 256>>>(1000, sink);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
