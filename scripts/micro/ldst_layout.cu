// Micro-test: register layout of tcgen05.ld.16x256b and of stmatrix.m8n8.trans (used by the transposing epilogue).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(uint32_t* out_ld, __half* out_mat) {
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) __half mat[2][8][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tmem_base_s;
    {   // every warp writes its lane quadrant: value = lane_global * 100 + col
        uint32_t v[8];
        for (int c = 0; c < 8; ++c) v[c] = (uint32_t)((warp * 32 + lane) * 100 + c);
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 1) {  // quadrant 1, second 16-lane half (lanes 48..63)
        uint32_t r0, r1, r2, r3;
        const uint32_t taddr = tb + ((uint32_t)(32 + 16) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        out_ld[lane * 4 + 0] = r0; out_ld[lane * 4 + 1] = r1; out_ld[lane * 4 + 2] = r2; out_ld[lane * 4 + 3] = r3;
        // stmatrix.trans: reg A = (ch = lane/4, pos 2(lane%4), +1) ; value encodes ch*10 + pos
        const int ch = lane >> 2, p = 2 * (lane & 3);
        __half2 a = __floats2half2_rn((float)(ch * 10 + p), (float)(ch * 10 + p + 1));
        __half2 b = __floats2half2_rn((float)(100 + ch * 10 + p), (float)(100 + ch * 10 + p + 1));
        const uint32_t ra = *reinterpret_cast<uint32_t*>(&a), rb = *reinterpret_cast<uint32_t*>(&b);
        // row addresses: threads 0-7 -> matrix 0 rows, threads 8-15 -> matrix 1 rows
        const uint32_t addr = smem_u32(&mat[(lane >> 3) & 1][lane & 7][0]);
        asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1, %2};" ::"r"(addr), "r"(ra), "r"(rb) : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 128; i += 128) out_mat[i] = (&mat[0][0][0])[i];
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb));
}
int main() {
    uint32_t* d; __half* m; cudaMalloc(&d, 128 * 4); cudaMalloc(&m, 256);
    k<<<1, 128>>>(d, m);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    uint32_t h[128]; __half hm[128];
    cudaMemcpy(h, d, 512, cudaMemcpyDeviceToHost); cudaMemcpy(hm, m, 256, cudaMemcpyDeviceToHost);
    for (int t = 0; t < 32; ++t) printf("thread %2d: r0 %5u r1 %5u r2 %5u r3 %5u\n", t, h[t * 4], h[t * 4 + 1], h[t * 4 + 2], h[t * 4 + 3]);
    for (int mi = 0; mi < 2; ++mi) { printf("matrix %d (rows = stored 16B rows):\n", mi); for (int r = 0; r < 8; ++r) { for (int c = 0; c < 8; ++c) printf("%5.0f", __half2float(hm[mi * 64 + r * 8 + c])); printf("\n"); } }
    return 0;
}
