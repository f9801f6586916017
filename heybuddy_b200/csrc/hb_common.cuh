// Shared helpers for the heybuddy_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/heybuddy_b200.h"

namespace hb {

// Thread-local last-error string behind hb_last_error().
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;  // kernels launched by this library (hb_launch_count)

#define HB_CUDA_OK(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            hb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return HB_ERR_CUDA;                                                           \
        }                                                                                 \
    } while (0)

// Checks a kernel launch and counts it.
#define HB_LAUNCHED()                        \
    do {                                     \
        HB_CUDA_OK(cudaGetLastError());      \
        hb::g_launches.fetch_add(1);         \
    } while (0)

#define HB_REQUIRE(cond, ...)                    \
    do {                                         \
        if (!(cond)) {                           \
            hb::set_error(__VA_ARGS__);          \
            return HB_ERR_INVALID;               \
        }                                        \
    } while (0)

// gemm_tf32.cu: C[M, N] = A[M, K] B[N, K]^T + bias on tcgen05 with three TF32 passes (fp32 accuracy); B rows / bias entries at or past the
// split come from the second pointer (nullptr: none)
bool gemm_tf32x3_ok(const float* A, int lda, const float* B0, const float* B1, int ldb, const float* C, int ldc, int M, int N, int K);
int gemm_tf32x3_tn(const float* A, int lda, const float* B0, const float* B1, int bsplit, int ldb, const float* bias0, const float* bias1,
                   int biassplit, float* C, int ldc, int M, int N, int K, cudaStream_t st);
int gemm_tf32_check_timeout();
int mlp_fused_check_timeout();

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- geometry of the hot path (mirrors heybuddy_b200/spec.py) ----
constexpr int kNFFT = 512;
constexpr int kWinLength = 400;
constexpr int kWinPad = (kNFFT - kWinLength) / 2;  // 56
constexpr int kHop = 160;
constexpr int kMels = 32;
constexpr int kMelBinLo = 2;    // first FFT bin with a non-zero filterbank row
constexpr int kMelBinHi = 122;  // one past the last
constexpr int kMelBand = kMelBinHi - kMelBinLo;  // 120
constexpr int kEmbWindow = 76;
constexpr int kEmbDim = 96;
constexpr int kNumConv = 20;
constexpr float kLeaky = 0.2f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace hb
