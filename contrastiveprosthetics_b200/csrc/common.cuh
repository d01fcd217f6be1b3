// Shared device/host helpers for libcpros (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cpros.h"

#define CP_NUM_SMS 148   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// every kernel launch of the library goes through this macro: it also feeds cp_launch_count()
extern unsigned long long g_cp_launches;
#define CP_CHECK_LAUNCH()                                  \
    do {                                                   \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return (int)e__;           \
        ++g_cp_launches;                                   \
    } while (0)

#define CP_CUDA(call)                                      \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

static inline int64_t cp_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t cp_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// packed fp32 FMA (Blackwell FFMA2): d.x += a.x*b.x, d.y += a.y*b.y in ONE instruction -- twice the fp32 rate of
// scalar FFMA, which issues every other cycle per scheduler on sm_100
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long& dd = reinterpret_cast<unsigned long long&>(d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(dd)
        : "l"(reinterpret_cast<const unsigned long long&>(a)), "l"(reinterpret_cast<const unsigned long long&>(b)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
