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

// Function attributes (dynamic shared-memory opt-in) are per DEVICE: CP_ONCE_PER_DEVICE runs its body the first time
// a call site is reached on each device of the process.  Idempotent bodies only (two host threads may both run one).
#include <atomic>
struct CpOncePerDevice {
    std::atomic<unsigned long long> mask{0};
    bool need(unsigned long long& bit) const {
        int d = 0;
        cudaGetDevice(&d);
        bit = 1ull << (d & 63);
        return !(mask.load(std::memory_order_acquire) & bit);
    }
    void done(unsigned long long bit) { mask.fetch_or(bit, std::memory_order_release); }
};
#define CP_ONCE_PER_DEVICE(...)                                \
    do {                                                       \
        static CpOncePerDevice once__;                         \
        unsigned long long bit__;                              \
        if (once__.need(bit__)) {                              \
            __VA_ARGS__;                                       \
            once__.done(bit__);                                \
        }                                                      \
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
// sum of w[i]^2 over [lo, hi) by one 256-thread CTA, in double: four independent chains per thread (the loads of a
// trip are in flight together), combined in a fixed order -> deterministic.  Result valid in thread 0.  Shared by
// cp_l2_forward and cp_step_prologue, which must produce the same norms bit for bit.
__device__ __forceinline__ double cta_sum_squares_256(const float* __restrict__ w, int64_t lo, int64_t hi,
                                                     double* red /*[8] shared*/) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int64_t i = lo + threadIdx.x;
    for (; i + 768 < hi; i += 1024) {
        const double a = (double)__ldg(w + i), b = (double)__ldg(w + i + 256), c = (double)__ldg(w + i + 512),
                     d = (double)__ldg(w + i + 768);
        s0 += a * a; s1 += b * b; s2 += c * c; s3 += d * d;
    }
    for (; i < hi; i += 256) {
        const double a = (double)__ldg(w + i);
        s0 += a * a;
    }
    double s = warp_sum((s0 + s1) + (s2 + s3));
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < 8; ++k) tot += red[k];
    return tot;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11): counter-based generator, 128 random
// bits per call.  Hand-rolled (about 70 integer instructions) because the dropout layers' BN-apply kernels are bound by
// the generator, not by HBM; keep-decisions compare the raw 32-bit words with an integer threshold (no float conversion).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
// keep mask of 4 consecutive elements: element index v*4.., dropout layer `layer`, P(drop) = thr / 2^32
__device__ __forceinline__ uchar4 dropout_keep4(uint64_t seed, uint64_t v, unsigned int layer, unsigned int thr) {
    const uint4 r = philox4x32_10(make_uint4((unsigned int)v, (unsigned int)(v >> 32), layer, 0x43505253u),
                                  make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
    uchar4 m;
    m.x = r.x >= thr; m.y = r.y >= thr; m.z = r.z >= thr; m.w = r.w >= thr;
    return m;
}
// threshold for drop probability p in [0, 1)
__device__ __forceinline__ unsigned int dropout_threshold(float p) {
    return (unsigned int)fminf(p * 4294967296.0f, 4294967040.0f);
}
