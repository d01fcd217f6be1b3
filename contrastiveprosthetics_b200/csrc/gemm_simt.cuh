// fp32 FFMA GEMM family for the encoder (the bit-faithful "parity" engine).
//
//   gemm_nt   : C[M,N] = act(A[M,K] . B^T + bias), B given as [N,K] (BMODE 0) or [K,N] (BMODE 1),
//               optional conv-view of A (1-D k=3 convolution over 12 positions as an implicit
//               GEMM, K = 3*64), fused bias + ReLU + per-column sum / sum-of-squares partials
//               (the BatchNorm statistics of models.py:17-35 without re-reading the activation).
//   gemm_tn   : C[z][Mo,No] = sum_r G[r,Mo] * A[r,No] over a row slice z (split-K weight gradient),
//               optional conv-view of A.
//
// Tiling: 256 threads, BK=16, double-buffered shared memory with register prefetch, each thread
// owns (BM/16)x(BN/16) outputs as 4x4 micro-tiles 64 apart (conflict-free float4 LDS).
#pragma once
#include "common.cuh"

#define GEMM_BK 16
#define CONV_POS 12      // EMG_DIM positions
#define CONV_CH 64

struct GemmNT {
    const float* A; int64_t M; int K; int lda;
    const float* B; int N; int ldb;
    const float* bias; float* C; int ldc;
    float* psum; float* psq;      // [ceil(M/BM)][N] partial column statistics, or null
    int relu;
};

// conv-view: logical A[r, tap*64 + c] = base[(r + tap - 1)*64 + c], zero outside the window of 12
__device__ __forceinline__ bool conv_valid(int64_t r, int k) {
    const int p = (int)(r % CONV_POS);
    return !((p == 0 && k < CONV_CH) || (p == CONV_POS - 1 && k >= 2 * CONV_CH));
}

template <int BM, int BN, int BMODE, bool ACONV>
__global__ void __launch_bounds__(256, 2)
gemm_nt_kernel(const GemmNT g) {
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int A_LD = (BM * GEMM_BK / 4) / 256;                   // float4 loads per thread
    constexpr int B_LD = (BN * GEMM_BK / 4 + 255) / 256;
    __shared__ __align__(16) float As[2][GEMM_BK][BM + 4];
    __shared__ __align__(16) float Bs[2][GEMM_BK][BN + 4];

    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    // 1-D grid, N tiles fastest: the CTAs that share an A row-panel are co-resident, so the panel
    // is fetched from HBM once and re-read from L2
    const int tiles_n = g.N / BN;
    const int64_t tile_m = blockIdx.x / tiles_n;
    const int64_t m0 = tile_m * BM;
    const int n0 = (int)(blockIdx.x % tiles_n) * BN;
    const int KT = g.K / GEMM_BK;

    float4 ra[A_LD], rb[B_LD];
    auto load_regs = [&](int kt) {
        const int k0 = kt * GEMM_BK;
#pragma unroll
        for (int q = 0; q < A_LD; ++q) {
            const int f = tid + q * 256, row = f / 4, kq = (f % 4) * 4;
            const int64_t r = m0 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < g.M) {
                if (ACONV) {
                    if (conv_valid(r, k0 + kq))
                        v = __ldg(reinterpret_cast<const float4*>(g.A + (r - 1) * CONV_CH + k0 + kq));
                } else {
                    v = __ldg(reinterpret_cast<const float4*>(g.A + r * g.lda + k0 + kq));
                }
            }
            ra[q] = v;
        }
#pragma unroll
        for (int q = 0; q < B_LD; ++q) {
            const int f = tid + q * 256;
            if (BMODE == 0) {
                const int row = f / 4, kq = (f % 4) * 4;
                if (row < BN)
                    rb[q] = __ldg(reinterpret_cast<const float4*>(g.B + (int64_t)(n0 + row) * g.ldb + k0 + kq));
            } else {
                const int kr = f / (BN / 4), nq = (f % (BN / 4)) * 4;
                if (kr < GEMM_BK)
                    rb[q] = __ldg(reinterpret_cast<const float4*>(g.B + (int64_t)(k0 + kr) * g.ldb + n0 + nq));
            }
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int q = 0; q < A_LD; ++q) {
            const int f = tid + q * 256, row = f / 4, kq = (f % 4) * 4;
            As[buf][kq + 0][row] = ra[q].x; As[buf][kq + 1][row] = ra[q].y;
            As[buf][kq + 2][row] = ra[q].z; As[buf][kq + 3][row] = ra[q].w;
        }
#pragma unroll
        for (int q = 0; q < B_LD; ++q) {
            const int f = tid + q * 256;
            if (BMODE == 0) {
                const int row = f / 4, kq = (f % 4) * 4;
                if (row < BN) {
                    Bs[buf][kq + 0][row] = rb[q].x; Bs[buf][kq + 1][row] = rb[q].y;
                    Bs[buf][kq + 2][row] = rb[q].z; Bs[buf][kq + 3][row] = rb[q].w;
                }
            } else {
                const int kr = f / (BN / 4), nq = (f % (BN / 4)) * 4;
                if (kr < GEMM_BK) *reinterpret_cast<float4*>(&Bs[buf][kr][nq]) = rb[q];
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_regs(0);
    store_smem(0);
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) load_regs(kt + 1);
#pragma unroll
        for (int k = 0; k < GEMM_BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int h = 0; h < TM / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][h * 64 + ty * 4]);
                a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < TN / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][h * 64 + tx * 4]);
                b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < KT) store_smem(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: bias, ReLU, store, column statistics
    float csum[TN], csq[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) { csum[j] = 0.f; csq[j] = 0.f; }
#pragma unroll
    for (int hj = 0; hj < TN / 4; ++hj) {
        const int col = n0 + hj * 64 + tx * 4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.bias) bv = __ldg(reinterpret_cast<const float4*>(g.bias + col));
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int64_t r = m0 + (i / 4) * 64 + ty * 4 + (i % 4);
            float v[4] = {acc[i][hj * 4 + 0] + bv.x, acc[i][hj * 4 + 1] + bv.y,
                          acc[i][hj * 4 + 2] + bv.z, acc[i][hj * 4 + 3] + bv.w};
            if (g.relu) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (r < g.M) {
                *reinterpret_cast<float4*>(g.C + r * g.ldc + col) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    csum[hj * 4 + j] += v[j];
                    csq[hj * 4 + j] = fmaf(v[j], v[j], csq[hj * 4 + j]);
                }
            }
        }
    }
    if (g.psum) {
        // [16][BN] scratch in the (now idle) tile buffers; the main loop ended with a barrier
        static_assert(BM >= BN, "scratch reuse assumes the A tile is at least as wide as BN");
        float* red_s = &As[0][0][0];
        float* red_q = &Bs[0][0][0];
#pragma unroll
        for (int hj = 0; hj < TN / 4; ++hj)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                red_s[ty * BN + hj * 64 + tx * 4 + j] = csum[hj * 4 + j];
                red_q[ty * BN + hj * 64 + tx * 4 + j] = csq[hj * 4 + j];
            }
        __syncthreads();
        if (tid < BN) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int y = 0; y < 16; ++y) { s += red_s[y * BN + tid]; q += red_q[y * BN + tid]; }
            g.psum[tile_m * g.N + n0 + tid] = s;
            g.psq[tile_m * g.N + n0 + tid] = q;
        }
    }
}

// --------------------------------------------------------------------------------- weight grad
struct GemmTN {
    const float* G; int ldg; int Mo;      // [R, ldg], output rows = columns of G
    const float* A; int lda; int No;      // [R, lda] (or conv-view base [R,64], No = 192)
    int64_t R; int64_t rows_per_split;
    float* P;                             // [gridDim.z][Mo][No]
    const float* G_lo; const float* A_lo; // optional low planes (operands stored split as hi + lo)
};

__device__ __forceinline__ float4 ld_planes(const float* hi, const float* lo, int64_t off) {
    float4 v = __ldg(reinterpret_cast<const float4*>(hi + off));
    if (lo) {
        const float4 l = __ldg(reinterpret_cast<const float4*>(lo + off));
        v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
    }
    return v;
}

template <int BM, int BN, bool ACONV>
__global__ void __launch_bounds__(256, 2)
gemm_tn_kernel(const GemmTN g) {
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int G_LD = (BM * GEMM_BK / 4 + 255) / 256;
    constexpr int A_LD = (BN * GEMM_BK / 4 + 255) / 256;
    __shared__ __align__(16) float Gs[2][GEMM_BK][BM + 4];
    __shared__ __align__(16) float As[2][GEMM_BK][BN + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int o0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
    const int64_t r_begin = (int64_t)blockIdx.z * g.rows_per_split;
    const int64_t r_end = min(g.R, r_begin + g.rows_per_split);
    const int KT = (int)((r_end - r_begin + GEMM_BK - 1) / GEMM_BK);

    float4 rg[G_LD], ra[A_LD];
    auto load_regs = [&](int kt) {
        const int64_t rb = r_begin + (int64_t)kt * GEMM_BK;
#pragma unroll
        for (int q = 0; q < G_LD; ++q) {
            const int f = tid + q * 256, kr = f / (BM / 4), cq = (f % (BM / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kr < GEMM_BK && rb + kr < r_end) v = ld_planes(g.G, g.G_lo, (rb + kr) * g.ldg + o0 + cq);
            rg[q] = v;
        }
#pragma unroll
        for (int q = 0; q < A_LD; ++q) {
            const int f = tid + q * 256, kr = f / (BN / 4), cq = (f % (BN / 4)) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int64_t r = rb + kr;
            if (kr < GEMM_BK && r < r_end) {
                if (ACONV) {
                    if (conv_valid(r, c0 + cq)) v = ld_planes(g.A, g.A_lo, (r - 1) * CONV_CH + c0 + cq);
                } else {
                    v = ld_planes(g.A, g.A_lo, r * g.lda + c0 + cq);
                }
            }
            ra[q] = v;
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int q = 0; q < G_LD; ++q) {
            const int f = tid + q * 256, kr = f / (BM / 4), cq = (f % (BM / 4)) * 4;
            if (kr < GEMM_BK) *reinterpret_cast<float4*>(&Gs[buf][kr][cq]) = rg[q];
        }
#pragma unroll
        for (int q = 0; q < A_LD; ++q) {
            const int f = tid + q * 256, kr = f / (BN / 4), cq = (f % (BN / 4)) * 4;
            if (kr < GEMM_BK) *reinterpret_cast<float4*>(&As[buf][kr][cq]) = ra[q];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    if (KT > 0) {
        load_regs(0);
        store_smem(0);
    }
    __syncthreads();
    for (int kt = 0; kt < KT; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < KT) load_regs(kt + 1);
#pragma unroll
        for (int k = 0; k < GEMM_BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int h = 0; h < (TM + 3) / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(&Gs[buf][k][h * 64 + ty * 4]);
                a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < (TN + 3) / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][h * 64 + tx * 4]);
                b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < KT) store_smem(buf ^ 1);
        __syncthreads();
    }
    float* P = g.P + (int64_t)blockIdx.z * g.Mo * g.No;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int o = o0 + (i / 4) * 64 + ty * 4 + (i % 4);
#pragma unroll
        for (int hj = 0; hj < TN / 4; ++hj) {
            const int c = c0 + hj * 64 + tx * 4;
            *reinterpret_cast<float4*>(P + (int64_t)o * g.No + c) =
                make_float4(acc[i][hj * 4 + 0], acc[i][hj * 4 + 1], acc[i][hj * 4 + 2], acc[i][hj * 4 + 3]);
        }
    }
}
