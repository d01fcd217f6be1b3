// Non-GEMM kernels of the EMG encoder: k=3 conv on 1 input channel, BatchNorm/AdaBN statistics
// finalisation, BN apply (+dropout), BN backward (reduce / apply fused with the ReLU mask and the
// bias gradient), the 512->16 projection, weight re-layouts.  All fp32; all HBM-bound: each
// activation is read/written in 16-byte lanes, per-column reductions use register accumulators
// + one shared-memory step and leave deterministic per-CTA partials (no float atomics).
//
// Activation layout: [rows, F] row-major, F = 64 (conv stages, rows = windows*12, channel
// contiguous) or F = 512 (linear stages, rows = windows).
#pragma once
#include "common.cuh"

template <int F> struct ColMap {
    static constexpr int QX = F / 4;            // float4 column groups
    static constexpr int RY = 256 / QX;         // row lanes per CTA
    static constexpr int ROWS = F == 512 ? 128 : (F == 256 ? 256 : 1024);   // rows per CTA (64 per thread)
};

// sum over the RY row lanes of a CTA; result for column c valid in threads with ry == 0
template <int F>
__device__ __forceinline__ void block_col_reduce(float (&v)[4], float* red /*[RY][F]*/, int qx, int ry) {
    constexpr int RY = ColMap<F>::RY;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[ry * F + qx * 4 + j] = v[j];
    __syncthreads();
    if (ry == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int y = 0; y < RY; ++y) s += red[y * F + qx * 4 + j];
            v[j] = s;
        }
    }
}

// ------------------------------------------------------------------------------------ conv1
// y1[(n,p), c] = relu(b[c] + sum_tap w[c,tap] * x[n, p+tap-1])   (models.py:255-256; a 3x3 conv
// with padding 1 on a 1x12 image only ever sees the middle kernel row)
//
// The stage costs 3 FMAs per output but its activation is 3 KB per window, so y1 is NEVER stored: every
// kernel that needs it (statistics, BN apply, BN backward) recomputes it from the 48-byte window -- the same
// fmaf chain each time, hence bit-identical values in every pass.
struct Conv1Taps { float w[4][3], b[4]; };          // 4 consecutive output channels
__device__ __forceinline__ Conv1Taps conv1_taps(const float* __restrict__ w9, const float* __restrict__ bias, int qx) {
    Conv1Taps t;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = qx * 4 + j;
        t.b[j] = __ldg(bias + c);
#pragma unroll
        for (int k = 0; k < 3; ++k) t.w[j][k] = __ldg(w9 + c * 9 + 3 + k);
    }
    return t;
}
// Thread layout of the conv1 kernels: 16 threads per window (one per group of 4 channels); a thread holds the
// 12 samples of its window in registers and walks the 12 positions, so there is no index arithmetic per element
// and a warp touches 2 x 256 contiguous bytes of the [n*12, 64] activation per position.
#define C1_WIN 64                                     // windows per CTA of the slab (statistics / backward) kernels
struct Conv1Window { float x[14]; };                // x[0] = x[13] = 0: the padding of models.py:255
__device__ __forceinline__ Conv1Window conv1_window(const float* __restrict__ x, int64_t w) {
    Conv1Window v;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + w * 12));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + w * 12) + 1);
    const float4 c = __ldg(reinterpret_cast<const float4*>(x + w * 12) + 2);
    v.x[0] = 0.f; v.x[13] = 0.f;
    v.x[1] = a.x; v.x[2] = a.y; v.x[3] = a.z; v.x[4] = a.w;
    v.x[5] = b.x; v.x[6] = b.y; v.x[7] = b.z; v.x[8] = b.w;
    v.x[9] = c.x; v.x[10] = c.y; v.x[11] = c.z; v.x[12] = c.w;
    return v;
}
// relu(conv1) at position p (taps x[p-1], x[p], x[p+1] = v.x[p], v.x[p+1], v.x[p+2]) for the thread's 4 channels
__device__ __forceinline__ void conv1_eval(const Conv1Taps& t, const Conv1Window& v, int p, float (&y)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = t.b[j];
        a = fmaf(t.w[j][0], v.x[p], a);
        a = fmaf(t.w[j][1], v.x[p + 1], a);
        a = fmaf(t.w[j][2], v.x[p + 2], a);
        y[j] = fmaxf(a, 0.f);
    }
}
// sum over the 16 window lanes of a CTA; result for the 4 channels of group qx valid in threads with wl == 0
__device__ __forceinline__ void conv1_block_reduce(float (&v)[4], float* red /*[16][64]*/, int qx, int wl) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[wl * 64 + qx * 4 + j] = v[j];
    __syncthreads();
    if (wl == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int y = 0; y < 16; ++y) s += red[y * 64 + qx * 4 + j];
            v[j] = s;
        }
    }
}

// statistics partials of relu(conv1(x)) (psum/psq non-null, [ceil(n/C1_WIN)][64]) and/or the activation itself
// (y non-null: parity tap)
__global__ void __launch_bounds__(256)
conv1_fwd_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ w9,
                 const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ psum,
                 float* __restrict__ psq) {
    __shared__ float red[16 * 64];
    const int qx = threadIdx.x % 16, wl = threadIdx.x / 16;
    const Conv1Taps t = conv1_taps(w9, bias, qx);
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    for (int wi = wl; wi < C1_WIN; wi += 16) {
        const int64_t w = (int64_t)blockIdx.x * C1_WIN + wi;
        if (w >= n) break;
        const Conv1Window xv = conv1_window(x, w);
#pragma unroll
        for (int p = 0; p < 12; ++p) {
            float v[4];
            conv1_eval(t, xv, p, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s[j] += v[j];
                q[j] = fmaf(v[j], v[j], q[j]);
            }
            if (y) reinterpret_cast<float4*>(y + (w * 12 + p) * 64)[qx] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    if (!psum) return;
    conv1_block_reduce(s, red, qx, wl);
    conv1_block_reduce(q, red, qx, wl);
    if (wl == 0) {
        *reinterpret_cast<float4*>(psum + (int64_t)blockIdx.x * 64 + qx * 4) = make_float4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<float4*>(psq + (int64_t)blockIdx.x * 64 + qx * 4) = make_float4(q[0], q[1], q[2], q[3]);
    }
}

// a1 = relu(conv1(x)) * scale + shift, written as fp32 or (SPLIT) as the two fp16 planes of the tensor-core engine
template <bool SPLIT>
__global__ void __launch_bounds__(256)
conv1_bn_apply_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ w9,
                      const float* __restrict__ bias, const float* __restrict__ scale,
                      const float* __restrict__ shift, float* __restrict__ a, float* __restrict__ a_lo,
                      const unsigned int* __restrict__ abound = nullptr, float* __restrict__ ascale_inv = nullptr) {
    const int qx = threadIdx.x % 16;
    const Conv1Taps t = conv1_taps(w9, bias, qx);
    float4 s = __ldg(reinterpret_cast<const float4*>(scale + qx * 4));
    float4 h = __ldg(reinterpret_cast<const float4*>(shift + qx * 4));
    if (SPLIT) {        // planes of a1 * S (power of two: exact), S from the stage's BatchNorm output bound
        const float S = plane_scale(__uint_as_float(__ldg(abound)));
        s.x *= S; s.y *= S; s.z *= S; s.w *= S;
        h.x *= S; h.y *= S; h.z *= S; h.w *= S;
        if (blockIdx.x == 0 && threadIdx.x == 0) *ascale_inv = 1.f / S;
    }
    for (int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 16; w < n;
         w += (int64_t)gridDim.x * (blockDim.x / 16)) {
        const Conv1Window xv = conv1_window(x, w);
#pragma unroll
        for (int p = 0; p < 12; ++p) {
            float y[4];
            conv1_eval(t, xv, p, y);
            const float4 o = make_float4(fmaf(y[0], s.x, h.x), fmaf(y[1], s.y, h.y), fmaf(y[2], s.z, h.z), fmaf(y[3], s.w, h.w));
            const int64_t v = (w * 12 + p) * 16 + qx;
            if (SPLIT) split_store4(o, reinterpret_cast<plane_t*>(a), reinterpret_cast<plane_t*>(a_lo), v);
            else reinterpret_cast<float4*>(a)[v] = o;
        }
    }
}

// BN backward of the conv1 stage, pass 1: partials of sum g and sum g*xh with xh from the recomputed activation
__global__ void __launch_bounds__(256)
conv1_bn_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ x, int64_t n,
                           const float* __restrict__ w9, const float* __restrict__ bias,
                           const float* __restrict__ mean, const float* __restrict__ istd,
                           float* __restrict__ p1, float* __restrict__ p2) {
    __shared__ float red[16 * 64];
    const int qx = threadIdx.x % 16, wl = threadIdx.x / 16;
    const Conv1Taps t = conv1_taps(w9, bias, qx);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + qx * 4));
    const float4 is = __ldg(reinterpret_cast<const float4*>(istd + qx * 4));
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    for (int wi = wl; wi < C1_WIN; wi += 16) {
        const int64_t w = (int64_t)blockIdx.x * C1_WIN + wi;
        if (w >= n) break;
        const Conv1Window xv = conv1_window(x, w);
        float4 gv[12];
#pragma unroll
        for (int p = 0; p < 12; ++p) gv[p] = __ldg(reinterpret_cast<const float4*>(g + (w * 12 + p) * 64) + qx);
#pragma unroll
        for (int p = 0; p < 12; ++p) {
            float y[4];
            conv1_eval(t, xv, p, y);
            s1[0] += gv[p].x; s1[1] += gv[p].y; s1[2] += gv[p].z; s1[3] += gv[p].w;
            s2[0] = fmaf(gv[p].x, (y[0] - mu.x) * is.x, s2[0]);
            s2[1] = fmaf(gv[p].y, (y[1] - mu.y) * is.y, s2[1]);
            s2[2] = fmaf(gv[p].z, (y[2] - mu.z) * is.z, s2[2]);
            s2[3] = fmaf(gv[p].w, (y[3] - mu.w) * is.w, s2[3]);
        }
    }
    conv1_block_reduce(s1, red, qx, wl);
    conv1_block_reduce(s2, red, qx, wl);
    if (wl == 0) {
        *reinterpret_cast<float4*>(p1 + (int64_t)blockIdx.x * 64 + qx * 4) = make_float4(s1[0], s1[1], s1[2], s1[3]);
        *reinterpret_cast<float4*>(p2 + (int64_t)blockIdx.x * 64 + qx * 4) = make_float4(s2[0], s2[1], s2[2], s2[3]);
    }
}

// pass 2: gz = 1[y>0] * gamma*istd * (g - m1 - xh*m2) stays in registers (the first layer has no data gradient);
// partials of the bias gradient sum gz -> pdb[blk][64] and of the weight gradient
// dW1[c,tap] = sum_(w,p) gz[(w,p),c] * x[w, p+tap-1] -> pdw[blk][tap][64]
__global__ void __launch_bounds__(256)
conv1_bn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ x, int64_t n,
                          const float* __restrict__ w9, const float* __restrict__ bias,
                          const float* __restrict__ mean, const float* __restrict__ istd,
                          const float* __restrict__ gamma, const float* __restrict__ m1,
                          const float* __restrict__ m2, float* __restrict__ pdb, float* __restrict__ pdw) {
    __shared__ float red[16 * 64];
    const int qx = threadIdx.x % 16, wl = threadIdx.x / 16;
    const Conv1Taps t = conv1_taps(w9, bias, qx);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + qx * 4));
    const float4 is = __ldg(reinterpret_cast<const float4*>(istd + qx * 4));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + qx * 4));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(m1 + qx * 4));
    const float4 a2 = __ldg(reinterpret_cast<const float4*>(m2 + qx * 4));
    const float k0 = ga.x * is.x, k1 = ga.y * is.y, k2 = ga.z * is.z, k3 = ga.w * is.w;
    float sb[4] = {0, 0, 0, 0}, w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0}, w2[4] = {0, 0, 0, 0};
    for (int wi = wl; wi < C1_WIN; wi += 16) {
        const int64_t w = (int64_t)blockIdx.x * C1_WIN + wi;
        if (w >= n) break;
        const Conv1Window xv = conv1_window(x, w);
        float4 gv[12];
#pragma unroll
        for (int p = 0; p < 12; ++p) gv[p] = __ldg(reinterpret_cast<const float4*>(g + (w * 12 + p) * 64) + qx);
#pragma unroll
        for (int p = 0; p < 12; ++p) {
            float y[4], o[4];
            conv1_eval(t, xv, p, y);
            o[0] = y[0] > 0.f ? k0 * (gv[p].x - a1.x - (y[0] - mu.x) * is.x * a2.x) : 0.f;
            o[1] = y[1] > 0.f ? k1 * (gv[p].y - a1.y - (y[1] - mu.y) * is.y * a2.y) : 0.f;
            o[2] = y[2] > 0.f ? k2 * (gv[p].z - a1.z - (y[2] - mu.z) * is.z * a2.z) : 0.f;
            o[3] = y[3] > 0.f ? k3 * (gv[p].w - a1.w - (y[3] - mu.w) * is.w * a2.w) : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sb[j] += o[j];
                w0[j] = fmaf(o[j], xv.x[p], w0[j]);
                w1[j] = fmaf(o[j], xv.x[p + 1], w1[j]);
                w2[j] = fmaf(o[j], xv.x[p + 2], w2[j]);
            }
        }
    }
    conv1_block_reduce(sb, red, qx, wl);
    conv1_block_reduce(w0, red, qx, wl);
    conv1_block_reduce(w1, red, qx, wl);
    conv1_block_reduce(w2, red, qx, wl);
    if (wl == 0) {
        *reinterpret_cast<float4*>(pdb + (int64_t)blockIdx.x * 64 + qx * 4) = make_float4(sb[0], sb[1], sb[2], sb[3]);
        float* o = pdw + (int64_t)blockIdx.x * 3 * 64;
        *reinterpret_cast<float4*>(o + 0 * 64 + qx * 4) = make_float4(w0[0], w0[1], w0[2], w0[3]);
        *reinterpret_cast<float4*>(o + 1 * 64 + qx * 4) = make_float4(w1[0], w1[1], w1[2], w1[3]);
        *reinterpret_cast<float4*>(o + 2 * 64 + qx * 4) = make_float4(w2[0], w2[1], w2[2], w2[3]);
    }
}

// ------------------------------------------------------------------------ partial reductions
// Sum P per-CTA partial rows of width `width` in double.  CTA = 32 columns x 32 partial lanes.
__device__ __forceinline__ double reduce_partials(const float* __restrict__ part, int P, int width,
                                                  int col, int lane, double* sm /*[32][33]*/) {
    double s = 0.0;
    if (col < width)
        for (int p = lane; p < P; p += 32) s += (double)__ldg(part + (int64_t)p * width + col);
    const int cx = threadIdx.x % 32;
    sm[lane * 33 + cx] = s;
    __syncthreads();
    double t = 0.0;
    if (lane == 0)
        for (int l = 0; l < 32; ++l) t += sm[l * 33 + cx];
    __syncthreads();
    return t;      // valid for lane == 0
}

// Two-level variant for long partial lists: grid = (ceil(width/32), RP_SLABS); every CTA reduces its slab
// of partial rows into `scratch[slab][q][width]` (double), the last CTA of a column group to finish
// (atomic ticket) adds the RP_SLABS slab sums in a fixed order -> deterministic.  Returns true in the
// threads (lane == 0, col < width) of that last CTA, with the totals in out[0..NQ).
#define RP_SLABS 16
// slabs (grid.y) for a list of P partial rows: one CTA per column group walks up to 64 rows by itself (32 lanes x 2) --
// no scratch round trip, no ticket -- which is every finalize of a small batch (328 windows: 3 .. 31 partial rows)
inline int rp_slabs(int P) { return P <= 64 ? 1 : RP_SLABS; }
template <int NQ>
__device__ __forceinline__ bool reduce_partials_2level(const float* const (&part)[NQ], int P, int width,
                                                       double* __restrict__ scratch, unsigned int* __restrict__ tickets,
                                                       double (&out)[NQ], double* sm /*[32*33]*/) {
    __shared__ bool is_last;
    const int cx = threadIdx.x % 32, lane = threadIdx.x / 32;
    const int col = blockIdx.x * 32 + cx;
    const int slab = blockIdx.y;
    const int nslab = (int)gridDim.y;                      // RP_SLABS, or 1 for a short partial list (rp_slabs below)
    const int per = (P + nslab - 1) / nslab;
    const int p0 = slab * per, p1 = min(P, p0 + per);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double s = 0.0;
        if (col < width)
            for (int p = p0 + lane; p < p1; p += 32) s += (double)__ldg(part[q] + (int64_t)p * width + col);
        sm[lane * 33 + cx] = s;
        __syncthreads();
        if (lane == 0) {
            double t = 0.0;
            for (int l = 0; l < 32; ++l) t += sm[l * 33 + cx];
            if (col < width) scratch[((int64_t)slab * NQ + q) * width + col] = t;
        }
        __syncthreads();
    }
    __threadfence();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(tickets + blockIdx.x, 1u);
        is_last = (t == (unsigned)nslab - 1u);
        if (is_last) tickets[blockIdx.x] = 0;          // re-arm for the next launch
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    if (lane != 0 || col >= width) return false;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double t = 0.0;
        for (int sl = 0; sl < nslab; ++sl) t += __ldcg(scratch + ((int64_t)sl * NQ + q) * width + col);
        out[q] = t;
    }
    return true;
}

// Bound on |gamma*xh + beta| of one BatchNorm channel, max-reduced over the channels of the stage into *abound (bit
// pattern of a non-negative float, zeroed per forward call): |xh| <= 16 is assumed and plane_scale() leaves another
// factor 2^8 of headroom, i.e. |xh| up to 4096 > sqrt(rows) for any batch this library is given.  The BN-apply kernels
// turn it into the power-of-two scale of the stage's fp16 planes.
__device__ __forceinline__ void bn_output_bound(unsigned int* abound, float gamma, float beta) {
    if (!abound) return;
    const float b = fmaf(fabsf(gamma), 16.f, fabsf(beta));
    if (b > 0.f && b < 3.0e38f) atomicMax(abound, __float_as_uint(b));
}

// BatchNorm statistics -> mean, inv-std, and the affine (scale, shift) the apply kernels use:
// y_bn = x*scale + shift with scale = gamma*istd, shift = beta - mean*scale (same arrangement as
// torch's CPU batch-norm transform).  mode: CP_BN_BATCH / _BATCH_UPDATE / _RUNNING.
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int P, int F,
                   int64_t R, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ run_mean, float* __restrict__ run_var, int mode, float momentum,
                   float eps, float* __restrict__ mean_o, float* __restrict__ istd_o,
                   float* __restrict__ scale_o, float* __restrict__ shift_o, double* __restrict__ scratch,
                   unsigned int* __restrict__ tickets, double* __restrict__ totals = nullptr,
                   unsigned int* __restrict__ abound = nullptr) {
    __shared__ double sm[32 * 33];
    const int col = blockIdx.x * 32 + threadIdx.x % 32, lane = threadIdx.x / 32;
    double mean, var;
    if (totals) {
        // SyncBN: only the rank-local column totals [sum | sum of squares | row count]; the caller all-reduces
        // them across ranks and bn_finalize_totals_kernel finishes the statistics
        const float* const parts[2] = {psum, psq};
        double tot[2];
        if (!reduce_partials_2level<2>(parts, P, F, scratch, tickets, tot, sm)) return;
        totals[col] = tot[0];
        totals[F + col] = tot[1];
        if (col == 0) totals[2 * F] = (double)R;
        return;
    }
    if (mode == CP_BN_RUNNING) {
        if (lane != 0 || col >= F || blockIdx.y != 0) return;
        mean = (double)run_mean[col];
        var = (double)run_var[col];
    } else {
        const float* const parts[2] = {psum, psq};
        double tot[2];
        if (!reduce_partials_2level<2>(parts, P, F, scratch, tickets, tot, sm)) return;
        const double s = tot[0], q = tot[1];
        mean = s / (double)R;
        var = q / (double)R - mean * mean;          // biased variance (normalisation)
        if (var < 0.0) var = 0.0;
        if (mode == CP_BN_BATCH_UPDATE) {
            const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
            run_mean[col] = (float)((1.0 - (double)momentum) * (double)run_mean[col] + (double)momentum * mean);
            run_var[col] = (float)((1.0 - (double)momentum) * (double)run_var[col] + (double)momentum * unbiased);
        }
    }
    const float istd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[col] * istd;
    mean_o[col] = (float)mean;
    istd_o[col] = istd;
    scale_o[col] = sc;
    shift_o[col] = beta[col] - (float)mean * sc;
    bn_output_bound(abound, gamma[col], beta[col]);
}

// SyncBN second half: statistics from the (all-reduced) totals [sum | sumsq | rows] of every rank
__global__ void __launch_bounds__(512)
bn_finalize_totals_kernel(const double* __restrict__ totals, int F, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float* __restrict__ run_mean, float* __restrict__ run_var,
                          int mode, float momentum, float eps, float* __restrict__ mean_o, float* __restrict__ istd_o,
                          float* __restrict__ scale_o, float* __restrict__ shift_o,
                          unsigned int* __restrict__ abound = nullptr) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= F) return;
    const double R = totals[2 * F];
    const double mean = totals[col] / R;
    double var = totals[F + col] / R - mean * mean;
    if (var < 0.0) var = 0.0;
    if (mode == CP_BN_BATCH_UPDATE) {
        const double unbiased = R > 1.0 ? var * R / (R - 1.0) : var;
        run_mean[col] = (float)((1.0 - (double)momentum) * (double)run_mean[col] + (double)momentum * mean);
        run_var[col] = (float)((1.0 - (double)momentum) * (double)run_var[col] + (double)momentum * unbiased);
    }
    const float istd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[col] * istd;
    mean_o[col] = (float)mean;
    istd_o[col] = istd;
    scale_o[col] = sc;
    shift_o[col] = beta[col] - (float)mean * sc;
    bn_output_bound(abound, gamma[col], beta[col]);
}

// SyncBN backward second half: m1 = sum g' / R, m2 = sum g'*xh / R over the rows of every rank
__global__ void __launch_bounds__(512)
bn_bwd_means_totals_kernel(const double* __restrict__ totals, int F, float* __restrict__ m1, float* __restrict__ m2) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= F) return;
    const double R = totals[2 * F];
    m1[col] = (float)(totals[col] / R);
    m2[col] = (float)(totals[F + col] / R);
}

// ---------------------------------------------------------------------------------- BN apply
// a = y*scale + shift, then (linear blocks 4..7) dropout: a * keep / (1-p)   (models.py:282-297)
// SPLIT: write the result as the two fp16 planes (hi -> a, lo -> a_lo; see gemm_tc.cuh) the tensor-core GEMMs consume.
// Dropout: `keep` holds a caller-provided mask, or (gen_p > 0) the mask is drawn here -- Philox4x32-10 (common.cuh)
// keyed by seed [+ *seed_offset * odd constant], counter = (element/4, layer) -- and stored to `keep` for the
// backward pass.
template <int F, bool SPLIT>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ y, float* __restrict__ a, float* __restrict__ a_lo, int64_t R,
                const float* __restrict__ scale, const float* __restrict__ shift,
                uint8_t* __restrict__ keep, float inv_keep, float gen_p, uint64_t seed, uint64_t layer,
                const unsigned long long* __restrict__ seed_offset = nullptr,
                const unsigned int* __restrict__ abound = nullptr, float* __restrict__ ascale_inv = nullptr) {
    const int64_t total = R * (F / 4);
    float S = 1.f;
    if (SPLIT) {        // planes of a * S: S from the BatchNorm output bound (times the dropout scale)
        S = plane_scale(__uint_as_float(__ldg(abound)) * inv_keep);
        if (blockIdx.x == 0 && threadIdx.x == 0) *ascale_inv = 1.f / S;
    }
    // device-resident step counter: lets a CUDA-graph replay of the same launch draw a fresh mask
    if (gen_p > 0.f && seed_offset) seed += __ldg(seed_offset) * 0x9E3779B97F4A7C15ull;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total;
         v += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(v % (F / 4)) * 4;
        const float4 x = __ldg(reinterpret_cast<const float4*>(y) + v);
        const float4 s = __ldg(reinterpret_cast<const float4*>(scale + c));
        const float4 t = __ldg(reinterpret_cast<const float4*>(shift + c));
        float4 o = make_float4(fmaf(x.x, s.x, t.x), fmaf(x.y, s.y, t.y), fmaf(x.z, s.z, t.z), fmaf(x.w, s.w, t.w));
        if (keep) {
            uchar4 m;
            if (gen_p > 0.f) {
                m = dropout_keep4(seed, (uint64_t)v, (unsigned int)layer, dropout_threshold(gen_p));
                reinterpret_cast<uchar4*>(keep)[v] = m;
            } else {
                m = *(reinterpret_cast<const uchar4*>(keep) + v);
            }
            o.x = m.x ? o.x * inv_keep : 0.f; o.y = m.y ? o.y * inv_keep : 0.f;
            o.z = m.z ? o.z * inv_keep : 0.f; o.w = m.w ? o.w * inv_keep : 0.f;
        }
        if (SPLIT) {
            split_store4(make_float4(o.x * S, o.y * S, o.z * S, o.w * S), reinterpret_cast<plane_t*>(a),
                         reinterpret_cast<plane_t*>(a_lo), v);
        } else {
            reinterpret_cast<float4*>(a)[v] = o;
        }
    }
}

// ------------------------------------------------------------------------------- BN backward
// g' = g * keep/(1-p);  xh = (y - mean)*istd;  partials: sum g', sum g'*xh
// post != null: the block is Linear -> BN -> ReLU (glove tower block 0, models.py:398-400): `post` is the
// block output relu(bn(y)), g' = g * 1[post > 0], and no ReLU mask follows the BN backward.
// Does the stage whose BN-backward sums were derived from the next layer's parameter gradients
// (bn_bwd_stats_from_wgrad_kernel) have to be recomputed the long way?  Yes when a gamma == 0 was met (*flag), or when
// the derivation cancelled badly: that kernel collects per-CTA pairs (E^2, D^2) with E_f = (|sum W dW| +
// |beta sum g'|) / |gamma| the magnitude of the terms and D_f = |sum g' xh| the result -- a weight-gradient error eps
// becomes eps * E / D in d_gamma (norm-wise over the stage); beyond 8x its last CTA raises the flag too.
#define WS_CANCEL_PARTS 64
__device__ __forceinline__ bool stage_needs_exact(const unsigned int* __restrict__ flag) { return __ldg(flag) != 0u; }

template <int F>
__device__ __forceinline__ void bn_bwd_reduce_body(const float* __restrict__ g, const float* __restrict__ y, int64_t R,
                                                   const uint8_t* __restrict__ keep, float inv_keep,
                                                   const float* __restrict__ mean, const float* __restrict__ istd,
                                                   float* __restrict__ p1, float* __restrict__ p2,
                                                   const float* __restrict__ post, unsigned int* __restrict__ gmax_bits,
                                                   float* red) {
    float gmax = 0.f;                     // max |g'| (feeds the fp16 plane scale of bn_bwd_apply_kernel<.., true>)
    const int qx = threadIdx.x % ColMap<F>::QX, ry = threadIdx.x / ColMap<F>::QX;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + qx * 4));
    const float4 is = __ldg(reinterpret_cast<const float4*>(istd + qx * 4));
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    const int64_t r0 = (int64_t)blockIdx.x * ColMap<F>::ROWS;
    for (int k = ry; k < ColMap<F>::ROWS; k += ColMap<F>::RY) {
        const int64_t r = r0 + k;
        if (r >= R) break;
        const int64_t v = r * (F / 4) + qx;
        float4 gv = __ldg(reinterpret_cast<const float4*>(g) + v);
        const float4 yv = __ldg(reinterpret_cast<const float4*>(y) + v);
        if (keep) {
            const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(keep) + v);
            gv.x = m.x ? gv.x * inv_keep : 0.f; gv.y = m.y ? gv.y * inv_keep : 0.f;
            gv.z = m.z ? gv.z * inv_keep : 0.f; gv.w = m.w ? gv.w * inv_keep : 0.f;
        }
        if (post) {
            const float4 pv = __ldg(reinterpret_cast<const float4*>(post) + v);
            gv.x = pv.x > 0.f ? gv.x : 0.f; gv.y = pv.y > 0.f ? gv.y : 0.f;
            gv.z = pv.z > 0.f ? gv.z : 0.f; gv.w = pv.w > 0.f ? gv.w : 0.f;
        }
        gmax = fmaxf(fmaxf(gmax, fmaxf(fabsf(gv.x), fabsf(gv.y))), fmaxf(fabsf(gv.z), fabsf(gv.w)));
        s1[0] += gv.x; s1[1] += gv.y; s1[2] += gv.z; s1[3] += gv.w;
        s2[0] = fmaf(gv.x, (yv.x - mu.x) * is.x, s2[0]);
        s2[1] = fmaf(gv.y, (yv.y - mu.y) * is.y, s2[1]);
        s2[2] = fmaf(gv.z, (yv.z - mu.z) * is.z, s2[2]);
        s2[3] = fmaf(gv.w, (yv.w - mu.w) * is.w, s2[3]);
    }
    block_col_reduce<F>(s1, red, qx, ry);
    block_col_reduce<F>(s2, red, qx, ry);
    if (ry == 0) {
        *reinterpret_cast<float4*>(p1 + (int64_t)blockIdx.x * F + qx * 4) = make_float4(s1[0], s1[1], s1[2], s1[3]);
        *reinterpret_cast<float4*>(p2 + (int64_t)blockIdx.x * F + qx * 4) = make_float4(s2[0], s2[1], s2[2], s2[3]);
    }
    if (gmax_bits) {
        gmax = warp_max(gmax);
        // non-negative floats order like their bit patterns; NaN / Inf gradients are the caller's problem
        if (threadIdx.x % 32 == 0 && gmax > 0.f) atomicMax(gmax_bits, __float_as_uint(gmax));
    }
}

template <int F>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t R,
                     const uint8_t* __restrict__ keep, float inv_keep, const float* __restrict__ mean,
                     const float* __restrict__ istd, float* __restrict__ p1, float* __restrict__ p2,
                     const float* __restrict__ post = nullptr, unsigned int* __restrict__ gmax_bits = nullptr,
                     const unsigned int* __restrict__ run_flag = nullptr) {
    __shared__ float red[ColMap<F>::RY * F];
    // fallback pass of the reduce-free BN backward: runs only when bn_bwd_stats_from_wgrad_kernel asked for it
    if (run_flag && !stage_needs_exact(run_flag)) return;
    bn_bwd_reduce_body<F>(g, y, R, keep, inv_keep, mean, istd, p1, p2, post, gmax_bits, red);
}

// The conditional (normally empty) fallback of the reduce-free BN backward as ONE launch: the reduce pass, then the last
// CTA to finish (ticket, re-armed) adds the P partial rows in index order, in double, and writes what
// bn_bwd_finalize_kernel writes.  One CTA walking every partial row is slow (~0.1 ms at 1,312 rows) -- it runs only when
// a stage's derivation from dW was rejected -- and it saves a graph node per stage in the common case.
template <int F>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_finalize_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t R,
                              const uint8_t* __restrict__ keep, float inv_keep, const float* __restrict__ mean,
                              const float* __restrict__ istd, float* __restrict__ p1, float* __restrict__ p2,
                              unsigned int* __restrict__ gmax_bits, const unsigned int* __restrict__ run_flag,
                              float* __restrict__ m1, float* __restrict__ m2, float* __restrict__ d_gamma,
                              float* __restrict__ d_beta, unsigned int* __restrict__ ticket) {
    __shared__ float red[ColMap<F>::RY * F];
    __shared__ bool last;
    if (!stage_needs_exact(run_flag)) return;
    bn_bwd_reduce_body<F>(g, y, R, keep, inv_keep, mean, istd, p1, p2, nullptr, gmax_bits, red);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        if (last) *ticket = 0;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int col = threadIdx.x; col < F; col += 256) {
        double a = 0.0, b = 0.0;
        for (unsigned int p = 0; p < gridDim.x; ++p) {
            a += (double)__ldcg(p1 + (int64_t)p * F + col);
            b += (double)__ldcg(p2 + (int64_t)p * F + col);
        }
        if (d_beta) d_beta[col] = (float)a;
        if (d_gamma) d_gamma[col] = (float)b;
        m1[col] = (float)(a / (double)R);
        m2[col] = (float)(b / (double)R);
    }
}

// d_gamma = sum g'*xh, d_beta = sum g';  m1 = d_beta/R, m2 = d_gamma/R
__global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int P, int F, int64_t R,
                       float* __restrict__ m1, float* __restrict__ m2, float* __restrict__ d_gamma,
                       float* __restrict__ d_beta, double* __restrict__ scratch, unsigned int* __restrict__ tickets,
                       double* __restrict__ totals = nullptr, const unsigned int* __restrict__ run_flag = nullptr) {
    __shared__ double sm[32 * 33];
    if (run_flag && !stage_needs_exact(run_flag)) return;          // see bn_bwd_reduce_kernel
    const int col = blockIdx.x * 32 + threadIdx.x % 32;
    const float* const parts[2] = {p1, p2};
    double tot[2];
    if (!reduce_partials_2level<2>(parts, P, F, scratch, tickets, tot, sm)) return;
    const double a = tot[0], b = tot[1];
    // d_gamma / d_beta stay rank-local (the parameter-gradient all-reduce averages them like every other
    // gradient); SyncBN only shares the two means that enter the data gradient
    if (d_beta) d_beta[col] = (float)a;
    if (d_gamma) d_gamma[col] = (float)b;
    if (totals) {
        totals[col] = a;
        totals[F + col] = b;
        if (col == 0) totals[2 * F] = (double)R;
        return;
    }
    m1[col] = (float)(a / (double)R);
    m2[col] = (float)(b / (double)R);
}

// gz = 1[y>0] * gamma*istd * (g' - m1 - xh*m2)     (BN backward, then ReLU backward);
// partial column sums of gz = bias gradient of the preceding Linear / Conv.
// SPLIT: gz is written as fp16 (hi, lo) planes of gz * S with S a power of two chosen from a bound on |gz|
// (max |gamma*istd| * max |g'| * 18: |m1| <= max|g'|, |m2| <= max|g'| E|xh| <= max|g'|, |xh| <= 16) so that
// S * bound = 2^8; 1/S goes to *gscale_inv for the consumers' epilogues (data gradient, weight gradient).
template <int F, bool SPLIT>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t R,
                    const uint8_t* __restrict__ keep, float inv_keep, const float* __restrict__ mean,
                    const float* __restrict__ istd, const float* __restrict__ gamma,
                    const float* __restrict__ m1, const float* __restrict__ m2, float* __restrict__ gz,
                    float* __restrict__ gz_lo, float* __restrict__ pdb, const float* __restrict__ post = nullptr,
                    const unsigned int* __restrict__ gmax_bits = nullptr, float* __restrict__ gscale_inv = nullptr,
                    unsigned int* __restrict__ g1max_out = nullptr /*max |gz| (bit pattern), zeroed per backward call*/,
                    int rows_per_cta = ColMap<F>::ROWS /*small batches: fewer rows per CTA, more CTAs (pdb rows = grid)*/) {
    __shared__ float red[ColMap<F>::RY * F];
    float zmax = 0.f;
    const int qx = threadIdx.x % ColMap<F>::QX, ry = threadIdx.x / ColMap<F>::QX;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + qx * 4));
    const float4 is = __ldg(reinterpret_cast<const float4*>(istd + qx * 4));
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + qx * 4));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(m1 + qx * 4));
    const float4 a2 = __ldg(reinterpret_cast<const float4*>(m2 + qx * 4));
    const float k0 = ga.x * is.x, k1 = ga.y * is.y, k2 = ga.z * is.z, k3 = ga.w * is.w;
    float S = 1.f;
    if (SPLIT) {
        __shared__ float kred[8];
        float km = warp_max(fmaxf(fmaxf(fabsf(k0), fabsf(k1)), fmaxf(fabsf(k2), fabsf(k3))));
        if (threadIdx.x % 32 == 0) kred[threadIdx.x / 32] = km;
        __syncthreads();
        km = kred[0];
#pragma unroll
        for (int w8 = 1; w8 < 8; ++w8) km = fmaxf(km, kred[w8]);
        S = plane_scale(km * __uint_as_float(__ldg(gmax_bits)) * 18.f);
        if (blockIdx.x == 0 && threadIdx.x == 0) *gscale_inv = 1.f / S;
    }
    float sb[4] = {0, 0, 0, 0};
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    for (int k = ry; k < rows_per_cta; k += ColMap<F>::RY) {
        const int64_t r = r0 + k;
        if (r >= R) break;
        const int64_t v = r * (F / 4) + qx;
        float4 gv = __ldg(reinterpret_cast<const float4*>(g) + v);
        const float4 yv = __ldg(reinterpret_cast<const float4*>(y) + v);
        if (keep) {
            const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(keep) + v);
            gv.x = m.x ? gv.x * inv_keep : 0.f; gv.y = m.y ? gv.y * inv_keep : 0.f;
            gv.z = m.z ? gv.z * inv_keep : 0.f; gv.w = m.w ? gv.w * inv_keep : 0.f;
        }
        const bool pre = post == nullptr;         // ReLU precedes the BN: mask the result with 1[y > 0]
        if (!pre) {
            const float4 pv = __ldg(reinterpret_cast<const float4*>(post) + v);
            gv.x = pv.x > 0.f ? gv.x : 0.f; gv.y = pv.y > 0.f ? gv.y : 0.f;
            gv.z = pv.z > 0.f ? gv.z : 0.f; gv.w = pv.w > 0.f ? gv.w : 0.f;
        }
        float4 o;
        o.x = (!pre || yv.x > 0.f) ? k0 * (gv.x - a1.x - (yv.x - mu.x) * is.x * a2.x) : 0.f;
        o.y = (!pre || yv.y > 0.f) ? k1 * (gv.y - a1.y - (yv.y - mu.y) * is.y * a2.y) : 0.f;
        o.z = (!pre || yv.z > 0.f) ? k2 * (gv.z - a1.z - (yv.z - mu.z) * is.z * a2.z) : 0.f;
        o.w = (!pre || yv.w > 0.f) ? k3 * (gv.w - a1.w - (yv.w - mu.w) * is.w * a2.w) : 0.f;
        if (SPLIT) {
            split_store4(make_float4(o.x * S, o.y * S, o.z * S, o.w * S), reinterpret_cast<plane_t*>(gz),
                         reinterpret_cast<plane_t*>(gz_lo), v);
        } else {
            reinterpret_cast<float4*>(gz)[v] = o;
        }
        sb[0] += o.x; sb[1] += o.y; sb[2] += o.z; sb[3] += o.w;
        zmax = fmaxf(fmaxf(zmax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
    }
    block_col_reduce<F>(sb, red, qx, ry);
    if (ry == 0)
        *reinterpret_cast<float4*>(pdb + (int64_t)blockIdx.x * F + qx * 4) = make_float4(sb[0], sb[1], sb[2], sb[3]);
    if (g1max_out) {
        zmax = warp_max(zmax);
        if (threadIdx.x % 32 == 0 && zmax > 0.f) atomicMax(g1max_out, __float_as_uint(zmax));
    }
}

// Coefficients of the BN + ReLU backward that the data-gradient GEMM of the layer above applies in its epilogue
// (tcg::EPI_BNBWD):  gz = 1[y > 0] (c1 g + c2 y + c3)  ==  1[y > 0] gamma istd (g - m1 - xh m2),
//   c1 = gamma istd,  c2 = -c1 istd m2,  c3 = -c1 m1 - c2 mean;
// and a bound on |gz| for the power-of-two scale of its fp16 planes: max|c1| * 18 * max|g| (|m1| <= max|g|, |m2| <=
// max|g|, |xh| <= 16, as in bn_bwd_apply_kernel) with max|g| <= max|G1 above| * max column L1 norm of the layer's W.
// One CTA of F threads.
template <int F>
__global__ void __launch_bounds__(F)
bn_bwd_coef_kernel(const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ istd,
                   const float* __restrict__ m1, const float* __restrict__ m2,
                   const unsigned int* __restrict__ g1max_above, const unsigned int* __restrict__ l1max,
                   float* __restrict__ c1, float* __restrict__ c2, float* __restrict__ c3, float* __restrict__ gz_bound) {
    __shared__ float red[F / 32];
    const int t = threadIdx.x;
    const float k = gamma[t] * istd[t];
    const float k2 = -k * istd[t] * m2[t];
    c1[t] = k;
    c2[t] = k2;
    c3[t] = fmaf(-k2, mean[t], -k * m1[t]);
    const float km = warp_max(fabsf(k));
    if (t % 32 == 0) red[t / 32] = km;
    __syncthreads();
    if (t == 0) {
        float m = red[0];
        for (int i = 1; i < F / 32; ++i) m = fmaxf(m, red[i]);
        *gz_bound = m * 18.f * (__uint_as_float(__ldg(g1max_above)) * __uint_as_float(__ldg(l1max)));
    }
}

// out[ch] = sum_p sum_pos partial[p][pos*64 + ch]: the conv2 bias gradient from the [tiles][768] column sums of the
// position-major pre-activation gradient (12 positions x 64 channels per window)
__global__ void __launch_bounds__(1024)
colsum_fold12_kernel(const float* __restrict__ part, int P, float* __restrict__ out) {
    __shared__ double sm[32 * 33];
    const int ch = blockIdx.x * 32 + threadIdx.x % 32, lane = threadIdx.x / 32;
    double s = 0.0;
    for (int p = lane; p < P; p += 32)
        for (int pos = 0; pos < 12; ++pos) s += (double)__ldg(part + (int64_t)p * 768 + pos * 64 + ch);
    const int cx = threadIdx.x % 32;
    sm[lane * 33 + cx] = s;
    __syncthreads();
    if (lane == 0) {
        double t = 0.0;
        for (int l = 0; l < 32; ++l) t += sm[l * 33 + cx];
        out[ch] = (float)t;
    }
}

// ------------------------------------------------- BN-backward sums without a pass over the activations
// For a BN stage whose output A = gamma*xh + beta feeds the next Linear layer UNMASKED (no dropout), the gradient
// g = dA = G1 . W of that layer is linear in G1, so the two sums the BN backward needs over all R rows are already
// contained in the layer's own parameter gradients db = colsum(G1), dW = G1^T . A:
//     sum_r g[r,c]          = sum_k db[k] W[k,c]
//     sum_r g[r,c] xh[r,c]  = sum_k W[k,c] (G1^T xh)[k,c] = (sum_k W[k,c] dW[k,c] - beta[c] sum_k db[k] W[k,c]) / gamma[c]
// -- a [512 x F] reduction instead of two reads of [R x F] (bn_bwd_reduce_kernel).  GROUP = 12 folds the 12 positions of
// a conv-stage channel (flatten column ch*12 + p, models.py:263) into its per-channel BN2d sums.  gamma == 0 makes the
// second sum unobservable from dW: the kernel then raises *zero_gamma_flag and the caller's conditional reduce pass
// (bn_bwd_reduce_kernel / bn_bwd_finalize_kernel with run_flag) recomputes the stage's sums the long way.
// Double accumulation in a fixed order (deterministic).  Outputs like bn_bwd_finalize_kernel.
// With dropout between the stage and the layer only the first identity is lost (see sum_g_in below).
template <int GROUP> struct WgradStats {
    // linear stages: 8 columns per CTA = 64 CTAs, 64 row lanes -> 8 rows per thread, every load of a thread in flight at
    // once (16 columns / 32 CTAs / 16 dependent trips took 9 us at any batch size: 7 such launches are 10 % of the
    // batch_size-8 step)
    static constexpr int COLS = GROUP == 1 ? 8 : 24;                   // columns of W per CTA (24 = 2 channels x 12 positions)
    static constexpr int CX = GROUP == 1 ? 8 : 32;                     // column slots of the 512 threads ...
    static constexpr int LANES = 512 / CX;                             // ... and row lanes (every thread works in the linear stages)
};
template <int GROUP>
__global__ void __launch_bounds__(512)
bn_bwd_stats_from_wgrad_kernel(const float* __restrict__ W, const float* __restrict__ dW, const float* __restrict__ db,
                               int K_out /*rows of W*/, int cols /*columns of W = F * GROUP*/, int64_t R /*rows per column*/,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               float* m1, float* __restrict__ m2, float* __restrict__ d_gamma,
                               float* __restrict__ d_beta, const float* sum_g_in = nullptr,
                               unsigned int* __restrict__ zero_gamma_flag = nullptr,
                               double* __restrict__ cancel = nullptr /*[gridDim.x][2]: see stage_needs_exact*/,
                               unsigned int* __restrict__ cancel_ticket = nullptr /*zero on entry, re-armed*/) {
    constexpr int COLS = WgradStats<GROUP>::COLS;
    constexpr int WS_LANES = WgradStats<GROUP>::LANES;
    __shared__ double s_a[WS_LANES][COLS], s_t[WS_LANES][COLS];
    __shared__ double s_E[COLS], s_D[COLS];
    const int cx = threadIdx.x % WgradStats<GROUP>::CX, ky = threadIdx.x / WgradStats<GROUP>::CX;
    const int col = blockIdx.x * COLS + cx;
    double a = 0.0, t = 0.0;
    if (cx < COLS && col < cols) {
#pragma unroll 8
        for (int k = ky; k < K_out; k += WS_LANES) {
            const double w = (double)__ldg(W + (size_t)k * cols + col);
            a += (double)__ldg(db + k) * w;
            t += w * (double)__ldg(dW + (size_t)k * cols + col);
        }
    }
    if (cx < COLS) { s_a[ky][cx] = a; s_t[ky][cx] = t; }
    __syncthreads();
    constexpr int NFEAT = COLS / GROUP;                   // BN features of this CTA: one thread each
    if (threadIdx.x < NFEAT) {
        const int f = blockIdx.x * NFEAT + threadIdx.x;
        if (f * GROUP < cols) {
            double sa = 0.0, stt = 0.0;
            for (int p = 0; p < GROUP; ++p)
                for (int y = 0; y < WS_LANES; ++y) {
                    sa += s_a[y][threadIdx.x * GROUP + p];
                    stt += s_t[y][threadIdx.x * GROUP + p];
                }
            const double ga = (double)__ldg(gamma + f), be = (double)__ldg(beta + f);
            // below a dropout mask (sum_g_in, GROUP == 1; may alias m1): g' = g * keep/(1-p) is no longer linear in G1,
            // its column sums come from the data-gradient GEMM's epilogue; with A = (gamma xh + beta) keep/(1-p) the
            // second identity still holds:  sum g' xh = (sum_k W dW - beta sum g') / gamma
            const double sum_g = sum_g_in ? (double)sum_g_in[f] : sa;
            const double sum_gx = ga != 0.0 ? (stt - be * sum_g) / ga : 0.0;
            // gamma == 0: d_gamma is not observable from dW -> the caller's (otherwise skipped) reduce pass runs
            if (ga == 0.0 && zero_gamma_flag) atomicOr(zero_gamma_flag, 1u);
            // magnitude of the cancelling terms vs the result (stage_needs_exact)
            const double Ef = ga != 0.0 ? (fabs(stt) + fabs(be * sum_g)) / fabs(ga) : 0.0;
            s_E[threadIdx.x] = fmin(Ef * Ef, 1e300);
            s_D[threadIdx.x] = fmin(sum_gx * sum_gx, 1e300);
            const double rows = (double)R * GROUP;
            m1[f] = (float)(sum_g / rows);
            m2[f] = (float)(sum_gx / rows);
            if (d_beta) d_beta[f] = (float)sum_g;
            if (d_gamma) d_gamma[f] = (float)sum_gx;
        } else {
            s_E[threadIdx.x] = 0.0;
            s_D[threadIdx.x] = 0.0;
        }
    }
    if (cancel) {
        __syncthreads();
        if (threadIdx.x == 0) {
            double E = 0.0, D = 0.0;
            for (int i = 0; i < NFEAT; ++i) { E += s_E[i]; D += s_D[i]; }
            cancel[2 * blockIdx.x] = E;
            cancel[2 * blockIdx.x + 1] = D;
            // the last CTA to finish adds the per-CTA pairs in a fixed order (deterministic whichever CTA that is) and
            // raises the stage's flag when the derivation cancelled by more than 8x norm-wise (stage_needs_exact)
            __threadfence();
            const unsigned int t = atomicAdd(cancel_ticket, 1u);
            if (t == gridDim.x - 1) {
                *cancel_ticket = 0;                   // re-arm for the next launch
                __threadfence();
                E = D = 0.0;
                for (unsigned int i = 0; i < gridDim.x; ++i) {
                    E += __ldcg(cancel + 2 * i);
                    D += __ldcg(cancel + 2 * i + 1);
                }
                if (E > 64.0 * D && zero_gamma_flag) atomicOr(zero_gamma_flag, 1u);
            }
        }
    }
}

// out[c] = sum_p partial[p][c]   (bias gradients, projection / conv1 weight gradients)
// remap: 0 identity; 1 conv1 weight: partial col = tap*64 + c -> out[c*9 + 3 + tap]
__global__ void __launch_bounds__(1024)
colsum_finalize_kernel(const float* __restrict__ part, int P, int width, float* __restrict__ out, int remap) {
    __shared__ double sm[32 * 33];
    const int col = blockIdx.x * 32 + threadIdx.x % 32, lane = threadIdx.x / 32;
    const double s = reduce_partials(part, P, width, col, lane, sm);
    if (lane != 0 || col >= width) return;
    if (remap == 1) out[(col % 64) * 9 + 3 + col / 64] = (float)s;
    else out[col] = (float)s;
}

// partial[blockIdx.y][c] = sum of G[r, c] over the 128-row slab blockIdx.y (layer-level bias grad)
__global__ void __launch_bounds__(256)
colsum_rows_kernel(const float* __restrict__ G, int64_t M, int N, float* __restrict__ partial) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x % 32, lane = threadIdx.x / 32;
    const int col = blockIdx.x * 32 + cx;
    const int64_t r0 = (int64_t)blockIdx.y * 128;
    float s = 0.f;
    if (col < N)
        for (int k = lane; k < 128 && r0 + k < M; k += 8) s += __ldg(G + (r0 + k) * N + col);
    red[lane][cx] = s;
    __syncthreads();
    if (lane == 0 && col < N) {
        float t = 0.f;
        for (int l = 0; l < 8; ++l) t += red[l][cx];
        partial[(int64_t)blockIdx.y * N + col] = t;
    }
}

// --------------------------------------------------------------------------------- projection
// emb[r, o] = sum_k a[r,k] * Wp[o,k]    (models.py:314, bias-free 512 -> 16)
// One warp per 4 rows: each lane owns 16 k-values of every row, the weight slice is read from shared
// memory once per 4 rows, and the 16 outputs are reduced across the warp with a halving butterfly.
#define PROJ_RPW 4
template <int K>
__global__ void __launch_bounds__(256)
proj_fwd_kernel(const float* __restrict__ a, const float* __restrict__ Wp, float* __restrict__ emb, int64_t R) {
    constexpr int JN = K / 128;                     // float4 chunks per lane
    __shared__ __align__(16) float W[CP_EMB_DIM][K];
    for (int e = threadIdx.x; e < CP_EMB_DIM * K / 4; e += 256)
        reinterpret_cast<float4*>(&W[0][0])[e] = __ldg(reinterpret_cast<const float4*>(Wp) + e);
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    for (int64_t r0 = ((int64_t)blockIdx.x * 8 + warp) * PROJ_RPW; r0 < R; r0 += (int64_t)gridDim.x * 8 * PROJ_RPW) {
        float4 x[PROJ_RPW][JN];
#pragma unroll
        for (int i = 0; i < PROJ_RPW; ++i)
#pragma unroll
            for (int j = 0; j < JN; ++j)
                x[i][j] = (r0 + i < R) ? __ldg(reinterpret_cast<const float4*>(a + (r0 + i) * K) + lane + 32 * j)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        float acc[PROJ_RPW][CP_EMB_DIM];
#pragma unroll
        for (int o = 0; o < CP_EMB_DIM; ++o) {
            float4 w[JN];
#pragma unroll
            for (int j = 0; j < JN; ++j) w[j] = *reinterpret_cast<const float4*>(&W[o][(lane + 32 * j) * 4]);
#pragma unroll
            for (int i = 0; i < PROJ_RPW; ++i) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < JN; ++j) {
                    s = fmaf(x[i][j].x, w[j].x, s); s = fmaf(x[i][j].y, w[j].y, s);
                    s = fmaf(x[i][j].z, w[j].z, s); s = fmaf(x[i][j].w, w[j].w, s);
                }
                acc[i][o] = s;
            }
        }
#pragma unroll
        for (int i = 0; i < PROJ_RPW; ++i) {
            // 16 values per lane -> lane l (l < 16) ends with the warp total of output l
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int k = 0; k < off; ++k) {
                    const float send = up ? acc[i][k] : acc[i][k + off];
                    const float keep = up ? acc[i][k + off] : acc[i][k];
                    acc[i][k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            const float tot = acc[i][0] + __shfl_xor_sync(0xffffffffu, acc[i][0], 16);
            if (lane < CP_EMB_DIM && r0 + i < R) emb[(r0 + i) * CP_EMB_DIM + lane] = tot;
        }
    }
}

// ga[r,k] = sum_o d[r,o] * Wp[o,k]
template <int K>
__global__ void __launch_bounds__(256)
proj_bwd_data_kernel(const float* __restrict__ d, const float* __restrict__ Wp, float* __restrict__ ga, int64_t R) {
    constexpr int KQ = K / 4, RL = 256 / KQ;        // float4 columns, row lanes per CTA
    const int q = threadIdx.x % KQ, rl = threadIdx.x / KQ;
    float4 w[CP_EMB_DIM];
#pragma unroll
    for (int o = 0; o < CP_EMB_DIM; ++o) w[o] = __ldg(reinterpret_cast<const float4*>(Wp + o * K) + q);
    for (int64_t r0 = ((int64_t)blockIdx.x * RL + rl) * 4; r0 < R; r0 += (int64_t)gridDim.x * RL * 4) {
        float4 dv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                dv[i][j] = (r0 + i < R) ? __ldg(reinterpret_cast<const float4*>(d + (r0 + i) * CP_EMB_DIM) + j)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (r0 + i >= R) break;
            const float dd[CP_EMB_DIM] = {dv[i][0].x, dv[i][0].y, dv[i][0].z, dv[i][0].w, dv[i][1].x, dv[i][1].y,
                                          dv[i][1].z, dv[i][1].w, dv[i][2].x, dv[i][2].y, dv[i][2].z, dv[i][2].w,
                                          dv[i][3].x, dv[i][3].y, dv[i][3].z, dv[i][3].w};
            float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int o = 0; o < CP_EMB_DIM; ++o) {
                const float2 d2 = make_float2(dd[o], dd[o]);
                ffma2(reinterpret_cast<float2*>(&o4)[0], d2, make_float2(w[o].x, w[o].y));
                ffma2(reinterpret_cast<float2*>(&o4)[1], d2, make_float2(w[o].z, w[o].w));
            }
            reinterpret_cast<float4*>(ga + (r0 + i) * K)[q] = o4;
        }
    }
}

// dWp[o,k] partial over a slab of rows: partial[blk][o*512 + k]
#define PROJ_W_ROWS 256
template <int K>
__global__ void __launch_bounds__(256)
proj_bwd_weight_kernel(const float* __restrict__ d, const float* __restrict__ a, int64_t R,
                       float* __restrict__ partial) {
    constexpr int KQ = K / 4, RL = 256 / KQ;        // float4 columns, row lanes per CTA
    __shared__ float4 red[RL][KQ];
    const int q = threadIdx.x % KQ, rl = threadIdx.x / KQ;
    float4 acc[CP_EMB_DIM];
#pragma unroll
    for (int o = 0; o < CP_EMB_DIM; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t r0 = (int64_t)blockIdx.x * PROJ_W_ROWS;
    for (int k = rl * 4; k < PROJ_W_ROWS; k += RL * 4) {     // 4 rows in flight per thread
        float4 x[4], dv[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t r = r0 + k + i;
            const bool ok = r < R;
            x[i] = ok ? __ldg(reinterpret_cast<const float4*>(a + r * K) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                dv[i][j] = ok ? __ldg(reinterpret_cast<const float4*>(d + r * CP_EMB_DIM) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float dd[CP_EMB_DIM] = {dv[i][0].x, dv[i][0].y, dv[i][0].z, dv[i][0].w, dv[i][1].x, dv[i][1].y,
                                          dv[i][1].z, dv[i][1].w, dv[i][2].x, dv[i][2].y, dv[i][2].z, dv[i][2].w,
                                          dv[i][3].x, dv[i][3].y, dv[i][3].z, dv[i][3].w};
            const float2 xlo = make_float2(x[i].x, x[i].y), xhi = make_float2(x[i].z, x[i].w);
#pragma unroll
            for (int o = 0; o < CP_EMB_DIM; ++o) {
                const float2 d2 = make_float2(dd[o], dd[o]);
                ffma2(reinterpret_cast<float2*>(&acc[o])[0], d2, xlo);
                ffma2(reinterpret_cast<float2*>(&acc[o])[1], d2, xhi);
            }
        }
    }
    float* out = partial + (int64_t)blockIdx.x * CP_EMB_DIM * K;
#pragma unroll
    for (int o = 0; o < CP_EMB_DIM; ++o) {
        __syncthreads();
        red[rl][q] = acc[o];
        __syncthreads();
        if (rl == 0) {
            float4 t = acc[o];
#pragma unroll
            for (int l = 1; l < RL; ++l) {
                const float4 u = red[l][q];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            reinterpret_cast<float4*>(out + o * K)[q] = t;
        }
    }
}

// ------------------------------------------------------------------------------ weight layouts
// Wc2 [o][tap*64+c]  = conv2_w[o][c][1][tap]         conv2 forward  (B operand, [N=64,K=192])
// Wc2d[c][tap*64+o]  = conv2_w[o][c][1][2-tap]       conv2 data-gradient
// W1p [o][p*64+c]    = fc1_w[o][c*12+p]              fc1 on the position-major flatten
// Wc2_lo / Wc2d_lo non-null: write the fp16 (hi, lo) planes (tensor-core engine) instead of fp32
// c1w / c1b: workspace copy of the conv1 parameters (the parity tap recomputes the unsaved conv1 activation)
// max |W| of the 7 linear weights (blockIdx.y = 0..6) and of conv2's middle kernel row (7) -> wmax[8] (bit patterns,
// zeroed per forward call): the power-of-two scales of the weight planes
struct WmaxArgs { const float* W[8]; int n[8]; };
__device__ __forceinline__ void weights_absmax_body(const WmaxArgs& a, unsigned int* __restrict__ wmax, int bx, int l,
                                                    int gx) {
    float m = 0.f;
    if (l != 7 && a.n[l] % 4 == 0 && (reinterpret_cast<uintptr_t>(a.W[l]) & 15) == 0) {
        // linear weights: 16-byte loads (max is order-free: same result as the scalar walk)
        const float4* __restrict__ w4 = reinterpret_cast<const float4*>(a.W[l]);
        for (int i = bx * blockDim.x + threadIdx.x; i < a.n[l] / 4; i += gx * blockDim.x) {
            const float4 v = __ldg(w4 + i);
            m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
    } else {
        for (int i = bx * blockDim.x + threadIdx.x; i < a.n[l]; i += gx * blockDim.x) {
            // conv2: only the middle row of the 3 x 3 kernels ever meets data
            if (l == 7 && (i % 9) / 3 != 1) continue;
            m = fmaxf(m, fabsf(__ldg(a.W[l] + i)));
        }
    }
    m = warp_max(m);
    if (threadIdx.x % 32 == 0 && m > 0.f && m < 3.0e38f) atomicMax(wmax + l, __float_as_uint(m));
}
__global__ void __launch_bounds__(256)
weights_absmax_kernel(const WmaxArgs a, unsigned int* __restrict__ wmax) {
    weights_absmax_body(a, wmax, blockIdx.x, blockIdx.y, gridDim.x);
}
// max over the columns k of sum_n |W_l[n, k]| for the 7 linear weights (blockIdx.y) -> l1max[7] (bit patterns, zeroed per
// forward call): |G1 . W| <= max|G1| * this, the bound the fused BN-backward epilogue scales its planes by
__device__ __forceinline__ void weights_col_l1_body(const WmaxArgs& a, unsigned int* __restrict__ l1max, int bx, int l) {
    const int K = a.n[l] / 512;
    const int k = bx * blockDim.x + threadIdx.x;
    float s = 0.f;
    if (k < K) {
#pragma unroll 8
        for (int n = 0; n < 512; ++n) s += fabsf(__ldg(a.W[l] + (size_t)n * K + k));
    }
    s = warp_max(s);
    if (threadIdx.x % 32 == 0 && s > 0.f && s < 3.0e38f) atomicMax(l1max + l, __float_as_uint(s));
}
// max over the rows j of sum_k |W_l[j, k]| (-> rowl1[l]) and max_j |b_l[j]| (-> bmax[l]) for the 7 linear layers
// (blockIdx.y; one warp per row): |A . W^T + b| <= max|A| * rowl1 + bmax, the a-priori bound the GEMM epilogue scales
// the fp16 planes of its OUTPUT by when the following BatchNorm is folded into the next layer (fold_bn_weights_kernel)
struct RowL1Args { const float* W[7]; const float* b[7]; int K[7]; };
__device__ __forceinline__ void weights_row_l1_body(const RowL1Args& a, unsigned int* __restrict__ rowl1,
                                                    unsigned int* __restrict__ bmax, int bx, int l) {
    const int j = bx * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
    if (j >= 512) return;
    const int K = a.K[l];
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += fabsf(__ldg(a.W[l] + (size_t)j * K + k));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
        if (s > 0.f && s < 3.0e38f) atomicMax(rowl1 + l, __float_as_uint(s));
        const float b = fabsf(__ldg(a.b[l] + j));
        if (b > 0.f && b < 3.0e38f) atomicMax(bmax + l, __float_as_uint(b));
    }
}
// The three bound passes over the weights are independent max-reductions (atomicMax: order-free, deterministic): ONE
// launch, the block index selects the pass -- [0, n_row) row-L1 (64 x 7), [.., + n_abs) abs-max (48 x 8),
// [.., + n_col) column-L1 (3 x 7); a pass that the configuration does not need has n_* = 0.
struct WBoundsArgs {
    RowL1Args ra;
    WmaxArgs wa;
    unsigned int *rowl1, *bmax, *wmax, *l1max;
    int n_row, n_abs, n_col;
};
__global__ void __launch_bounds__(256)
weights_bounds_kernel(const WBoundsArgs a) {
    int b = blockIdx.x;
    if (b < a.n_row) { weights_row_l1_body(a.ra, a.rowl1, a.bmax, b % 64, b / 64); return; }
    b -= a.n_row;
    if (b < a.n_abs) { weights_absmax_body(a.wa, a.wmax, b % 48, b / 48, 48); return; }
    b -= a.n_abs;
    weights_col_l1_body(a.wa, a.l1max, b % 3, b / 3);
}

// BatchNorm folded into the NEXT linear layer (no BN-apply pass between two linear blocks without dropout):
//   (y * scale + shift) . W^T + b  =  y . (W diag(scale))^T + (b + W . shift)
// One CTA per output row j: planes of W[j,k] * scale[k] * S (S = power of two from max|W| * max|scale|), folded bias
// b'[j] = b[j] + sum_k W[j,k] * shift[k], *wscale_inv_out = 1 / S.  K = 512 (linear layers 2..4 of the encoder).
__global__ void __launch_bounds__(128)
fold_bn_weights_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ scale,
                       const float* __restrict__ shift, const unsigned int* __restrict__ wmax,
                       plane_t* __restrict__ Wh, plane_t* __restrict__ Wl, float* __restrict__ bias_out,
                       float* __restrict__ wscale_inv_out) {
    constexpr int K = 512;
    __shared__ float red[4], redm[4];
    const int j = blockIdx.x, t = threadIdx.x, lane = t % 32, wp = t / 32;
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + t);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + t);
    const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)j * K) + t);
    float m = fmaxf(fmaxf(fabsf(sc.x), fabsf(sc.y)), fmaxf(fabsf(sc.z), fabsf(sc.w)));
    float d = fmaf(w.x, sh.x, fmaf(w.y, sh.y, fmaf(w.z, sh.z, w.w * sh.w)));
    m = warp_max(m);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    if (lane == 0) { red[wp] = d; redm[wp] = m; }
    __syncthreads();
    d = (red[0] + red[1]) + (red[2] + red[3]);
    m = fmaxf(fmaxf(redm[0], redm[1]), fmaxf(redm[2], redm[3]));
    const float bound = __uint_as_float(__ldg(wmax)) * m;
    const float S = (bound > 0.f && bound < 3.0e38f) ? exp2f(-ceilf(log2f(bound))) : 1.f;
    if (t == 0) {
        bias_out[j] = __ldg(b + j) + d;
        if (j == 0) *wscale_inv_out = 1.f / S;
    }
    split_store4(make_float4(w.x * sc.x * S, w.y * sc.y * S, w.z * sc.z * S, w.w * sc.w * S), Wh + (size_t)j * K,
                 Wl + (size_t)j * K, t);
}

// S with max|W| * S in (1/2, 1]
__device__ __forceinline__ float weight_scale(const unsigned int* wmax_slot) {
    const float m = __uint_as_float(__ldg(wmax_slot));
    return m > 0.f ? exp2f(-ceilf(log2f(m))) : 1.f;
}

__global__ void __launch_bounds__(256)
prep_weights_kernel(const float* __restrict__ conv2_w, const float* __restrict__ fc1_w,
                    float* __restrict__ Wc2, float* __restrict__ Wc2d, float* __restrict__ W1p,
                    float* __restrict__ Wc2_lo, float* __restrict__ Wc2d_lo,
                    const float* __restrict__ conv1_w, const float* __restrict__ conv1_b,
                    float* __restrict__ c1w, float* __restrict__ c1b,
                    const unsigned int* __restrict__ wmax = nullptr, float* __restrict__ wscale_inv = nullptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float S2 = Wc2_lo ? weight_scale(wmax + 7) : 1.f;
    if (Wc2_lo && i == 0) wscale_inv[7] = 1.f / S2;
    if (i < 64 * 9) c1w[i] = __ldg(conv1_w + i);
    if (i < 64) c1b[i] = __ldg(conv1_b + i);
    if (i < 64 * 192) {
        const int o = i / 192, k = i % 192, tap = k / 64, c = k % 64;
        const float a = __ldg(conv2_w + (o * 64 + c) * 9 + 3 + tap);
        // same flat index read as [c'][tap*64 + o'] with c' = o, o' = c
        const float b = __ldg(conv2_w + (c * 64 + o) * 9 + 3 + (2 - tap));
        if (Wc2_lo) {
            split_f16(a * S2, reinterpret_cast<plane_t*>(Wc2)[i], reinterpret_cast<plane_t*>(Wc2_lo)[i]);
            split_f16(b * S2, reinterpret_cast<plane_t*>(Wc2d)[i], reinterpret_cast<plane_t*>(Wc2d_lo)[i]);
        } else {
            Wc2[i] = a;
            Wc2d[i] = b;
        }
    }
    if (i < 512 * 768) {
        const int o = i / 768, k = i % 768, p = k / 64, c = k % 64;
        W1p[i] = __ldg(fc1_w + o * 768 + c * 12 + p);
    }
}

// Tensor-core engine weight planes, per linear layer l (K = 768 for l = 0 on the permuted flatten):
//   fwd  Wh/Wl [512][K]   = split(W (l = 0: W1p))          B operand of Y = A . W^T      (K-major)
//   dgrad Wth/Wtl [K][512] = split(transpose)                B operand of dA = G . W       (K-major)
// One launch for all 7 layers: blockIdx.y = layer.
struct PrepTcArgs {
    const float* W[7];
    plane_t *Wh[7], *Wl[7], *Wth[7], *Wtl[7];
    const unsigned int* wmax;      // [8] from weights_absmax_kernel
    float* wscale_inv;             // [8] 1 / (power-of-two scale of the layer's weight planes)
};
__global__ void __launch_bounds__(256)
prep_weights_tc_kernel(const PrepTcArgs a) {
    // 32 x 32 tiles (grid: K/32, 512/32, layer), 256 threads = 32 x 8: straight planes written along k, transposed planes
    // along o through shared memory (the element-per-thread version scattered 2-byte writes 1 KB apart: 37 us for 2 M weights)
    __shared__ uint32_t tile[32][33];                 // hi | lo << 16
    const int l = blockIdx.z;
    const int K = l == 0 ? 768 : 512;
    const int k0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    if (k0 >= K) return;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    const float S = weight_scale(a.wmax + l);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.wscale_inv[l] = 1.f / S;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int o = o0 + r, k = k0 + tx;
        // fc1 runs on the position-major flatten: column p*64+c of the operand is column c*12+p of the parameter
        const float w = l == 0 ? __ldg(a.W[0] + o * 768 + (k % 64) * 12 + k / 64) : __ldg(a.W[l] + (size_t)o * K + k);
        plane_t h, lo;
        split_f16(w * S, h, lo);
        a.Wh[l][(size_t)o * K + k] = h;
        a.Wl[l][(size_t)o * K + k] = lo;
        tile[r][tx] = (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(lo) << 16);
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, o = o0 + tx;
        const uint32_t v = tile[tx][r];
        a.Wth[l][(size_t)k * 512 + o] = __ushort_as_half((unsigned short)(v & 0xffffu));
        a.Wtl[l][(size_t)k * 512 + o] = __ushort_as_half((unsigned short)(v >> 16));
    }
}

// dW = sum_z P[z]  with the inverse re-layouts.  mode 0: identity; 1: fc1 (cols p*64+c -> c*12+p);
// 2: conv2 (cols tap*64+c -> [o][c][1][tap] of a zero-initialised (64,64,3,3) tensor);
// 3: conv2 from the transposed tensor-core partials P[z][256][64] (row tap*64+c, column o)
// fold_scale / fold_shift / fold_db non-null (mode 0): the A operand was the PRE-BatchNorm activation y of a stage whose BN
// is folded into this layer (fold_bn_weights_kernel): dW[o,k] = sum_r g[r,o] (y[r,k] scale[k] + shift[k])
//                                                             = scale[k] * (G^T y)[o,k] + shift[k] * db[o]
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ P, int S, int Mo, int No, float* __restrict__ out, int mode,
                    const float* __restrict__ scale = nullptr, const float* __restrict__ scale2 = nullptr,
                    const float* __restrict__ fold_scale = nullptr, const float* __restrict__ fold_shift = nullptr,
                    const float* __restrict__ fold_db = nullptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Mo * No) return;
    // undoes the power-of-two scales of the G planes and of the activation planes
    const double sc = (scale ? (double)__ldg(scale) : 1.0) * (scale2 ? (double)__ldg(scale2) : 1.0);
    if (mode == 3) {
        const int m = i / 64, o = i % 64;                 // Mo = 192 rows used of 256, No = 64
        double s = 0.0;
        for (int z = 0; z < S; ++z) s += (double)__ldg(P + ((int64_t)z * 256 + m) * 64 + o);
        out[(o * 64 + m % 64) * 9 + 3 + m / 64] = (float)(s * sc);
        return;
    }
    double s = 0.0;
    for (int z = 0; z < S; ++z) s += (double)__ldg(P + (int64_t)z * Mo * No + i);
    s *= sc;
    const int o = i / No, k = i % No;
    if (fold_scale) s = s * (double)__ldg(fold_scale + k) + (double)__ldg(fold_shift + k) * (double)__ldg(fold_db + o);
    if (mode == 0) out[i] = (float)s;
    else if (mode == 1) out[o * 768 + (k % 64) * 12 + k / 64] = (float)s;
    else out[(o * 64 + k % 64) * 9 + 3 + k / 64] = (float)s;
}
