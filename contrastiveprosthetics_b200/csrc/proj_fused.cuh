// Last linear block + 512 -> 16 projection, fused (models.py:293-297, 310-315).
//
// The output A7 = dropout(BN(Y7)) of the last block feeds only the bias-free projection emb = A7 . Wp^T, and the
// gradient w.r.t. A7 is the rank-16 product d_emb . Wp.  Neither [n, 512] tensor is worth a trip through HBM:
//   forward   bn_apply_proj_kernel     Y7 (+ mask) -> emb                    (was: BN apply -> A7 -> projection)
//   backward  proj_bwd_reduce_kernel   Y7, mask, d_emb -> dWp partials + the two BN-backward sums + max|g'|
//             proj_bwd_apply_kernel    Y7, mask, d_emb -> pre-activation gradient planes + bias-gradient partials
//                                      (was: dWp kernel, d_emb . Wp -> G0, BN reduce, BN apply: G0 written once and
//                                       read twice, A7 read once)
// A7 and d_emb . Wp are recomputed per element from registers (16 FMAs against the thread's 4 columns of Wp).
//
// Row streaming: the kernels hold 64-128 accumulator / weight registers per thread, so memory-level
// parallelism cannot come from thread count.  Each persistent CTA instead pulls slabs of SR rows (Y7 rows are
// 2 KB contiguous, masks 512 B, d_emb 64 B) into a 4-stage shared-memory ring with cp.async.bulk (one elected
// thread, mbarrier completion): ~120 KB in flight per SM independent of the register budget.
// Thread layout: 128 threads per row (one float4 of columns each), 2 row lanes (4 in the forward kernel).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace pf {

constexpr int F = 512, QX = F / 4;
constexpr int SR = 16;                                   // rows per stage
constexpr int NST = 4;                                   // ring depth
constexpr int Y_BYTES = SR * F * 4, K_BYTES = SR * F, D_BYTES = SR * CP_EMB_DIM * 4;
constexpr int STAGE = Y_BYTES + K_BYTES + D_BYTES;       // 41,984 B
constexpr int BAR_BYTES = 128;
constexpr int SMEM = BAR_BYTES + NST * STAGE;            // 168,064 B (ring reused as reduction scratch at the end)

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tc::smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}

// Walks the slabs blockIdx.x, blockIdx.x + gridDim.x, ...; body(row0, rows, ys, ks, ds) sees one staged slab
// (ys [SR][F] floats, ks [SR][F] mask bytes, ds [SR][16] floats) and must not touch it after returning.
template <bool HAS_KEEP, bool HAS_D, typename Body>
__device__ __forceinline__ void stream_rows(const float* __restrict__ y, const uint8_t* __restrict__ keep,
                                            const float* __restrict__ d, int64_t n, uint8_t* smem, Body body) {
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint8_t* ring = smem + BAR_BYTES;
    const int64_t n_slabs = (n + SR - 1) / SR;
    const int64_t mine = blockIdx.x < n_slabs ? (n_slabs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) tc::mbar_init(&full[s], 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int64_t k) {
        const int s = (int)(k % NST);
        const int64_t row0 = (blockIdx.x + k * gridDim.x) * SR;
        const uint32_t rows = (uint32_t)min((int64_t)SR, n - row0);
        uint8_t* st = ring + s * STAGE;
        tc::mbar_expect_tx(&full[s], rows * (F * 4 + (HAS_KEEP ? F : 0) + (HAS_D ? CP_EMB_DIM * 4 : 0)));
        bulk_load(st, y + row0 * F, rows * F * 4, &full[s]);
        if (HAS_KEEP) bulk_load(st + Y_BYTES, keep + row0 * F, rows * F, &full[s]);
        if (HAS_D) bulk_load(st + Y_BYTES + K_BYTES, d + row0 * CP_EMB_DIM, rows * CP_EMB_DIM * 4, &full[s]);
    };
    if (threadIdx.x == 0)
        for (int64_t k = 0; k < NST - 1 && k < mine; ++k) issue(k);
    for (int64_t k = 0; k < mine; ++k) {
        // stage (k-1) % NST was released by the __syncthreads that ended iteration k-1
        if (threadIdx.x == 0 && k + NST - 1 < mine) issue(k + NST - 1);
        const int s = (int)(k % NST);
        tc::mbar_wait(&full[s], (uint32_t)((k / NST) & 1));
        const int64_t row0 = (blockIdx.x + k * gridDim.x) * SR;
        const uint8_t* st = ring + s * STAGE;
        body(row0, (int)min((int64_t)SR, n - row0), reinterpret_cast<const float4*>(st),
             reinterpret_cast<const uchar4*>(st + Y_BYTES), reinterpret_cast<const float4*>(st + Y_BYTES + K_BYTES));
        __syncthreads();
    }
}

__device__ __forceinline__ float4 keep_scale(const float4& v, const uchar4& m, float inv_keep) {
    return make_float4(m.x ? v.x * inv_keep : 0.f, m.y ? v.y * inv_keep : 0.f, m.z ? v.z * inv_keep : 0.f,
                       m.w ? v.w * inv_keep : 0.f);
}

// ------------------------------------------------------------------------------------------ forward
// MODE 0: no dropout; 1: caller-provided mask in `keep`; 2: mask drawn here (same Philox stream as
// bn_apply_kernel: key seed [+ step * odd constant], counter (element / 4, layer)) and stored to `keep`.
constexpr int FWD_THREADS = 512;                         // 4 row lanes: the Philox chain needs warps, not registers
template <int MODE>
__global__ void __launch_bounds__(FWD_THREADS, 1)
bn_apply_proj_kernel(const float* __restrict__ y, int64_t n, const float* __restrict__ scale,
                     const float* __restrict__ shift, uint8_t* __restrict__ keep, float inv_keep, float gen_p,
                     uint64_t seed, uint64_t layer, const unsigned long long* __restrict__ seed_offset,
                     const float* __restrict__ Wp, float* __restrict__ emb) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ float part[SR][4][CP_EMB_DIM];
    const int qx = threadIdx.x % QX, rl = threadIdx.x / QX;
    const int lane = threadIdx.x % 32, wq = (threadIdx.x / 32) % 4;
    float4 wp[CP_EMB_DIM];
#pragma unroll
    for (int o = 0; o < CP_EMB_DIM; ++o) wp[o] = __ldg(reinterpret_cast<const float4*>(Wp + o * F) + qx);
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + qx);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + qx);
    if (MODE == 2 && seed_offset) seed += __ldg(seed_offset) * 0x9E3779B97F4A7C15ull;
    const unsigned int thr = dropout_threshold(gen_p);

    stream_rows<MODE == 1, false>(y, keep, nullptr, n, smem, [&](int64_t row0, int rows, const float4* ys,
                                                                 const uchar4* ks, const float4*) {
        for (int r = rl; r < rows; r += FWD_THREADS / QX) {
            const float4 x = ys[r * QX + qx];
            float4 a = make_float4(fmaf(x.x, sc.x, sh.x), fmaf(x.y, sc.y, sh.y), fmaf(x.z, sc.z, sh.z), fmaf(x.w, sc.w, sh.w));
            if (MODE == 1) a = keep_scale(a, ks[r * QX + qx], inv_keep);
            if (MODE == 2) {
                const int64_t v = (row0 + r) * QX + qx;
                const uchar4 m = dropout_keep4(seed, (uint64_t)v, (unsigned int)layer, thr);
                reinterpret_cast<uchar4*>(keep)[v] = m;
                a = keep_scale(a, m, inv_keep);
            }
            float acc[CP_EMB_DIM];
#pragma unroll
            for (int o = 0; o < CP_EMB_DIM; ++o)
                acc[o] = fmaf(a.x, wp[o].x, fmaf(a.y, wp[o].y, fmaf(a.z, wp[o].z, a.w * wp[o].w)));
            // 16 values per lane -> lane l (l < 16) ends with the warp total of output l
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int k = 0; k < off; ++k) {
                    const float send = up ? acc[k] : acc[k + off];
                    const float kept = up ? acc[k + off] : acc[k];
                    acc[k] = kept + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            const float tot = acc[0] + __shfl_xor_sync(0xffffffffu, acc[0], 16);
            if (lane < CP_EMB_DIM) part[r][wq][lane] = tot;
        }
        __syncthreads();
        const int r = threadIdx.x / CP_EMB_DIM, o = threadIdx.x % CP_EMB_DIM;          // first 256 threads = SR x 16
        if (r < rows) emb[(row0 + r) * CP_EMB_DIM + o] = (part[r][0][o] + part[r][1][o]) + (part[r][2][o] + part[r][3][o]);
    });
}

// ------------------------------------------------------------------------------- backward, pass 1
// per CTA: pdw[blk][o*512 + c] = sum_r d[r,o] * a7[r,c];  p1[blk][c] = sum_r g'[r,c];  p2[blk][c] = sum_r g'[r,c]*xh[r,c]
// with a7 = dropout(BN(y)), g' = (d . Wp) * keep/(1-p), xh = (y - mean) * istd;  max|g'| -> *gmax_bits
template <bool HAS_KEEP>
__global__ void __launch_bounds__(256, 1)
proj_bwd_reduce_kernel(const float* __restrict__ y, const uint8_t* __restrict__ keep, const float* __restrict__ d_emb,
                       int64_t n, float inv_keep, const float* __restrict__ scale, const float* __restrict__ shift,
                       const float* __restrict__ mean, const float* __restrict__ istd, const float* __restrict__ Wp,
                       float* __restrict__ p1, float* __restrict__ p2, float* __restrict__ pdw,
                       unsigned int* __restrict__ gmax_bits) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int qx = threadIdx.x % QX, rl = threadIdx.x / QX;
    float4 wp[CP_EMB_DIM], acc[CP_EMB_DIM];
#pragma unroll
    for (int o = 0; o < CP_EMB_DIM; ++o) {
        wp[o] = __ldg(reinterpret_cast<const float4*>(Wp + o * F) + qx);
        acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + qx);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + qx);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + qx);
    const float4 is = __ldg(reinterpret_cast<const float4*>(istd) + qx);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    float gmax = 0.f;

    stream_rows<HAS_KEEP, true>(y, keep, d_emb, n, smem, [&](int64_t, int rows, const float4* ys, const uchar4* ks,
                                                             const float4* ds) {
        for (int r = rl; r < rows; r += 2) {
            const float4 x = ys[r * QX + qx];
            float4 a = make_float4(fmaf(x.x, sc.x, sh.x), fmaf(x.y, sc.y, sh.y), fmaf(x.z, sc.z, sh.z), fmaf(x.w, sc.w, sh.w));
            uchar4 m = make_uchar4(1, 1, 1, 1);
            if (HAS_KEEP) {
                m = ks[r * QX + qx];
                a = keep_scale(a, m, inv_keep);
            }
            const float2 alo = make_float2(a.x, a.y), ahi = make_float2(a.z, a.w);
            float2 glo[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, ghi[2] = {glo[0], glo[0]};   // 4 short chains
#pragma unroll
            for (int q = 0; q < CP_EMB_DIM / 4; ++q) {
                const float4 d4 = ds[r * (CP_EMB_DIM / 4) + q];           // broadcast read
                const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = q * 4 + j;
                    const float2 d2 = make_float2(dd[j], dd[j]);
                    ffma2(reinterpret_cast<float2*>(&acc[o])[0], d2, alo);
                    ffma2(reinterpret_cast<float2*>(&acc[o])[1], d2, ahi);
                    ffma2(glo[j & 1], d2, make_float2(wp[o].x, wp[o].y));
                    ffma2(ghi[j & 1], d2, make_float2(wp[o].z, wp[o].w));
                }
            }
            float4 g = make_float4(glo[0].x + glo[1].x, glo[0].y + glo[1].y, ghi[0].x + ghi[1].x, ghi[0].y + ghi[1].y);
            if (HAS_KEEP) g = keep_scale(g, m, inv_keep);
            gmax = fmaxf(fmaxf(gmax, fmaxf(fabsf(g.x), fabsf(g.y))), fmaxf(fabsf(g.z), fabsf(g.w)));
            s1.x += g.x; s1.y += g.y; s1.z += g.z; s1.w += g.w;
            s2.x = fmaf(g.x, (x.x - mu.x) * is.x, s2.x);
            s2.y = fmaf(g.y, (x.y - mu.y) * is.y, s2.y);
            s2.z = fmaf(g.z, (x.z - mu.z) * is.z, s2.z);
            s2.w = fmaf(g.w, (x.w - mu.w) * is.w, s2.w);
        }
    });
    // fold the odd row lane into the even one through the (now idle) ring, then write the CTA's partials
    float4* red = reinterpret_cast<float4*>(smem + BAR_BYTES);              // [18][QX] float4 = 36 KB
    if (rl == 1) {
#pragma unroll
        for (int o = 0; o < CP_EMB_DIM; ++o) red[o * QX + qx] = acc[o];
        red[16 * QX + qx] = s1;
        red[17 * QX + qx] = s2;
    }
    __syncthreads();
    if (rl == 0) {
        float* out = pdw + (int64_t)blockIdx.x * CP_EMB_DIM * F;
#pragma unroll
        for (int o = 0; o < CP_EMB_DIM; ++o) {
            const float4 u = red[o * QX + qx];
            reinterpret_cast<float4*>(out + o * F)[qx] = make_float4(acc[o].x + u.x, acc[o].y + u.y, acc[o].z + u.z, acc[o].w + u.w);
        }
        const float4 u1 = red[16 * QX + qx], u2 = red[17 * QX + qx];
        reinterpret_cast<float4*>(p1 + (int64_t)blockIdx.x * F)[qx] = make_float4(s1.x + u1.x, s1.y + u1.y, s1.z + u1.z, s1.w + u1.w);
        reinterpret_cast<float4*>(p2 + (int64_t)blockIdx.x * F)[qx] = make_float4(s2.x + u2.x, s2.y + u2.y, s2.z + u2.z, s2.w + u2.w);
    }
    if (gmax_bits) {
        gmax = warp_max(gmax);
        if (threadIdx.x % 32 == 0 && gmax > 0.f) atomicMax(gmax_bits, __float_as_uint(gmax));
    }
}

// ------------------------------------------------------------------------------- backward, pass 2
// gz = 1[y>0] * gamma*istd * (g' - m1 - xh*m2), g' recomputed from d_emb; pdb[blk][c] = sum_r gz[r,c] (bias gradient).
// SPLIT: gz written as fp16 (hi, lo) planes of gz * S (same power-of-two rule as bn_bwd_apply_kernel), 1/S -> *gscale_inv
template <bool HAS_KEEP, bool SPLIT>
__global__ void __launch_bounds__(256, 1)
proj_bwd_apply_kernel(const float* __restrict__ y, const uint8_t* __restrict__ keep, const float* __restrict__ d_emb,
                      int64_t n, float inv_keep, const float* __restrict__ mean, const float* __restrict__ istd,
                      const float* __restrict__ gamma, const float* __restrict__ m1, const float* __restrict__ m2,
                      const float* __restrict__ Wp, float* __restrict__ gz, float* __restrict__ gz_lo,
                      float* __restrict__ pdb, const unsigned int* __restrict__ gmax_bits,
                      float* __restrict__ gscale_inv, unsigned int* __restrict__ g1max_out = nullptr) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int qx = threadIdx.x % QX, rl = threadIdx.x / QX;
    float zmax = 0.f;
    float4 wp[CP_EMB_DIM];
#pragma unroll
    for (int o = 0; o < CP_EMB_DIM; ++o) wp[o] = __ldg(reinterpret_cast<const float4*>(Wp + o * F) + qx);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + qx);
    const float4 is = __ldg(reinterpret_cast<const float4*>(istd) + qx);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + qx);
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(m1) + qx);
    const float4 a2 = __ldg(reinterpret_cast<const float4*>(m2) + qx);
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
        const float bound = km * __uint_as_float(__ldg(gmax_bits)) * 18.f;
        S = plane_scale(bound);
        if (blockIdx.x == 0 && threadIdx.x == 0) *gscale_inv = 1.f / S;
    }
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f);

    stream_rows<HAS_KEEP, true>(y, keep, d_emb, n, smem, [&](int64_t row0, int rows, const float4* ys, const uchar4* ks,
                                                             const float4* ds) {
        for (int r = rl; r < rows; r += 2) {
            const float4 x = ys[r * QX + qx];
            float2 glo[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, ghi[2] = {glo[0], glo[0]};   // same order as pass 1
#pragma unroll
            for (int q = 0; q < CP_EMB_DIM / 4; ++q) {
                const float4 d4 = ds[r * (CP_EMB_DIM / 4) + q];
                const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = q * 4 + j;
                    const float2 d2 = make_float2(dd[j], dd[j]);
                    ffma2(glo[j & 1], d2, make_float2(wp[o].x, wp[o].y));
                    ffma2(ghi[j & 1], d2, make_float2(wp[o].z, wp[o].w));
                }
            }
            float4 g = make_float4(glo[0].x + glo[1].x, glo[0].y + glo[1].y, ghi[0].x + ghi[1].x, ghi[0].y + ghi[1].y);
            if (HAS_KEEP) g = keep_scale(g, ks[r * QX + qx], inv_keep);
            float4 o;
            o.x = x.x > 0.f ? k0 * (g.x - a1.x - (x.x - mu.x) * is.x * a2.x) : 0.f;
            o.y = x.y > 0.f ? k1 * (g.y - a1.y - (x.y - mu.y) * is.y * a2.y) : 0.f;
            o.z = x.z > 0.f ? k2 * (g.z - a1.z - (x.z - mu.z) * is.z * a2.z) : 0.f;
            o.w = x.w > 0.f ? k3 * (g.w - a1.w - (x.w - mu.w) * is.w * a2.w) : 0.f;
            const int64_t v = (row0 + r) * QX + qx;
            if (SPLIT)
                split_store4(make_float4(o.x * S, o.y * S, o.z * S, o.w * S), reinterpret_cast<plane_t*>(gz),
                             reinterpret_cast<plane_t*>(gz_lo), v);
            else
                reinterpret_cast<float4*>(gz)[v] = o;
            sb.x += o.x; sb.y += o.y; sb.z += o.z; sb.w += o.w;
            zmax = fmaxf(fmaxf(zmax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
        }
    });
    if (g1max_out) {                  // max |gz|: bounds the next data gradient (fused BN-backward epilogue)
        zmax = warp_max(zmax);
        if (threadIdx.x % 32 == 0 && zmax > 0.f) atomicMax(g1max_out, __float_as_uint(zmax));
    }
    float4* red = reinterpret_cast<float4*>(smem + BAR_BYTES);
    if (rl == 1) red[qx] = sb;
    __syncthreads();
    if (rl == 0) {
        const float4 u = red[qx];
        reinterpret_cast<float4*>(pdb + (int64_t)blockIdx.x * F)[qx] = make_float4(sb.x + u.x, sb.y + u.y, sb.z + u.z, sb.w + u.w);
    }
}

inline int grid_for(int64_t n) {
    const int64_t slabs = cp_cdiv(n, SR);
    return (int)(slabs < CP_NUM_SMS ? slabs : CP_NUM_SMS);
}

template <typename K>
inline int set_smem(K kernel) {
    CP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    return CP_OK;
}

}  // namespace pf
