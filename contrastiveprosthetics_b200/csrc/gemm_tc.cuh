// Tensor-core engine: fp32-accurate GEMMs on tcgen05 (kind::f16) with a 3-product fp16 split
//   x = hi + lo/2048, hi = fp16(x), lo = fp16((x - hi) * 2048);   a.b ~= hi_a.hi_b + (hi_a.lo_b + lo_a.hi_b)/2048
// (two 11-bit significands = 22 bits per operand; dropped term lo.lo <= 2^-22 |a||b|), accumulated in fp32
// in TMEM.  Same accuracy as the 3xTF32 split it replaces (tf32 also carries an 11-bit significand) at TWICE
// the tensor-pipe rate and HALF the operand bytes (HBM, TMA and shared-memory traffic); fp16's narrow exponent
// is handled by the producers: activations / weights are O(1) after BatchNorm, pre-activation gradients are
// written with a per-layer power-of-two scale (bn_bwd_apply_kernel) that the consumers' epilogues undo.
// Operands arrive pre-split as separate fp16 planes (written by the producing elementwise kernels), so the
// main loop is pure TMA -> shared memory -> tcgen05.mma with no register traffic.
//
//   gemm_tc_nt : C[M,N] = act(A[M,K] . B[N,K]^T + bias) + per-column sum / sum-of-squares partials
//                (both operands K-major, 128B-swizzled tiles).           forward + data gradient
//   gemm_tc_tn : P[z][Mo,No] = sum_{r in slab z} G[r,Mo] * A[r,No]       weight gradient (split-K)
//                (both operands MN-major, 128B-swizzled tiles)
//
// CTA = 6 warps: warp 0 TMA producer (one elected lane), warp 1 MMA issuer (one elected lane; owns
// the TMEM allocation), warps 2-5 epilogue (TMEM -> registers -> global).  Persistent: grid =
// min(#tiles, 148); two TMEM accumulators (2 x 128 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1; 3-stage shared-memory ring (A_hi, A_lo, B_hi, B_lo: 4 x 16 KB per stage).
#pragma once
#include <cstdlib>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"

typedef __half plane_t;                            // element type of the (hi, lo) operand planes
#define CP_LO_SCALE 2048.f                         // lo plane = (x - hi) * 2^11
#define CP_LO_INV (1.f / 2048.f)

// x -> (hi, lo) fp16 planes: hi = fp16(x), lo = fp16((x - hi) * 2048); x = hi + lo/2048 to 22 bits.
// Every producer multiplies by a per-tensor power of two first (plane_scale below: activations from the BatchNorm
// affine's bound, weights from max|W|, gradients from max|g'|) so that the values sit in fp16's normal range whatever
// the magnitude of the tensor; the clamp only keeps hi finite for a value 256x beyond its bound.
// S with bound * S = 2^8 (exact power of two; 1 for a zero / non-finite bound)
__device__ __forceinline__ float plane_scale(float bound) {
    return (bound > 0.f && bound < 3.0e38f) ? exp2f(8.f - ceilf(log2f(bound))) : 1.f;
}
__device__ __forceinline__ void split_f16(float x, plane_t& hi, plane_t& lo) {
    x = fminf(fmaxf(x, -65000.f), 65000.f);
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * CP_LO_SCALE);
}

namespace tcg {

constexpr int BM = 128, BN = 128, BK = 64;        // BK halves = 128 bytes = one swizzle row
constexpr int STAGES = 3;
constexpr int MAX_STAGES = 4;
constexpr int TILE_BYTES = BM * BK * 2;           // 16 KB per operand plane per stage
constexpr int STAGE_BYTES = 4 * TILE_BYTES;       // A_hi, A_lo, B_hi, B_lo
constexpr int UMMA_K = 16;                        // fp16: 32 bytes of K per instruction
constexpr int THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int TMEM_COLS = 512;                    // 2 buffers x (main + correction accumulator) x 128 fp32 columns
constexpr int ACC_COLS = 2 * BN;                  // columns per buffer
// conv-view tiles: 10 windows x 12 positions = 120 of the 128 MMA rows carry data
constexpr int CONV_WIN = 10, CONV_ROWS = 120;

struct Smem {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    float csum[4][BN];
    float csq[4][BN];
};
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + (int)sizeof(Smem);

// geometry of the K-major kernel as a function of the N tile
constexpr int OUT_BOX = 32 * 32 * 4;                  // epilogue staging: one 32 x 32 fp32 box per warp (TMA store)
template <int BN_> struct NtCfg {
    static constexpr int B_TILE = BN_ * BK * 2;
    static constexpr int STAGE = 2 * TILE_BYTES + 2 * B_TILE;
    static constexpr int NSTAGES = BN_ == 128 ? 3 : 4;
    static constexpr int ACC = 2 * BN_;               // main + correction accumulator
    static constexpr int TMEM = 4 * BN_;              // double buffered
    static constexpr int SMEM = NSTAGES * STAGE + 4 * OUT_BOX + 1024 + (int)sizeof(Smem);
};

// Epilogue store of one warp's 32 rows x 32 columns: the lane's 32 values go to the warp's staging box in
// the 128B-swizzled layout (conflict-free 16-byte shared stores), then ONE TMA store writes the box -- rows
// beyond M are clipped by the tensor map.  Direct per-lane global stores (32 half-used sectors per instruction)
// made the epilogue LSU-bound once the fp16 main loop halved the time per tile.
__device__ __forceinline__ void store_box_tma(const CUtensorMap* tm_c, uint8_t* box, const float (&v)[32], int lane,
                                              int col, int64_t row0) {
    if (lane == 0) tc::tma_store_wait_read();          // the previous store of this warp has left the box
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) =
            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    tc::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        tc::tma_store_2d(tm_c, box, col, (int)row0);
        tc::tma_store_commit();
    }
}

struct NtArgs {
    float* C; int ldc;
    const float* bias;
    float* psum; float* psq;        // [ceil(M/128)][N] or null
    int64_t M; int N; int K;
    int relu;
    const float* out_scale;         // device scalar multiplied into the result before bias (undoes the
                                    // power-of-two scale of a gradient operand), or null
    int fast;                       // CP_ENGINE_TC_FP16: hi planes only, ONE tensor-core product (11-bit operands)
    unsigned int* gmax_bits;        // non-null: atomicMax of the bit pattern of max |C| (feeds the fp16 plane scale of the
                                    // BN backward that consumes C when its reduce pass is skipped)
    const uint8_t* keep;            // non-null (with psum): C is the gradient w.r.t. a dropout output; psum then holds the
    float inv_keep;                 // column sums of C * keep / (1-p) (mask [M][N] bytes), psq is not written, and
                                    // gmax_bits takes the masked maximum
    const float* out_scale2;        // second device scalar multiplied into the result (the other operand's plane scale), or null
    // ---- EPI_BNBWD (pair kernel): C = the gradient w.r.t. the BatchNorm output of the stage below is NOT stored; the
    // epilogue turns it into the gradient w.r.t. that stage's pre-activation,
    //     gz = 1[y > 0] * (c1 * g + c2 * y + c3)        (BN backward + ReLU backward; coefficients per BN channel =
    //                                                     column % bn_period, from bn_bwd_coef_kernel)
    // and writes gz * S as fp16 (hi, lo) planes through tm_c / tm_c2, with S = plane_scale(*gz_bound); psum receives
    // the per-tile column sums of gz (bias gradient partials), *g1max_out the maximum |gz|, *gscale_inv_out = 1 / S.
    const float* Y; int ldy;        // post-ReLU pre-BN activation of the stage below, [M][N] fp32
    const float *c1, *c2, *c3;      // [bn_period]
    int bn_period;
    const float* gz_bound;          // device scalar: bound on |gz| (bn_bwd_coef_kernel)
    unsigned int* g1max_out;
    float* gscale_inv_out;
    const unsigned int* skip_flag;  // non-null: a device word; the kernel returns at once when it is zero (the conditional
                                    // plain data-gradient launch of the exact fallback, see stage_needs_exact)
    // ---- EPI_YPLANES (pair kernel): besides C (fp32, kept for the backward pass) the epilogue writes C * S as fp16
    // (hi, lo) planes through tm_c2 / tm_y -- the operand of the NEXT layer's GEMM when the BatchNorm between the two
    // is folded into that layer's weights (no BN-apply pass).  S = plane_scale(in_bound * row_l1 + bias_max): an
    // a-priori bound on |A . B^T + bias| from the bound on the input planes' values, max_j sum_k |B[j,k]| and max |bias|
    // (bit patterns of non-negative floats); *yscale_inv_out = 1 / S.
    const unsigned int *y_in_bound, *y_row_l1, *y_bias_max;
    float* yscale_inv_out;
};

// warp-transposing reduction: on return v[0] of lane l = sum over the 32 lanes of their v[l]
__device__ __forceinline__ void warp_col_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// v (this thread's 32 consecutive columns of row `row`) *= keep / (1-p); masked max -> gmax_bits; on return v[0] of
// lane l = the warp's column sum of column col + l.  The 32 mask bytes of a thread are one full 32-byte sector.
struct Mask32 { uint4 m[2]; };
__device__ __forceinline__ Mask32 load_mask32(const NtArgs& g, int64_t row, int col, bool row_ok) {
    Mask32 k;
    k.m[0] = k.m[1] = make_uint4(0, 0, 0, 0);
    if (row_ok) {
        const uint4* mp = reinterpret_cast<const uint4*>(g.keep + row * (int64_t)g.N + col);
        k.m[0] = __ldg(mp);
        k.m[1] = __ldg(mp + 1);
    }
    return k;
}
__device__ __forceinline__ void masked_col_sums(float (&v)[32], const NtArgs& g, const Mask32& k, int lane) {
    const uint8_t* mb = reinterpret_cast<const uint8_t*>(k.m);
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        v[j] = mb[j] ? v[j] * g.inv_keep : 0.f;        // rows beyond M: mask bytes 0
        mx = fmaxf(mx, fabsf(v[j]));
    }
    if (g.gmax_bits) {
        mx = warp_max(mx);
        if (lane == 0 && mx > 0.f) atomicMax(g.gmax_bits, __float_as_uint(mx));
    }
    warp_col_reduce32(v, lane);
}

// CONV: the A operand is the conv-view of a [windows][12][64] activation (k = 3 convolution as an
// implicit GEMM, K = 3 taps x 64 channels = 3 k-blocks of 64): tm_a_* are 3-D maps (64 ch, 12 pos,
// windows) and k-block kb loads the box at (0, tap-1, window0) -- positions -1 and 12 are
// out of range and arrive as zeros, which is exactly the padding of models.py:255,259.
template <int BN_, bool CONV, bool FAST>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_nt_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                  const __grid_constant__ CUtensorMap tm_c, const NtArgs g) {
    using Cfg = NtCfg<BN_>;
    constexpr int ROWS = CONV ? CONV_ROWS : BM;                 // data rows per tile
    constexpr uint32_t A_BYTES = ROWS * BK * 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* out_boxes = tiles + Cfg::NSTAGES * Cfg::STAGE;                     // 4 x 4 KB, 1024-aligned
    Smem* sm = reinterpret_cast<Smem*>(out_boxes + 4 * OUT_BOX);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int tiles_n = g.N / BN_;
    const int64_t tiles_m = (g.M + ROWS - 1) / ROWS;
    const int64_t n_tiles = tiles_m * tiles_n;
    const int kblocks = g.K / BK;

    if (warp == 0 && tc::elect_one()) {
        tc::prefetch_tmap(&tm_a_hi); tc::prefetch_tmap(&tm_a_lo);
        tc::prefetch_tmap(&tm_b_hi); tc::prefetch_tmap(&tm_b_lo);
        tc::prefetch_tmap(&tm_c);
        for (int s = 0; s < Cfg::NSTAGES; ++s) { tc::mbar_init(&sm->full[s], 1); tc::mbar_init(&sm->empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&sm->tmem_full[a], 1); tc::mbar_init(&sm->tmem_empty[a], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(&sm->tmem_base, Cfg::TMEM);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int64_t tile_m = t / tiles_n;
                const int n0 = (int)(t % tiles_n) * BN_;
                for (int kb = 0; kb < kblocks; ++kb) {
                    tc::mbar_wait(&sm->empty[s], ph ^ 1);
                    uint8_t* st = tiles + s * Cfg::STAGE;
                    tc::mbar_expect_tx(&sm->full[s], (FAST ? 1 : 2) * (A_BYTES + Cfg::B_TILE));
                    if (CONV) {
                        const int p0 = kb - 1, w0 = (int)tile_m * CONV_WIN;
                        tc::tma_load_3d(st, &tm_a_hi, &sm->full[s], 0, p0, w0);
                        if (!FAST) tc::tma_load_3d(st + TILE_BYTES, &tm_a_lo, &sm->full[s], 0, p0, w0);
                    } else {
                        tc::tma_load_2d(st, &tm_a_hi, &sm->full[s], kb * BK, (int)tile_m * BM);
                        if (!FAST) tc::tma_load_2d(st + TILE_BYTES, &tm_a_lo, &sm->full[s], kb * BK, (int)tile_m * BM);
                    }
                    tc::tma_load_2d(st + 2 * TILE_BYTES, &tm_b_hi, &sm->full[s], kb * BK, n0);
                    if (!FAST) tc::tma_load_2d(st + 2 * TILE_BYTES + Cfg::B_TILE, &tm_b_lo, &sm->full[s], kb * BK, n0);
                    if (++s == Cfg::NSTAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_f16(BM, BN_, 0, 0);
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                tc::mbar_wait(&sm->tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                // two accumulators per tile: the hi.hi chain and the (small) correction chain are kept
                // apart so that the accumulator's round-toward-zero steps on the big chain are 1/3 as many
                const uint32_t d = tmem_base + acc * Cfg::ACC;
                const uint32_t dc = d + BN_;
                for (int kb = 0; kb < kblocks; ++kb) {
                    tc::mbar_wait(&sm->full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t base = tc::smem_u32(tiles + s * Cfg::STAGE);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t ko = k * UMMA_K * 2;                       // bytes along K inside the swizzle row
                        const uint64_t a_hi = tc::smem_desc_sw128(base + ko, 16, 1024);
                        const uint64_t a_lo = tc::smem_desc_sw128(base + TILE_BYTES + ko, 16, 1024);
                        const uint64_t b_hi = tc::smem_desc_sw128(base + 2 * TILE_BYTES + ko, 16, 1024);
                        const uint64_t b_lo = tc::smem_desc_sw128(base + 2 * TILE_BYTES + Cfg::B_TILE + ko, 16, 1024);
                        if (!FAST) {
                            tc::mma_f16(dc, a_lo, b_hi, idesc, (kb | k) != 0);
                            tc::mma_f16(dc, a_hi, b_lo, idesc, 1);
                        }
                        tc::mma_f16(d, a_hi, b_hi, idesc, (kb | k) != 0);
                    }
                    tc::mma_commit(&sm->empty[s]);                                // frees the stage when the MMAs retire
                    if (++s == Cfg::NSTAGES) { s = 0; ph ^= 1; }
                }
                tc::mma_commit(&sm->tmem_full[acc]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp % 4;                      // TMEM lane quadrant this warp may access
        const int et = threadIdx.x - 64;             // 0..127
        const float oscale = (g.out_scale ? __ldg(g.out_scale) : 1.f) * (g.out_scale2 ? __ldg(g.out_scale2) : 1.f);
        const float cscale = CP_LO_INV * oscale;
        int it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const int64_t tile_m = t / tiles_n;
            const int n0 = (int)(t % tiles_n) * BN_;
            const int rl = q * 32 + lane;
            const int64_t row = tile_m * ROWS + rl;
            const bool row_ok = rl < ROWS && row < g.M;
            tc::mbar_wait(&sm->tmem_full[acc], (it >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN_ / 32; ++c) {
                float v[32], vc[32];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * Cfg::ACC + c * 32;
                tc::tmem_ld32(ta, v);
                if (!FAST) tc::tmem_ld32(ta + BN_, vc);
                tc::tmem_ld_wait();
                if (FAST) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= oscale;
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaf(vc[j], cscale, v[j] * oscale);
                }
                const int col = n0 + c * 32;
                if (g.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + j));
                        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                    }
                }
                if (g.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (g.gmax_bits && !g.keep) {
                    float mx = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fabsf(v[j]));
                    mx = warp_max(row_ok ? mx : 0.f);
                    if (lane == 0 && mx > 0.f) atomicMax(g.gmax_bits, __float_as_uint(mx));
                }
                if (CONV) {
                    // one [120 rows][32 columns] box per CTA and chunk (the tile's last 8 MMA rows carry no data)
                    if (et == 0) tc::tma_store_wait_read();
                    tc::named_bar_sync(2, EPI_THREADS);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(out_boxes + rl * 128 + ((j ^ (rl & 7)) << 4)) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    tc::fence_proxy_async();
                    tc::named_bar_sync(2, EPI_THREADS);
                    if (et == 0) {
                        tc::tma_store_2d(&tm_c, out_boxes, col, (int)(tile_m * ROWS));
                        tc::tma_store_commit();
                    }
                } else {
                    store_box_tma(&tm_c, out_boxes + q * OUT_BOX, v, lane, col, tile_m * ROWS + q * 32);
                }
                if (g.psum && g.keep) {
                    masked_col_sums(v, g, load_mask32(g, row, col, row_ok), lane);
                    sm->csum[q][c * 32 + lane] = v[0];
                    sm->csq[q][c * 32 + lane] = 0.f;
                } else if (g.psum) {
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        v[j] = row_ok ? v[j] : 0.f;
                        sq[j] = v[j] * v[j];
                    }
                    warp_col_reduce32(v, lane);
                    warp_col_reduce32(sq, lane);
                    sm->csum[q][c * 32 + lane] = v[0];
                    sm->csq[q][c * 32 + lane] = sq[0];
                }
            }
            // accumulator drained -> hand it back to the MMA warp
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sm->tmem_empty[acc]);
            if (g.psum) {
                tc::named_bar_sync(1, EPI_THREADS);
                if (et < BN_) {
                    const float s = sm->csum[0][et] + sm->csum[1][et] + sm->csum[2][et] + sm->csum[3][et];
                    const float qq = sm->csq[0][et] + sm->csq[1][et] + sm->csq[2][et] + sm->csq[3][et];
                    g.psum[tile_m * g.N + n0 + et] = s;
                    g.psq[tile_m * g.N + n0 + et] = qq;
                }
                tc::named_bar_sync(1, EPI_THREADS);
            }
        }
        if (lane == 0) tc::tma_store_wait_read();                 // shared memory must outlive the last TMA store
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, Cfg::TMEM);
}

// ---------------------------------------------------------------------------- CTA-pair forward / dgrad
// Same math as gemm_tc_nt_kernel<128>, issued as cta_group::2 MMAs on a 256 x 128 tile owned by a pair of CTAs
// (cluster 2x1x1).  The single-CTA kernel is bound by L2 -> SM bandwidth (every CTA re-fetches the full
// 128-column B tile: 512 KB of operand loads per 128 x 128 tile, ~9 TB/s chip-wide); here each CTA stages its
// own 128 rows of A but only HALF of the B tile (64 of the 128 rows) -- the tensor cores of the pair read each
// other's half -- so operand loads drop to 384 KB per 128 x 128 tile and shared-memory reads per MMA by 25 %.
// The leader CTA (cluster rank 0) issues every MMA; TMA completions of both CTAs are accounted on the leader's
// `full` barriers; tcgen05.commit multicasts the `empty` / `tmem_full` arrivals to both CTAs.  TMEM per CTA:
// 2 buffers x (128 main + 128 correction) columns, so the epilogue (8 warps) overlaps the next main loop.
namespace pair {
constexpr int THREADS2 = 320;            // TMA warp, MMA warp, 8 epilogue warps
constexpr int EPI2 = 256;
constexpr int B_HALF = (BN / 2) * BK * 2;                 // 8 KB: this CTA's 64 rows of a B plane
constexpr int STAGE2 = 2 * TILE_BYTES + 2 * B_HALF;       // A_hi, A_lo, B_hi/2, B_lo/2 = 48 KB
// operand ring depth: 4 stages when the epilogue boxes leave room for them (plain epilogue: 192 + 32 KB), 3 with the
// 64 KB of boxes of the fused BN-backward epilogue
constexpr int MAX_STAGES2 = 4;
__host__ __device__ constexpr int stages2(int epi) { return epi == 1 ? 3 : 4; }
struct Smem2 {
    uint64_t full[MAX_STAGES2], empty[MAX_STAGES2], tmem_full[2], tmem_empty[2];
    uint64_t ybar[16];              // EPI_BNBWD: arrival of the two activation boxes of each epilogue warp
    uint32_t tmem_base;
    uint32_t pad;
    float csum[4][BN];              // per TMEM-quadrant column sums of a tile (sums, then sums of squares: two phases)
};
// no alignment slack: the dynamic shared-memory array is declared __align__(1024)
constexpr int SMEM2 = stages2(0) * STAGE2 + 8 * OUT_BOX + (int)sizeof(Smem2);
// EPI_BNBWD: per epilogue warp two 4 KB boxes, one per 32-column chunk: first the TMA-loaded [32 rows][32 fp32]
// activation tile, then (once it is in registers) the chunk's [32][32] fp16 hi and lo planes on their way out
constexpr int SMEM2_BNBWD = stages2(1) * STAGE2 + 8 * 2 * OUT_BOX + (int)sizeof(Smem2);
static_assert(SMEM2 <= 232448 && SMEM2_BNBWD <= 232448, "shared memory per CTA");
}  // namespace pair

constexpr int EPI_STD = 0, EPI_BNBWD = 1, EPI_YPLANES = 2;

__device__ __forceinline__ bool nt_skip(const NtArgs& g) { return g.skip_flag && __ldg(g.skip_flag) == 0u; }

template <bool FAST, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pair::THREADS2, 1)
gemm_tc_nt_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                       const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                       const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_c2,
                       const __grid_constant__ CUtensorMap tm_y, const NtArgs g) {
    using namespace pair;
    constexpr int STAGES2 = stages2(EPI);
    if (nt_skip(g)) return;                      // uniform over the grid: taken before any barrier / TMEM allocation
    extern __shared__ __align__(1024) uint8_t smem_pair[];
    uint8_t* tiles = smem_pair;                  // 1024-aligned (128B-swizzled TMA boxes / UMMA descriptors)
    uint8_t* out_boxes = tiles + STAGES2 * STAGE2;                              // 8 x 4 KB (EPI_BNBWD: 8 x 8 KB), 1024-aligned
    Smem2* sm = reinterpret_cast<Smem2*>(out_boxes + 8 * OUT_BOX * (EPI == EPI_BNBWD ? 2 : 1));

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = tc::cluster_ctarank();
    const bool leader = rank == 0;
    const int tiles_n = g.N / BN;
    const int64_t tiles_m = (g.M + 2 * BM - 1) / (2 * BM);
    const int64_t n_tiles = tiles_m * tiles_n;
    const int kblocks = g.K / BK;
    const int64_t cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;

    if (warp == 0 && tc::elect_one()) {
        tc::prefetch_tmap(&tm_a_hi); tc::prefetch_tmap(&tm_a_lo);
        tc::prefetch_tmap(&tm_b_hi); tc::prefetch_tmap(&tm_b_lo);
        for (int s = 0; s < STAGES2; ++s) { tc::mbar_init(&sm->full[s], 1); tc::mbar_init(&sm->empty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(&sm->tmem_full[a], 1);
            tc::mbar_init(&sm->tmem_empty[a], 16);       // 8 epilogue warps x 2 CTAs arrive on the leader's copy
        }
        if (EPI == EPI_BNBWD)
            for (int a = 0; a < 16; ++a) tc::mbar_init(&sm->ybar[a], 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc_pair(&sm->tmem_base, 512);
        tc::tmem_relinquish_pair();
    }
    tc::tc_fence_before();
    tc::cluster_sync();
    tc::tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (tc::elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
                const int m0 = (int)(t / tiles_n) * 2 * BM + (int)rank * BM;
                const int n0 = (int)(t % tiles_n) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < kblocks; ++kb) {
                    tc::mbar_wait(&sm->empty[s], ph ^ 1);
                    uint8_t* st = tiles + s * STAGE2;
                    if (leader) tc::mbar_expect_tx(&sm->full[s], FAST ? STAGE2 : 2 * STAGE2);   // bytes of both CTAs
                    tc::tma_load_2d_pair(st, &tm_a_hi, &sm->full[s], kb * BK, m0);
                    if (!FAST) tc::tma_load_2d_pair(st + TILE_BYTES, &tm_a_lo, &sm->full[s], kb * BK, m0);
                    tc::tma_load_2d_pair(st + 2 * TILE_BYTES, &tm_b_hi, &sm->full[s], kb * BK, n0);
                    if (!FAST) tc::tma_load_2d_pair(st + 2 * TILE_BYTES + B_HALF, &tm_b_lo, &sm->full[s], kb * BK, n0);
                    if (++s == STAGES2) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_f16(2 * BM, BN, 0, 0);
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int64_t t = cluster_id; t < n_tiles; t += n_clusters, ++it) {
                const int acc = it & 1;
                tc::mbar_wait(&sm->tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d = tmem_base + acc * ACC_COLS;
                const uint32_t dc = d + BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    tc::mbar_wait(&sm->full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t base = tc::smem_u32(tiles + s * STAGE2);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t ko = k * UMMA_K * 2;
                        const uint64_t a_hi = tc::smem_desc_sw128(base + ko, 16, 1024);
                        const uint64_t a_lo = tc::smem_desc_sw128(base + TILE_BYTES + ko, 16, 1024);
                        const uint64_t b_hi = tc::smem_desc_sw128(base + 2 * TILE_BYTES + ko, 16, 1024);
                        const uint64_t b_lo = tc::smem_desc_sw128(base + 2 * TILE_BYTES + B_HALF + ko, 16, 1024);
                        if (!FAST) {
                            tc::mma_f16_pair(dc, a_lo, b_hi, idesc, (kb | k) != 0);
                            tc::mma_f16_pair(dc, a_hi, b_lo, idesc, 1);
                        }
                        tc::mma_f16_pair(d, a_hi, b_hi, idesc, (kb | k) != 0);
                    }
                    tc::mma_commit_pair(&sm->empty[s]);
                    if (++s == STAGES2) { s = 0; ph ^= 1; }
                }
                tc::mma_commit_pair(&sm->tmem_full[acc]);
            }
        }
    } else {
        // ===================== epilogue (warps 2..9 of both CTAs) =====================
        const int q = warp % 4;                      // TMEM lane quadrant
        const int half = (warp - 2) / 4;             // column half: 0 -> 0..63, 1 -> 64..127
        const int et = threadIdx.x - 64;             // 0..255
        const float oscale = (g.out_scale ? __ldg(g.out_scale) : 1.f) * (g.out_scale2 ? __ldg(g.out_scale2) : 1.f);
        const float cscale = CP_LO_INV * oscale;
        int it = 0;
        if (EPI == EPI_BNBWD) {
            // ---- fused BatchNorm + ReLU backward of the stage below (see NtArgs)
            const float S = plane_scale(__ldg(g.gz_bound));
            if (blockIdx.x == 0 && et == 0) *g.gscale_inv_out = 1.f / S;
            uint8_t* boxes = out_boxes + (warp - 2) * 2 * OUT_BOX;            // chunk c uses boxes + c * OUT_BOX
            uint64_t* ybar = &sm->ybar[(warp - 2) * 2];
            float gzmax = 0.f;
            for (int64_t t = cluster_id; t < n_tiles; t += n_clusters, ++it) {
                const int acc = it & 1;
                const int64_t tile_m = (t / tiles_n) * 2 + rank;
                const int n0 = (int)(t % tiles_n) * BN;
                const int row0 = (int)(tile_m * BM + q * 32);
                // the warp's two [32 x 32] activation tiles of the stage below, fetched (TMA, rows beyond M arrive as
                // zeros) while the tile's MMAs are still running
                if (lane == 0) {
                    tc::tma_store_wait_read();                     // the previous tile's plane stores have left the boxes
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        tc::mbar_expect_tx(&ybar[c], OUT_BOX);
                        tc::tma_load_2d(boxes + c * OUT_BOX, &tm_y, &ybar[c], n0 + half * 64 + c * 32, row0);
                    }
                }
                tc::mbar_wait(&sm->tmem_full[acc], (it >> 1) & 1);
                tc::tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int cl = half * 64 + c * 32;
                    uint8_t* box = boxes + c * OUT_BOX;
                    float v[32];
                    {
                        float vc[32];
                        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + cl;
                        tc::tmem_ld32(ta, v);
                        if (!FAST) tc::tmem_ld32(ta + BN, vc);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = FAST ? v[j] * oscale : fmaf(vc[j], cscale, v[j] * oscale);
                    }
                    {
                        tc::mbar_wait(&ybar[c], it & 1);
                        float4 yv[8];                                  // row `lane` of the 128B-swizzled box
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            yv[j] = *reinterpret_cast<const float4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4));
                        __syncwarp();                                  // every lane has its row: the box may be overwritten
                        const int ch0 = (n0 + cl) % g.bn_period;           // 32 consecutive BN channels
                        const float* yy = reinterpret_cast<const float*>(yv);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 k1 = __ldg(reinterpret_cast<const float4*>(g.c1 + ch0 + j));
                            const float4 k2 = __ldg(reinterpret_cast<const float4*>(g.c2 + ch0 + j));
                            const float4 k3 = __ldg(reinterpret_cast<const float4*>(g.c3 + ch0 + j));
                            const float kk1[4] = {k1.x, k1.y, k1.z, k1.w}, kk2[4] = {k2.x, k2.y, k2.z, k2.w},
                                        kk3[4] = {k3.x, k3.y, k3.z, k3.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float yj = yy[j + u];
                                const float z = yj > 0.f ? fmaf(kk1[u], v[j + u], fmaf(kk2[u], yj, kk3[u])) : 0.f;
                                gzmax = fmaxf(gzmax, fabsf(z));
                                v[j + u] = z;
                            }
                        }
                    }
                    // planes of gz * S into the chunk's box: hi [32 rows][64 B] then lo [32 rows][64 B], no swizzle
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        uint32_t hq[4], lq[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float a = fminf(fmaxf(v[j8 * 8 + 2 * u] * S, -65000.f), 65000.f);
                            const float b = fminf(fmaxf(v[j8 * 8 + 2 * u + 1] * S, -65000.f), 65000.f);
                            const __half2 h2 = __floats2half2_rn(a, b);
                            const float2 hf = __half22float2(h2);
                            const __half2 l2 = __floats2half2_rn((a - hf.x) * CP_LO_SCALE, (b - hf.y) * CP_LO_SCALE);
                            hq[u] = *reinterpret_cast<const uint32_t*>(&h2);
                            lq[u] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                        *reinterpret_cast<uint4*>(box + lane * 64 + j8 * 16) = make_uint4(hq[0], hq[1], hq[2], hq[3]);
                        *reinterpret_cast<uint4*>(box + OUT_BOX / 2 + lane * 64 + j8 * 16) = make_uint4(lq[0], lq[1], lq[2], lq[3]);
                    }
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tc::tma_store_2d(&tm_c, box, n0 + cl, row0);
                        tc::tma_store_2d(&tm_c2, box + OUT_BOX / 2, n0 + cl, row0);
                        tc::tma_store_commit();
                    }
                    if (g.psum) {                    // column sums of gz: bias gradient partials
                        warp_col_reduce32(v, lane);
                        sm->csum[q][cl + lane] = v[0];
                    }
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&sm->tmem_empty[acc]), 0));
                if (g.psum) {
                    tc::named_bar_sync(1, EPI2);
                    if (et < BN && tile_m * BM < g.M)
                        g.psum[tile_m * g.N + n0 + et] = sm->csum[0][et] + sm->csum[1][et] + sm->csum[2][et] + sm->csum[3][et];
                    tc::named_bar_sync(1, EPI2);
                }
            }
            gzmax = warp_max(gzmax);
            if (lane == 0 && gzmax > 0.f && g.g1max_out) atomicMax(g.g1max_out, __float_as_uint(gzmax));
            if (lane == 0) tc::tma_store_wait_read();
        } else {
        float yS = 1.f;
        if (EPI == EPI_YPLANES) {
            yS = plane_scale(fmaf(__uint_as_float(__ldg(g.y_in_bound)), __uint_as_float(__ldg(g.y_row_l1)),
                                  __uint_as_float(__ldg(g.y_bias_max))));
            if (blockIdx.x == 0 && et == 0) *g.yscale_inv_out = 1.f / yS;
        }
        for (int64_t t = cluster_id; t < n_tiles; t += n_clusters, ++it) {
            const int acc = it & 1;
            const int64_t tile_m = (t / tiles_n) * 2 + rank;            // 128-row tile index of this CTA
            const int n0 = (int)(t % tiles_n) * BN;
            const int64_t row = tile_m * BM + q * 32 + lane;
            const bool row_ok = row < g.M;
            float cs[2] = {0.f, 0.f}, cq[2] = {0.f, 0.f};               // this lane's column sum / sum of squares per chunk
            // dropout mask of this thread's two chunks, fetched while the tile's MMAs are still running
            Mask32 mk[2];
            if (g.keep) {
                mk[0] = load_mask32(g, row, n0 + half * 64, row_ok);
                mk[1] = load_mask32(g, row, n0 + half * 64 + 32, row_ok);
            }
            tc::mbar_wait(&sm->tmem_full[acc], (it >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int cl = half * 64 + c * 32;                      // column inside the tile
                float v[32], vc[32];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + cl;
                tc::tmem_ld32(ta, v);
                if (!FAST) tc::tmem_ld32(ta + BN, vc);
                tc::tmem_ld_wait();
                if (FAST) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= oscale;
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaf(vc[j], cscale, v[j] * oscale);
                }
                const int col = n0 + cl;
                if (g.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + j));
                        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                    }
                }
                if (g.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (g.gmax_bits && !g.keep) {
                    float mx = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fabsf(v[j]));
                    mx = warp_max(row_ok ? mx : 0.f);
                    if (lane == 0 && mx > 0.f) atomicMax(g.gmax_bits, __float_as_uint(mx));
                }
                store_box_tma(&tm_c, out_boxes + (warp - 2) * OUT_BOX, v, lane, col, tile_m * BM + q * 32);
                if (g.psum && g.keep) {
                    masked_col_sums(v, g, mk[c], lane);
                    cs[c] = v[0];
                    cq[c] = 0.f;
                } else if (g.psum) {
                    // column sums / sums of squares of the warp's 32 x 32 chunk, read back from the staged fp32 box (lane l
                    // = column l; word (r, l) of the 128B-swizzled box: conflict-free) -- two register-transposing butterfly
                    // reductions (2 x 31 shuffles + selects) cost more issue slots than the rest of the epilogue together
                    const uint8_t* box = out_boxes + (warp - 2) * OUT_BOX;
                    const int64_t left = g.M - (tile_m * BM + q * 32);
                    const int nvalid = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
                    float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        float x = *reinterpret_cast<const float*>(box + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2)));
                        x = r < nvalid ? x : 0.f;
                        s4[r & 3] += x;
                        q4[r & 3] = fmaf(x, x, q4[r & 3]);
                    }
                    cs[c] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                    cq[c] = (q4[0] + q4[1]) + (q4[2] + q4[3]);
                }
                if (EPI == EPI_YPLANES) {
                    // planes of C * S through the SAME box once the fp32 store has read it (a second box per warp would
                    // cost the operand ring its 4th stage -- measured: 184 -> 277 us per launch at K = 512):
                    // hi [32 rows][64 B] then lo [32 rows][64 B], no swizzle
                    uint8_t* pbox = out_boxes + (warp - 2) * OUT_BOX;
                    uint32_t hq[16], lq[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const float a = fminf(fmaxf(v[2 * u] * yS, -65000.f), 65000.f);
                        const float b = fminf(fmaxf(v[2 * u + 1] * yS, -65000.f), 65000.f);
                        const __half2 h2 = __floats2half2_rn(a, b);
                        const float2 hf = __half22float2(h2);
                        const __half2 l2 = __floats2half2_rn((a - hf.x) * CP_LO_SCALE, (b - hf.y) * CP_LO_SCALE);
                        hq[u] = *reinterpret_cast<const uint32_t*>(&h2);
                        lq[u] = *reinterpret_cast<const uint32_t*>(&l2);
                    }
                    if (lane == 0) tc::tma_store_wait_read();
                    __syncwarp();
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        *reinterpret_cast<uint4*>(pbox + lane * 64 + j8 * 16) = make_uint4(hq[4 * j8], hq[4 * j8 + 1], hq[4 * j8 + 2], hq[4 * j8 + 3]);
                        *reinterpret_cast<uint4*>(pbox + OUT_BOX / 2 + lane * 64 + j8 * 16) = make_uint4(lq[4 * j8], lq[4 * j8 + 1], lq[4 * j8 + 2], lq[4 * j8 + 3]);
                    }
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tc::tma_store_2d(&tm_c2, pbox, col, (int)(tile_m * BM + q * 32));
                        tc::tma_store_2d(&tm_y, pbox + OUT_BOX / 2, col, (int)(tile_m * BM + q * 32));
                        tc::tma_store_commit();
                    }
                }
            }
            // accumulator drained -> hand it back to the leader's MMA warp
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&sm->tmem_empty[acc]), 0));
            if (g.psum) {
                // sum over the 4 TMEM quadrants (warps) through one shared array: sums first, then sums of squares
                const bool writer = et < BN && tile_m * BM < g.M;
                sm->csum[q][half * 64 + lane] = cs[0];
                sm->csum[q][half * 64 + 32 + lane] = cs[1];
                tc::named_bar_sync(1, EPI2);
                if (writer)
                    g.psum[tile_m * g.N + n0 + et] = sm->csum[0][et] + sm->csum[1][et] + sm->csum[2][et] + sm->csum[3][et];
                if (g.keep) {                                   // masked column sums: no second moment
                    if (writer) g.psq[tile_m * g.N + n0 + et] = 0.f;
                    tc::named_bar_sync(1, EPI2);
                } else {
                    tc::named_bar_sync(1, EPI2);
                    sm->csum[q][half * 64 + lane] = cq[0];
                    sm->csum[q][half * 64 + 32 + lane] = cq[1];
                    tc::named_bar_sync(1, EPI2);
                    if (writer)
                        g.psq[tile_m * g.N + n0 + et] = sm->csum[0][et] + sm->csum[1][et] + sm->csum[2][et] + sm->csum[3][et];
                    tc::named_bar_sync(1, EPI2);
                }
            }
        }
        }
        if (lane == 0) tc::tma_store_wait_read();
    }
    tc::tc_fence_before();
    tc::cluster_sync();
    if (warp == 1) tc::tmem_dealloc_pair(tmem_base, 512);
}

// ---------------------------------------------------------------------------- weight gradient
// P[z][o, c] = sum over rows r of slab z of G[r, o] * A[r, c].  Both operands are MN-major: a stage
// holds, per plane, 2 blocks of [64 rows r][64 halves] (one 3-D TMA box {64, 64, 2}, plain 128B swizzle);
// MMA K-step j reads the 16 rows at +j*2048 B (two 8-row swizzle groups, SBO = 1024 B), the two 64-wide
// MN blocks are LBO = 8192 B apart.
// The fp32 accumulators in TMEM are rounded toward zero at every MMA, so a chain over ~20k rows
// would carry a ~1e-4 bias: the chain is cut every CHUNK_KB k-blocks (512 rows) and the epilogue
// threads keep the running sum in registers (round-to-nearest adds), 128 per thread.
constexpr int CHUNK_KB = 8;
constexpr int MN_BLOCK = BK * 128;                // bytes of one [64 rows][64 halves] block

struct TnArgs {
    float* P;                 // [splits][Mo][No]
    int Mo, No;
    int64_t R;                // total rows
    int64_t rows_per_split;   // multiple of BK
    int fast;                 // hi planes only
};

template <bool FAST>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_tn_kernel(const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
                  const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const TnArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Smem* sm = reinterpret_cast<Smem*>(tiles + STAGES * STAGE_BYTES);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int o0 = blockIdx.y * BM, c0 = blockIdx.x * BN;
    const int64_t r_begin = (int64_t)blockIdx.z * g.rows_per_split;
    const int64_t r_end = min(g.R, r_begin + g.rows_per_split);
    const int kblocks = (int)((r_end - r_begin + BK - 1) / BK);
    const int chunks = (kblocks + CHUNK_KB - 1) / CHUNK_KB;

    if (warp == 0 && tc::elect_one()) {
        tc::prefetch_tmap(&tm_g_hi); tc::prefetch_tmap(&tm_g_lo);
        tc::prefetch_tmap(&tm_a_hi); tc::prefetch_tmap(&tm_a_lo);
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&sm->full[s], 1); tc::mbar_init(&sm->empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&sm->tmem_full[a], 1); tc::mbar_init(&sm->tmem_empty[a], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(&sm->tmem_base, TMEM_COLS);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        if (tc::elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                tc::mbar_wait(&sm->empty[s], ph ^ 1);
                uint8_t* st = tiles + s * STAGE_BYTES;
                const int r = (int)(r_begin + (int64_t)kb * BK);
                tc::mbar_expect_tx(&sm->full[s], FAST ? STAGE_BYTES / 2 : STAGE_BYTES);
                tc::tma_load_3d(st + 0 * TILE_BYTES, &tm_g_hi, &sm->full[s], 0, r, o0 / 64);
                if (!FAST) tc::tma_load_3d(st + 1 * TILE_BYTES, &tm_g_lo, &sm->full[s], 0, r, o0 / 64);
                tc::tma_load_3d(st + 2 * TILE_BYTES, &tm_a_hi, &sm->full[s], 0, r, c0 / 64);
                if (!FAST) tc::tma_load_3d(st + 3 * TILE_BYTES, &tm_a_lo, &sm->full[s], 0, r, c0 / 64);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_f16(BM, BN, 1, 1);
            int s = 0; uint32_t ph = 0;
            int kb = 0;
            for (int ch = 0; ch < chunks; ++ch) {
                const int acc = ch & 1;
                tc::mbar_wait(&sm->tmem_empty[acc], ((ch >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d = tmem_base + acc * ACC_COLS;
                const uint32_t dc = d + BN;
                const int kb_end = min(kblocks, kb + CHUNK_KB);
                for (bool first = true; kb < kb_end; ++kb) {
                    tc::mbar_wait(&sm->full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t base = tc::smem_u32(tiles + s * STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t ko = k * 2048;                 // 16 rows of 128 B
                        const uint64_t g_hi = tc::smem_desc_sw128(base + 0 * TILE_BYTES + ko, MN_BLOCK, 1024);
                        const uint64_t g_lo = tc::smem_desc_sw128(base + 1 * TILE_BYTES + ko, MN_BLOCK, 1024);
                        const uint64_t a_hi = tc::smem_desc_sw128(base + 2 * TILE_BYTES + ko, MN_BLOCK, 1024);
                        const uint64_t a_lo = tc::smem_desc_sw128(base + 3 * TILE_BYTES + ko, MN_BLOCK, 1024);
                        const uint32_t accum = (first && k == 0) ? 0u : 1u;
                        if (!FAST) {
                            tc::mma_f16(dc, g_lo, a_hi, idesc, accum);
                            tc::mma_f16(dc, g_hi, a_lo, idesc, 1);
                        }
                        tc::mma_f16(d, g_hi, a_hi, idesc, accum);
                    }
                    first = false;
                    tc::mma_commit(&sm->empty[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                tc::mma_commit(&sm->tmem_full[acc]);
            }
        }
    } else {
        const int q = warp % 4;
        float sum[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) sum[j] = 0.f;
        for (int ch = 0; ch < chunks; ++ch) {
            const int acc = ch & 1;
            tc::mbar_wait(&sm->tmem_full[acc], (ch >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                float v[32], vc[32];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + c * 32;
                tc::tmem_ld32(ta, v);
                if (!FAST) tc::tmem_ld32(ta + BN, vc);
                tc::tmem_ld_wait();
                if (FAST) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c * 32 + j] += v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c * 32 + j] += fmaf(vc[j], CP_LO_INV, v[j]);
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sm->tmem_empty[acc]);
        }
        float* dst = g.P + ((int64_t)blockIdx.z * g.Mo + o0 + q * 32 + lane) * g.No + c0;
#pragma unroll
        for (int j = 0; j < BN; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------- CTA-pair weight gradient
// Same math as gemm_tc_tn_kernel on a 256 (o) x 128 (c) tile owned by a pair of CTAs (cluster 2x1x1, cta_group::2
// MMAs issued by the leader).  The single-CTA kernel streams 64 KB of operands per k-block and 128 x 128 tile --
// 83 B/clk/SM at the tensor rate of the 3-product split, the L2 -> SM wall the K-major kernel hit before it was
// paired.  Here each CTA stages its own 128 o-columns of G (2 MN blocks per plane) but only ONE 64-wide MN block of
// the A planes -- the pair's tensor cores read each other's half -- so a k-block costs 48 KB per CTA (62 B/clk/SM)
// and the ring is 4 stages deep.  Barriers as in gemm_tc_nt_pair_kernel: TMA completions of both CTAs land on the
// leader's `full`, tcgen05.commit multicasts `empty` / `tmem_full` to both CTAs, the epilogue warps of both CTAs hand
// the accumulator back on the leader's `tmem_empty`.  Chain cutting and the register-side running sums are unchanged.
namespace tnp {
constexpr int A_HALF = (BN / 2) * BK * 2;                 // 8 KB: this CTA's [64 rows r][64 halves] block of an A plane
constexpr int STAGE = 2 * TILE_BYTES + 2 * A_HALF;        // G_hi, G_lo, A_hi/2, A_lo/2 = 48 KB
constexpr int NSTAGES = 4;
constexpr int SMEM = NSTAGES * STAGE + (int)sizeof(Smem); // the dynamic array is declared __align__(1024)
static_assert(NSTAGES <= MAX_STAGES && SMEM <= 232448, "shared memory per CTA");
}  // namespace tnp

template <bool FAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_tc_tn_pair_kernel(const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
                       const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                       const TnArgs g) {
    using namespace tnp;
    extern __shared__ __align__(1024) uint8_t smem_tnp[];
    uint8_t* tiles = smem_tnp;
    Smem* sm = reinterpret_cast<Smem*>(tiles + NSTAGES * STAGE);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = tc::cluster_ctarank();
    const bool leader = rank == 0;
    const int o0 = (blockIdx.x / 2) * 2 * BM + (int)rank * BM;      // this CTA's 128 output rows
    const int c0 = blockIdx.y * BN;                                 // the pair's 128 output columns
    const int64_t r_begin = (int64_t)blockIdx.z * g.rows_per_split;
    const int64_t r_end = min(g.R, r_begin + g.rows_per_split);
    const int kblocks = (int)((r_end - r_begin + BK - 1) / BK);
    const int chunks = (kblocks + CHUNK_KB - 1) / CHUNK_KB;

    if (warp == 0 && tc::elect_one()) {
        tc::prefetch_tmap(&tm_g_hi); tc::prefetch_tmap(&tm_g_lo);
        tc::prefetch_tmap(&tm_a_hi); tc::prefetch_tmap(&tm_a_lo);
        for (int s = 0; s < NSTAGES; ++s) { tc::mbar_init(&sm->full[s], 1); tc::mbar_init(&sm->empty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(&sm->tmem_full[a], 1);
            tc::mbar_init(&sm->tmem_empty[a], 8);        // 4 epilogue warps x 2 CTAs arrive on the leader's copy
        }
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc_pair(&sm->tmem_base, TMEM_COLS);
        tc::tmem_relinquish_pair();
    }
    tc::tc_fence_before();
    tc::cluster_sync();
    tc::tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        if (tc::elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                tc::mbar_wait(&sm->empty[s], ph ^ 1);
                uint8_t* st = tiles + s * STAGE;
                const int r = (int)(r_begin + (int64_t)kb * BK);
                if (leader) tc::mbar_expect_tx(&sm->full[s], FAST ? STAGE : 2 * STAGE);     // bytes of both CTAs
                tc::tma_load_3d_pair(st, &tm_g_hi, &sm->full[s], 0, r, o0 / 64);
                if (!FAST) tc::tma_load_3d_pair(st + TILE_BYTES, &tm_g_lo, &sm->full[s], 0, r, o0 / 64);
                tc::tma_load_3d_pair(st + 2 * TILE_BYTES, &tm_a_hi, &sm->full[s], 0, r, c0 / 64 + (int)rank);
                if (!FAST) tc::tma_load_3d_pair(st + 2 * TILE_BYTES + A_HALF, &tm_a_lo, &sm->full[s], 0, r, c0 / 64 + (int)rank);
                if (++s == NSTAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader && tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_f16(2 * BM, BN, 1, 1);
            int s = 0; uint32_t ph = 0;
            int kb = 0;
            for (int ch = 0; ch < chunks; ++ch) {
                const int acc = ch & 1;
                tc::mbar_wait(&sm->tmem_empty[acc], ((ch >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d = tmem_base + acc * ACC_COLS;
                const uint32_t dc = d + BN;
                const int kb_end = min(kblocks, kb + CHUNK_KB);
                for (bool first = true; kb < kb_end; ++kb) {
                    tc::mbar_wait(&sm->full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t base = tc::smem_u32(tiles + s * STAGE);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t ko = k * 2048;                 // 16 rows of 128 B
                        const uint64_t g_hi = tc::smem_desc_sw128(base + ko, MN_BLOCK, 1024);
                        const uint64_t g_lo = tc::smem_desc_sw128(base + TILE_BYTES + ko, MN_BLOCK, 1024);
                        const uint64_t a_hi = tc::smem_desc_sw128(base + 2 * TILE_BYTES + ko, MN_BLOCK, 1024);
                        const uint64_t a_lo = tc::smem_desc_sw128(base + 2 * TILE_BYTES + A_HALF + ko, MN_BLOCK, 1024);
                        const uint32_t accum = (first && k == 0) ? 0u : 1u;
                        if (!FAST) {
                            tc::mma_f16_pair(dc, g_lo, a_hi, idesc, accum);
                            tc::mma_f16_pair(dc, g_hi, a_lo, idesc, 1);
                        }
                        tc::mma_f16_pair(d, g_hi, a_hi, idesc, accum);
                    }
                    first = false;
                    tc::mma_commit_pair(&sm->empty[s]);
                    if (++s == NSTAGES) { s = 0; ph ^= 1; }
                }
                tc::mma_commit_pair(&sm->tmem_full[acc]);
            }
        }
    } else {
        const int q = warp % 4;
        float sum[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) sum[j] = 0.f;
        for (int ch = 0; ch < chunks; ++ch) {
            const int acc = ch & 1;
            tc::mbar_wait(&sm->tmem_full[acc], (ch >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                float v[32], vc[32];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + c * 32;
                tc::tmem_ld32(ta, v);
                if (!FAST) tc::tmem_ld32(ta + BN, vc);
                tc::tmem_ld_wait();
                if (FAST) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c * 32 + j] += v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c * 32 + j] += fmaf(vc[j], CP_LO_INV, v[j]);
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(tc::mapa(tc::smem_u32(&sm->tmem_empty[acc]), 0));
        }
        float* dst = g.P + ((int64_t)blockIdx.z * g.Mo + o0 + q * 32 + lane) * g.No + c0;
#pragma unroll
        for (int j = 0; j < BN; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
    }
    tc::tc_fence_before();
    tc::cluster_sync();
    if (warp == 1) tc::tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------- conv2 weight gradient
// P[z][tap*64 + c][o] = sum over the windows of slab z, positions p, of X[w, p+tap-1, c] * G[(w,p), o]
// (the transposed weight gradient of the k = 3 convolution).  M side = conv-view of X (192 columns,
// padded to two 128-row MMA tiles: tile 0 = taps 0,1, tile 1 = tap 2 + 64 unused rows), N side = the
// 64 output channels of G.  A k-block is 4 windows = 48 rows: per plane the M side is up to two
// [48 rows][64 ch] boxes of the 3-D conv map shifted by tap-1 (zero padded by TMA), the N side one
// [48][64] box of G.  Same MN-major layout / register-side chain cutting as gemm_tc_tn.
constexpr int CW_WIN = 4, CW_ROWS = 48;                   // windows / rows per k-block
constexpr int CW_BLOCK = CW_ROWS * 128;                   // bytes of one [48][64 halves] block
constexpr int CW_STAGE = (2 + 1) * 2 * CW_BLOCK;          // (X: 2 blocks, G: 1 block) x (hi, lo)
constexpr int CW_STAGES = 3;
constexpr int CW_CHUNK_KB = 11;                           // 528 rows per accumulation chain
constexpr int CW_SMEM = CW_STAGES * CW_STAGE + 1024 + (int)sizeof(Smem);

struct CwArgs {
    float* P;                    // [splits][256][64]
    int64_t windows;
    int64_t win_per_split;       // multiple of CW_WIN
    int fast;                    // hi planes only
};

template <bool FAST>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_tn_conv_kernel(const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                       const __grid_constant__ CUtensorMap tm_g_hi, const __grid_constant__ CUtensorMap tm_g_lo,
                       const CwArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    Smem* sm = reinterpret_cast<Smem*>(tiles + CW_STAGES * CW_STAGE);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int mt = blockIdx.x;                               // 0: taps 0,1   1: tap 2
    const int n_blocks = mt == 0 ? 2 : 1;                    // 64-channel blocks (taps) that carry data
    const int64_t w_begin = (int64_t)blockIdx.y * g.win_per_split;
    const int64_t w_end = min(g.windows, w_begin + g.win_per_split);
    const int kblocks = (int)((w_end - w_begin + CW_WIN - 1) / CW_WIN);
    const int chunks = (kblocks + CW_CHUNK_KB - 1) / CW_CHUNK_KB;

    if (warp == 0 && tc::elect_one()) {
        tc::prefetch_tmap(&tm_x_hi); tc::prefetch_tmap(&tm_x_lo);
        tc::prefetch_tmap(&tm_g_hi); tc::prefetch_tmap(&tm_g_lo);
        for (int s = 0; s < CW_STAGES; ++s) { tc::mbar_init(&sm->full[s], 1); tc::mbar_init(&sm->empty[s], 1); }
        for (int a = 0; a < 2; ++a) { tc::mbar_init(&sm->tmem_full[a], 1); tc::mbar_init(&sm->tmem_empty[a], 4); }
        tc::fence_barrier_init();
    }
    if (warp == 1) {
        tc::tmem_alloc(&sm->tmem_base, 256);                 // 2 buffers x (main + correction) x 64 columns
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;

    if (warp == 0) {
        if (tc::elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < kblocks; ++kb) {
                tc::mbar_wait(&sm->empty[s], ph ^ 1);
                uint8_t* st = tiles + s * CW_STAGE;
                const int w0 = (int)(w_begin + (int64_t)kb * CW_WIN);
                tc::mbar_expect_tx(&sm->full[s], (uint32_t)(n_blocks + 1) * (FAST ? 1 : 2) * CW_BLOCK);
                for (int b = 0; b < n_blocks; ++b) {
                    const int tap = mt * 2 + b;
                    tc::tma_load_3d(st + b * CW_BLOCK, &tm_x_hi, &sm->full[s], 0, tap - 1, w0);
                    if (!FAST) tc::tma_load_3d(st + (2 + b) * CW_BLOCK, &tm_x_lo, &sm->full[s], 0, tap - 1, w0);
                }
                // G: [rows][64] -> one [48][64] block per plane
                tc::tma_load_2d(st + 4 * CW_BLOCK, &tm_g_hi, &sm->full[s], 0, w0 * 12);
                if (!FAST) tc::tma_load_2d(st + 5 * CW_BLOCK, &tm_g_lo, &sm->full[s], 0, w0 * 12);
                if (++s == CW_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (tc::elect_one()) {
            constexpr uint32_t idesc = tc::idesc_f16(BM, 64, 1, 1);
            int s = 0; uint32_t ph = 0;
            int kb = 0;
            for (int ch = 0; ch < chunks; ++ch) {
                const int acc = ch & 1;
                tc::mbar_wait(&sm->tmem_empty[acc], ((ch >> 1) & 1) ^ 1);
                tc::tc_fence_after();
                const uint32_t d = tmem_base + acc * 128;
                const uint32_t dc = d + 64;
                const int kb_end = min(kblocks, kb + CW_CHUNK_KB);
                for (bool first = true; kb < kb_end; ++kb) {
                    tc::mbar_wait(&sm->full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t base = tc::smem_u32(tiles + s * CW_STAGE);
#pragma unroll
                    for (int k = 0; k < CW_ROWS / UMMA_K; ++k) {
                        const uint32_t ko = k * 2048;
                        const uint64_t x_hi = tc::smem_desc_sw128(base + 0 * CW_BLOCK + ko, CW_BLOCK, 1024);
                        const uint64_t x_lo = tc::smem_desc_sw128(base + 2 * CW_BLOCK + ko, CW_BLOCK, 1024);
                        const uint64_t g_hi = tc::smem_desc_sw128(base + 4 * CW_BLOCK + ko, CW_BLOCK, 1024);
                        const uint64_t g_lo = tc::smem_desc_sw128(base + 5 * CW_BLOCK + ko, CW_BLOCK, 1024);
                        const uint32_t accum = (first && k == 0) ? 0u : 1u;
                        if (!FAST) {
                            tc::mma_f16(dc, x_lo, g_hi, idesc, accum);
                            tc::mma_f16(dc, x_hi, g_lo, idesc, 1);
                        }
                        tc::mma_f16(d, x_hi, g_hi, idesc, accum);
                    }
                    first = false;
                    tc::mma_commit(&sm->empty[s]);
                    if (++s == CW_STAGES) { s = 0; ph ^= 1; }
                }
                tc::mma_commit(&sm->tmem_full[acc]);
            }
        }
    } else {
        const int q = warp % 4;
        float sum[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) sum[j] = 0.f;
        for (int ch = 0; ch < chunks; ++ch) {
            const int acc = ch & 1;
            tc::mbar_wait(&sm->tmem_full[acc], (ch >> 1) & 1);
            tc::tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float v[32], vc[32];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 128 + c * 32;
                tc::tmem_ld32(ta, v);
                if (!FAST) tc::tmem_ld32(ta + 64, vc);
                tc::tmem_ld_wait();
                if (FAST) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c * 32 + j] += v[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c * 32 + j] += fmaf(vc[j], CP_LO_INV, v[j]);
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sm->tmem_empty[acc]);
        }
        const int m = mt * 128 + q * 32 + lane;
        if (m < 192) {
            float* dst = g.P + ((int64_t)blockIdx.y * 256 + m) * 64;
#pragma unroll
            for (int j = 0; j < 64; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major fp16 [rows, cols] (leading dimension ld halves), box {64 cols, box_rows}, 128B swizzle,
// out-of-range elements read as zero
inline int make_tmap_2d(CUtensorMap* m, const plane_t* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return CP_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<plane_t*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CP_OK : CP_ERR_ARG;
}

// row-major fp16 [rows, cols] viewed as (64 cols, rows, cols/64 blocks): box {64, 64 rows, 2 blocks}
// lands in shared memory as 2 x [64 rows][128 B], i.e. the MN-major 128B-swizzle canonical layout
inline int make_tmap_mn(CUtensorMap* m, const plane_t* base, int64_t rows, int64_t cols, int64_t ld, int blocks = 2) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return CP_ERR_UNSUPPORTED;
    cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(cols / 64)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 128};
    cuuint32_t box[3] = {64, 64, (cuuint32_t)blocks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<plane_t*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CP_OK : CP_ERR_ARG;
}

// CTA-pair (cta_group::2) weight-gradient kernel; CP_TN_PAIR=0 in the environment selects the single-CTA kernel (A/B runs)
inline bool use_tn_pair() {
    static const bool on = [] { const char* e = getenv("CP_TN_PAIR"); return !(e && e[0] == '0'); }();
    return on;
}
// returns the number of splits written to P ([splits][Mo][No]) through *splits_out
inline int launch_tn(const plane_t* G_hi, const plane_t* G_lo, int ldg, int Mo, const plane_t* A_hi, const plane_t* A_lo,
                     int lda, int No, int64_t R, float* P, size_t p_capacity_elems, int* splits_out,
                     cudaStream_t st, int fast = 0, bool alone = false) {
    if (Mo % BM != 0 || No % BN != 0 || ldg % 8 != 0 || lda % 8 != 0 || R <= 0) return CP_ERR_ARG;
    CUtensorMap tg_hi, tg_lo, ta_hi, ta_lo;
    int rc;
    if ((rc = make_tmap_mn(&tg_hi, G_hi, R, Mo, ldg)) != CP_OK) return rc;
    if ((rc = make_tmap_mn(&tg_lo, G_lo, R, Mo, ldg)) != CP_OK) return rc;
    if ((rc = make_tmap_mn(&ta_hi, A_hi, R, No, lda)) != CP_OK) return rc;
    if ((rc = make_tmap_mn(&ta_lo, A_lo, R, No, lda)) != CP_OK) return rc;
    CP_ONCE_PER_DEVICE({
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    });
    const int tiles = (Mo / BM) * (No / BN);
    // one wave on ~2/3 of the SMs: this GEMM runs on a side stream next to the HBM-bound BN-backward kernels, and
    // each of its CTAs pins 48 K registers (255 x 192 threads), leaving room for ONE 256-thread BN CTA on that SM.
    // Measured at M = 167,936: 9 splits (144 CTAs) 11.10 ms/step, 6 splits 10.96, 4 splits 11.06, 3 splits 11.35
    int S = (CP_NUM_SMS * 2 / 3) / tiles;
    if (alone) S = CP_NUM_SMS / tiles;                            // in line on the caller's stream: every SM
    const int64_t max_s = cp_cdiv(R, (int64_t)BK * CHUNK_KB);
    if (S > max_s) S = (int)max_s;
    const int64_t cap = (int64_t)(p_capacity_elems / ((size_t)Mo * No));
    if (S > cap) S = (int)cap;
    if (S < 1) S = 1;
    const int64_t rps = cp_cdiv(cp_cdiv(R, S), BK) * BK;
    S = (int)cp_cdiv(R, rps);
    TnArgs g{P, Mo, No, R, rps, fast};
    if (use_tn_pair() && Mo % (2 * BM) == 0) {
        CUtensorMap ta_hi1, ta_lo1;                                   // A boxes of ONE 64-wide MN block: half a tile per CTA
        if ((rc = make_tmap_mn(&ta_hi1, A_hi, R, No, lda, 1)) != CP_OK) return rc;
        if ((rc = make_tmap_mn(&ta_lo1, A_lo, R, No, lda, 1)) != CP_OK) return rc;
        CP_ONCE_PER_DEVICE({
            CP_CUDA(cudaFuncSetAttribute(gemm_tc_tn_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tnp::SMEM));
            CP_CUDA(cudaFuncSetAttribute(gemm_tc_tn_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tnp::SMEM));
        });
        const dim3 grid(2 * (Mo / (2 * BM)), No / BN, S);
        if (fast) gemm_tc_tn_pair_kernel<true><<<grid, THREADS, tnp::SMEM, st>>>(tg_hi, tg_lo, ta_hi1, ta_lo1, g);
        else gemm_tc_tn_pair_kernel<false><<<grid, THREADS, tnp::SMEM, st>>>(tg_hi, tg_lo, ta_hi1, ta_lo1, g);
        CP_CHECK_LAUNCH();
        *splits_out = S;
        return CP_OK;
    }
    if (fast) gemm_tc_tn_kernel<true><<<dim3(No / BN, Mo / BM, S), THREADS, SMEM_BYTES, st>>>(tg_hi, tg_lo, ta_hi, ta_lo, g);
    else gemm_tc_tn_kernel<false><<<dim3(No / BN, Mo / BM, S), THREADS, SMEM_BYTES, st>>>(tg_hi, tg_lo, ta_hi, ta_lo, g);
    CP_CHECK_LAUNCH();
    *splits_out = S;
    return CP_OK;
}

// [windows][12][64] fp16 activation as (64 channels, position, window) boxes {64, 12, box_windows}
inline int make_tmap_conv(CUtensorMap* m, const plane_t* base, int64_t windows, int box_windows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return CP_ERR_UNSUPPORTED;
    cuuint64_t dims[3] = {64, 12, (cuuint64_t)windows};
    cuuint64_t strides[2] = {64 * 2, 12 * 64 * 2};
    cuuint32_t box[3] = {64, 12, (cuuint32_t)box_windows};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<plane_t*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CP_OK : CP_ERR_ARG;
}

// fp32 output [rows, cols] (leading dimension ld floats): 32 x 32 boxes, 128B swizzle (TMA store, rows clipped)
inline int make_tmap_out(CUtensorMap* m, float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows = 32) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return CP_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CP_OK : CP_ERR_ARG;
}

// fp16 plane output [rows, cols] (leading dimension ld halves): 32 x 32 boxes (64-byte rows), no swizzle (TMA store)
inline int make_tmap_plane_out(CUtensorMap* m, plane_t* base, int64_t rows, int64_t cols, int64_t ld) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return CP_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? CP_OK : CP_ERR_ARG;
}

template <int BN_, bool CONV, bool FAST>
inline int launch_nt_cfg(const CUtensorMap& ta_hi, const CUtensorMap& ta_lo, const CUtensorMap& tb_hi,
                         const CUtensorMap& tb_lo, const CUtensorMap& tc_out, const NtArgs& g, int64_t tiles_m,
                         cudaStream_t st) {
    CP_ONCE_PER_DEVICE({
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_nt_kernel<BN_, CONV, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     NtCfg<BN_>::SMEM));
    });
    const int64_t n_tiles = tiles_m * (g.N / BN_);
    const int grid = (int)(n_tiles < CP_NUM_SMS ? n_tiles : CP_NUM_SMS);
    gemm_tc_nt_kernel<BN_, CONV, FAST><<<grid, THREADS, NtCfg<BN_>::SMEM, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, tc_out, g);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

static bool g_use_pair = true;       // CTA-pair (cta_group::2) kernel for the plain (non-conv) K-major GEMMs
inline int set_pair_attrs() {
    CP_ONCE_PER_DEVICE({
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_nt_pair_kernel<false, EPI_STD>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::SMEM2));
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_nt_pair_kernel<true, EPI_STD>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::SMEM2));
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_nt_pair_kernel<false, EPI_BNBWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::SMEM2_BNBWD));
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_nt_pair_kernel<true, EPI_BNBWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::SMEM2_BNBWD));
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_nt_pair_kernel<false, EPI_YPLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::SMEM2));
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_nt_pair_kernel<true, EPI_YPLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::SMEM2));
    });
    return CP_OK;
}
// plane outputs of launch_nt (EPI_YPLANES, see NtArgs)
struct YPlanes {
    plane_t *hi, *lo;                                       // [M][N] each
    const unsigned int *in_bound, *row_l1, *bias_max;       // device words (bit patterns)
    float* scale_inv_out;                                   // device scalar
};
inline int launch_nt(const plane_t* A_hi, const plane_t* A_lo, int64_t M, int K, int lda, const plane_t* B_hi,
                     const plane_t* B_lo, int N, int ldb, const float* bias, float* C, int ldc, float* psum,
                     float* psq, int relu, cudaStream_t st, const float* out_scale = nullptr, int fast = 0,
                     unsigned int* gmax_bits = nullptr, const uint8_t* keep = nullptr, float inv_keep = 1.f,
                     const float* out_scale2 = nullptr, const unsigned int* skip_flag = nullptr,
                     const YPlanes* yp = nullptr) {
    if (K % BK != 0 || N % BN != 0 || lda % 8 != 0 || ldb % 8 != 0 || ldc % 4 != 0) return CP_ERR_ARG;
    if (yp && !(g_use_pair && M > BM)) return CP_ERR_UNSUPPORTED;      // the plane epilogue exists in the pair kernel only
    if (keep && (ldc != N || ((uintptr_t)keep) % 16 != 0)) return CP_ERR_ARG;     // mask laid out like a dense [M][N] C
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo, tc_out;
    int rc;
    if ((rc = make_tmap_out(&tc_out, C, M, N, ldc)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&ta_hi, A_hi, M, K, lda, BM)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&ta_lo, A_lo, M, K, lda, BM)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tb_hi, B_hi, N, K, ldb, BN)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tb_lo, B_lo, N, K, ldb, BN)) != CP_OK) return rc;
    NtArgs g{C, ldc, bias, psum, psq, M, N, K, relu, out_scale, fast, gmax_bits, keep, inv_keep, out_scale2};
    g.skip_flag = skip_flag;
    if (g_use_pair && M > BM) {
        CUtensorMap tb_hi2, tb_lo2;                                   // B boxes of 64 rows: half a tile per CTA
        if ((rc = make_tmap_2d(&tb_hi2, B_hi, N, K, ldb, BN / 2)) != CP_OK) return rc;
        if ((rc = make_tmap_2d(&tb_lo2, B_lo, N, K, ldb, BN / 2)) != CP_OK) return rc;
        if ((rc = set_pair_attrs()) != CP_OK) return rc;
        const int64_t n_tiles = cp_cdiv(M, 2 * BM) * (N / BN);
        const int clusters = (int)(n_tiles < CP_NUM_SMS / 2 ? n_tiles : CP_NUM_SMS / 2);
        if (yp) {
            CUtensorMap ty_hi, ty_lo;
            if ((rc = make_tmap_plane_out(&ty_hi, yp->hi, M, N, N)) != CP_OK) return rc;
            if ((rc = make_tmap_plane_out(&ty_lo, yp->lo, M, N, N)) != CP_OK) return rc;
            g.y_in_bound = yp->in_bound; g.y_row_l1 = yp->row_l1; g.y_bias_max = yp->bias_max;
            g.yscale_inv_out = yp->scale_inv_out;
            if (fast) gemm_tc_nt_pair_kernel<true, EPI_YPLANES><<<2 * clusters, pair::THREADS2, pair::SMEM2, st>>>(ta_hi, ta_lo, tb_hi2, tb_lo2, tc_out, ty_hi, ty_lo, g);
            else gemm_tc_nt_pair_kernel<false, EPI_YPLANES><<<2 * clusters, pair::THREADS2, pair::SMEM2, st>>>(ta_hi, ta_lo, tb_hi2, tb_lo2, tc_out, ty_hi, ty_lo, g);
            CP_CHECK_LAUNCH();
            return CP_OK;
        }
        if (fast) gemm_tc_nt_pair_kernel<true, EPI_STD><<<2 * clusters, pair::THREADS2, pair::SMEM2, st>>>(ta_hi, ta_lo, tb_hi2, tb_lo2, tc_out, tc_out, tc_out, g);
        else gemm_tc_nt_pair_kernel<false, EPI_STD><<<2 * clusters, pair::THREADS2, pair::SMEM2, st>>>(ta_hi, ta_lo, tb_hi2, tb_lo2, tc_out, tc_out, tc_out, g);
        CP_CHECK_LAUNCH();
        return CP_OK;
    }
    if (skip_flag) return CP_ERR_UNSUPPORTED;                         // conditional launches exist for the pair kernel only
    return fast ? launch_nt_cfg<128, false, true>(ta_hi, ta_lo, tb_hi, tb_lo, tc_out, g, cp_cdiv(M, BM), st)
                : launch_nt_cfg<128, false, false>(ta_hi, ta_lo, tb_hi, tb_lo, tc_out, g, cp_cdiv(M, BM), st);
}

// Data gradient of a linear layer fused with the BatchNorm + ReLU backward of the stage below (EPI_BNBWD, see NtArgs):
//   g = A . B^T * out_scale * out_scale2 ;  gz = 1[Y > 0] (c1 g + c2 Y + c3) ;  planes (G_hi, G_lo) = split(gz * S)
// A: [M][K] planes of the layer's pre-activation gradient, B: [N][K] planes of W^T, Y: [M][N] fp32, G planes [M][N].
// pdb: [ceil(M/128)][N] column sums of gz.  Needs M > 128 (the CTA-pair kernel).
inline bool bnbwd_supported(int64_t M) { return g_use_pair && M > BM; }
inline int launch_nt_bnbwd(const plane_t* A_hi, const plane_t* A_lo, int64_t M, int K, const plane_t* B_hi,
                           const plane_t* B_lo, int N, const float* Y, const float* c1, const float* c2, const float* c3,
                           int bn_period, const float* gz_bound, plane_t* G_hi, plane_t* G_lo, float* pdb,
                           unsigned int* g1max_out, float* gscale_inv_out, const float* out_scale,
                           const float* out_scale2, int fast, cudaStream_t st) {
    if (!bnbwd_supported(M) || K % BK != 0 || N % BN != 0 || bn_period % 32 != 0 || ((uintptr_t)Y) % 16 != 0)
        return CP_ERR_ARG;
    CUtensorMap ta_hi, ta_lo, tb_hi2, tb_lo2, tg_hi, tg_lo, ty;
    int rc;
    if ((rc = make_tmap_out(&ty, const_cast<float*>(Y), M, N, N)) != CP_OK) return rc;        // load boxes: 32 fp32 x 32 rows
    if ((rc = make_tmap_2d(&ta_hi, A_hi, M, K, K, BM)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&ta_lo, A_lo, M, K, K, BM)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tb_hi2, B_hi, N, K, K, BN / 2)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tb_lo2, B_lo, N, K, K, BN / 2)) != CP_OK) return rc;
    if ((rc = make_tmap_plane_out(&tg_hi, G_hi, M, N, N)) != CP_OK) return rc;   // store boxes: 32 halves x 32 rows
    if ((rc = make_tmap_plane_out(&tg_lo, G_lo, M, N, N)) != CP_OK) return rc;
    if ((rc = set_pair_attrs()) != CP_OK) return rc;
    NtArgs g{nullptr, N, nullptr, pdb, nullptr, M, N, K, 0, out_scale, fast, nullptr, nullptr, 1.f, out_scale2};
    g.Y = Y; g.ldy = N;
    g.c1 = c1; g.c2 = c2; g.c3 = c3;
    g.bn_period = bn_period;
    g.gz_bound = gz_bound;
    g.g1max_out = g1max_out;
    g.gscale_inv_out = gscale_inv_out;
    const int64_t n_tiles = cp_cdiv(M, 2 * BM) * (N / BN);
    const int clusters = (int)(n_tiles < CP_NUM_SMS / 2 ? n_tiles : CP_NUM_SMS / 2);
    if (fast) gemm_tc_nt_pair_kernel<true, EPI_BNBWD><<<2 * clusters, pair::THREADS2, pair::SMEM2_BNBWD, st>>>(ta_hi, ta_lo, tb_hi2, tb_lo2, tg_hi, tg_lo, ty, g);
    else gemm_tc_nt_pair_kernel<false, EPI_BNBWD><<<2 * clusters, pair::THREADS2, pair::SMEM2_BNBWD, st>>>(ta_hi, ta_lo, tb_hi2, tb_lo2, tg_hi, tg_lo, ty, g);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// conv2 as implicit GEMM: C[(w,p), o] = act(sum_{tap,c} X[w, p+tap-1, c] * B[o, tap*64+c] + bias[o]);
// X planes are [windows][12][64], B planes [64][192]; partial statistics rows = ceil(windows/10)
inline int launch_conv_nt(const plane_t* X_hi, const plane_t* X_lo, int64_t windows, const plane_t* B_hi,
                          const plane_t* B_lo, const float* bias, float* C, float* psum, float* psq, int relu,
                          cudaStream_t st, const float* out_scale = nullptr, int fast = 0,
                          const float* out_scale2 = nullptr) {
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo, tc_out;
    int rc;
    if ((rc = make_tmap_out(&tc_out, C, windows * 12, 64, 64, CONV_ROWS)) != CP_OK) return rc;
    if ((rc = make_tmap_conv(&ta_hi, X_hi, windows, CONV_WIN)) != CP_OK) return rc;
    if ((rc = make_tmap_conv(&ta_lo, X_lo, windows, CONV_WIN)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tb_hi, B_hi, 64, 192, 192, 64)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tb_lo, B_lo, 64, 192, 192, 64)) != CP_OK) return rc;
    NtArgs g{C, 64, bias, psum, psq, windows * 12, 64, 192, relu, out_scale, fast, nullptr, nullptr, 1.f, out_scale2};
    return fast ? launch_nt_cfg<64, true, true>(ta_hi, ta_lo, tb_hi, tb_lo, tc_out, g, cp_cdiv(windows, CONV_WIN), st)
                : launch_nt_cfg<64, true, false>(ta_hi, ta_lo, tb_hi, tb_lo, tc_out, g, cp_cdiv(windows, CONV_WIN), st);
}

// conv2 weight gradient; P capacity >= splits*256*64 floats; *splits_out = number of slabs written
inline int launch_conv_tn(const plane_t* X_hi, const plane_t* X_lo, const plane_t* G_hi, const plane_t* G_lo,
                          int64_t windows, float* P, size_t p_capacity_elems, int* splits_out, cudaStream_t st,
                          int fast = 0) {
    CUtensorMap tx_hi, tx_lo, tg_hi, tg_lo;
    int rc;
    if ((rc = make_tmap_conv(&tx_hi, X_hi, windows, CW_WIN)) != CP_OK) return rc;
    if ((rc = make_tmap_conv(&tx_lo, X_lo, windows, CW_WIN)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tg_hi, G_hi, windows * 12, 64, 64, CW_ROWS)) != CP_OK) return rc;
    if ((rc = make_tmap_2d(&tg_lo, G_lo, windows * 12, 64, 64, CW_ROWS)) != CP_OK) return rc;
    CP_ONCE_PER_DEVICE({
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_tn_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM));
        CP_CUDA(cudaFuncSetAttribute(gemm_tc_tn_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CW_SMEM));
    });
    int S = CP_NUM_SMS / 2;
    const int64_t max_s = cp_cdiv(windows, (int64_t)CW_WIN * CW_CHUNK_KB);
    if (S > max_s) S = (int)max_s;
    const int64_t cap = (int64_t)(p_capacity_elems / (256 * 64));
    if (S > cap) S = (int)cap;
    if (S < 1) S = 1;
    const int64_t wps = cp_cdiv(cp_cdiv(windows, S), CW_WIN) * CW_WIN;
    S = (int)cp_cdiv(windows, wps);
    CwArgs g{P, windows, wps, fast};
    if (fast) gemm_tc_tn_conv_kernel<true><<<dim3(2, S), THREADS, CW_SMEM, st>>>(tx_hi, tx_lo, tg_hi, tg_lo, g);
    else gemm_tc_tn_conv_kernel<false><<<dim3(2, S), THREADS, CW_SMEM, st>>>(tx_hi, tx_lo, tg_hi, tg_lo, g);
    CP_CHECK_LAUNCH();
    *splits_out = S;
    return CP_OK;
}

}  // namespace tcg

// four consecutive elements -> 8-byte stores into the two planes (element index v*4)
__device__ __forceinline__ void split_store4(const float4& x, plane_t* hi, plane_t* lo, int64_t v) {
    plane_t h[4], l[4];
    split_f16(x.x, h[0], l[0]); split_f16(x.y, h[1], l[1]);
    split_f16(x.z, h[2], l[2]); split_f16(x.w, h[3], l[3]);
    reinterpret_cast<uint2*>(hi)[v] = *reinterpret_cast<const uint2*>(h);
    reinterpret_cast<uint2*>(lo)[v] = *reinterpret_cast<const uint2*>(l);
}

__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ x, plane_t* __restrict__ hi, plane_t* __restrict__ lo, int64_t n4,
                    float scale) {
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += (int64_t)gridDim.x * blockDim.x) {
        float4 a = __ldg(reinterpret_cast<const float4*>(x) + v);
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        split_store4(a, hi, lo, v);
    }
}
