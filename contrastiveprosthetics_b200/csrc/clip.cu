// K3': batch x batch (CLIP) contrastive head -- BASELINE.json config 5.
//
// Generalises Model.forward's contrastive branch (models.py:112-130: L2-normalise both towers, inner
// products) from the per-group 41 x 41 `bmm` to ONE B x B similarity matrix with the CLIP loss the
// reference is "modeled after" (models.py:65): symmetric cross-entropy with arange targets,
//     S = s * Ehat . Ghat^T,   s = exp(logit_scale)  (models.py:81,129),
//     loss = 1/(2B) sum_i [ LSE_j S_ij - S_ii ] + 1/(2B) sum_j [ LSE_i S_ij - S_jj ].
// The B x B matrix (17 GB at B = 65,536) is never materialised: every kernel below recomputes 128 x 64
// tiles of it in registers.  Because both towers are L2-normalised, |S_ij| <= s, so exp(S_ij - s) can
// neither overflow nor needs a running maximum: the row and column sums are plain sums.
//
// One kernel, `clip_sweep_kernel`, serves all four passes; a CTA owns 64 "own" rows and sweeps over
// all "loop" rows:
//   sums:   own_sum[i] = sum_j exp(s*(x_i.y_j - 1))                          (+ first-max argmax)
//   grads:  d_own[i,:] = coef * sum_j exp(s*(x_i.y_j - 1)) * (1/own_sum[i] + 1/loop_sum[j]) * y_j
// row pass:    own = Ehat (this rank's rows), loop = Ghat (all B rows)
// column pass: own = Ghat (all B rows),       loop = Ehat (this rank's rows)  -> partial sums that the
//              host all-reduces (column sums) / reduce-scatters (d Ghat) across ranks.
// Results are deterministic (fixed-order shuffles, no float atomics).
#include <cstdlib>
#include <type_traits>
#include <cuda_fp16.h>
#include "common.cuh"

namespace {
constexpr int D = CP_EMB_DIM;
constexpr int OWN = 64;          // own rows per CTA  (16 row lanes x 4)
constexpr int LOOP = 128;        // loop rows per tile (16 column lanes x 8)
constexpr int THREADS = 256;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Sum v[0..15] over the 16 lanes that share an own-row group; lane l (of the 16) ends with the
// total of element l in v[0] (halving butterfly, fixed order).
__device__ __forceinline__ float reduce16_scatter(float (&v)[D], int l16) {
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        const bool up = (l16 & off) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
            const float send = up ? v[k] : v[k + off];
            const float keep = up ? v[k + off] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

struct SweepArgs {
    const float* X;          // (n_own, 16) normalised rows
    const float* Yt;         // (16, ld_y) normalised loop rows, k-major
    int64_t n_own, n_loop, ld_y;
    float a;                 // s * log2(e)
    // sums pass
    float* own_sum;          // (n_own)
    int32_t* own_arg;        // (n_own) first-max loop index, or null
    // gradient pass
    const float* own_sum_in; // (n_own)
    const float* loop_sum;   // (n_loop)
    float coef;
    float* d_own;            // (n_own, 16)
};

template <bool GRAD>
__global__ void __launch_bounds__(THREADS, 1) clip_sweep_kernel(const SweepArgs g) {
    __shared__ __align__(16) float Ys[2][D][LOOP];
    __shared__ __align__(16) float Ls[2][LOOP];          // 1 / loop_sum of the tile (GRAD)
    __shared__ __align__(16) float Xs[D][OWN];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int64_t i0 = (int64_t)blockIdx.x * OWN + ty * 4;

    // own rows, k-major in shared memory: Xs[k][row]  (one LDS.128 per k gives this thread's 4 rows)
    {
        const int row = tid / 4, q = tid % 4;
        const int64_t r = (int64_t)blockIdx.x * OWN + row;
        const float4 v = r < g.n_own ? __ldg(reinterpret_cast<const float4*>(g.X + r * D) + q)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        Xs[4 * q + 0][row] = v.x; Xs[4 * q + 1][row] = v.y; Xs[4 * q + 2][row] = v.z; Xs[4 * q + 3][row] = v.w;
    }
    float inv_own[4] = {0.f, 0.f, 0.f, 0.f};
    if (GRAD) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
            if (i0 + ii < g.n_own) inv_own[ii] = 1.f / __ldg(g.own_sum_in + i0 + ii);
    }

    const int64_t n_tiles = (g.n_loop + LOOP - 1) / LOOP;
    auto load_tile = [&](int64_t t, int buf) {
        const int64_t j0 = t * LOOP;
        // 16 x 128 floats = 512 16-byte chunks, two per thread
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int c = tid + q * THREADS, k = c / (LOOP / 4), jq = (c % (LOOP / 4)) * 4;
            float* dst = &Ys[buf][k][jq];
            if (j0 + jq < g.ld_y) cp_async16(dst, g.Yt + (int64_t)k * g.ld_y + j0 + jq);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (GRAD && tid < LOOP) {
            const int64_t j = j0 + tid;
            Ls[buf][tid] = j < g.n_loop ? 1.f / __ldg(g.loop_sum + j) : 0.f;
        }
        cp_async_commit();
    };

    float acc_sum[4] = {0.f, 0.f, 0.f, 0.f};
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int best_j[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    float2 d2[2][D];                                     // rows (0,1) and (2,3) of this thread, packed
    if (GRAD) {
#pragma unroll
        for (int ip = 0; ip < 2; ++ip)
#pragma unroll
            for (int k = 0; k < D; ++k) d2[ip][k] = make_float2(0.f, 0.f);
    }

    load_tile(0, 0);
    for (int64_t t = 0; t < n_tiles; ++t) {
        const int buf = (int)(t & 1);
        if (t + 1 < n_tiles) {
            load_tile(t + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        // similarities of this thread's 4 x 8 micro-tile, two own rows per packed FMA
        float2 s2[2][8];
#pragma unroll
        for (int ip = 0; ip < 2; ++ip)
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) s2[ip][jj] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float4 y0 = *reinterpret_cast<const float4*>(&Ys[buf][k][tx * 4]);
            const float4 y1 = *reinterpret_cast<const float4*>(&Ys[buf][k][64 + tx * 4]);
            const float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
            const float4 xv = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
            const float2 xp[2] = {make_float2(xv.x, xv.y), make_float2(xv.z, xv.w)};
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float2 yy = make_float2(y[jj], y[jj]);
                ffma2(s2[0][jj], xp[0], yy);
                ffma2(s2[1][jj], xp[1], yy);
            }
        }
        float s[4][8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            s[0][jj] = s2[0][jj].x; s[1][jj] = s2[0][jj].y; s[2][jj] = s2[1][jj].x; s[3][jj] = s2[1][jj].y;
        }
        // this thread's 8 loop columns: tx*4 .. tx*4+3 and 64 + tx*4 .. (two conflict-free 16-byte lanes per row)
        const int64_t jbase = t * LOOP + tx * 4;
#define CLIP_COL(jj) (jbase + ((jj) < 4 ? (jj) : 60 + (jj)))
        if (!GRAD) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const bool ok = CLIP_COL(jj) < g.n_loop;
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const float e = ok ? ex2(fmaf(s[ii][jj], g.a, -g.a)) : 0.f;
                    acc_sum[ii] += e;
                    if (ok && s[ii][jj] > best[ii]) { best[ii] = s[ii][jj]; best_j[ii] = (int)CLIP_COL(jj); }
                }
            }
        } else {
            const float4 l0 = *reinterpret_cast<const float4*>(&Ls[buf][tx * 4]);
            const float4 l1 = *reinterpret_cast<const float4*>(&Ls[buf][64 + tx * 4]);
            const float il[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const bool ok = CLIP_COL(jj) < g.n_loop;
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const float e = ok ? ex2(fmaf(s[ii][jj], g.a, -g.a)) : 0.f;
                    s[ii][jj] = e * (inv_own[ii] + il[jj]);
                }
            }
            // d[ii][k] += sum_jj c[ii][jj] * y[k][jj], two own rows per packed FMA
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float4 y0 = *reinterpret_cast<const float4*>(&Ys[buf][k][tx * 4]);
                const float4 y1 = *reinterpret_cast<const float4*>(&Ys[buf][k][64 + tx * 4]);
                const float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float2 yy = make_float2(y[jj], y[jj]);
                    ffma2(d2[0][k], make_float2(s[0][jj], s[1][jj]), yy);
                    ffma2(d2[1][k], make_float2(s[2][jj], s[3][jj]), yy);
                }
            }
        }
        __syncthreads();
    }

    // combine the 16 column lanes of each own row (fixed order)
    if (!GRAD) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            float v = acc_sum[ii];
            float b = best[ii];
            int bj = best_j[ii];
#pragma unroll
            for (int off = 8; off >= 1; off >>= 1) {
                v += __shfl_xor_sync(0xffffffffu, v, off);
                const float ob = __shfl_xor_sync(0xffffffffu, b, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                if (ob > b || (ob == b && oj < bj)) { b = ob; bj = oj; }
            }
            if (tx == 0 && i0 + ii < g.n_own) {
                g.own_sum[i0 + ii] = v;
                if (g.own_arg) g.own_arg[i0 + ii] = bj;
            }
        }
    } else {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            float dsum[D];
#pragma unroll
            for (int k = 0; k < D; ++k) dsum[k] = (ii & 1) ? d2[ii >> 1][k].y : d2[ii >> 1][k].x;
            const float tot = reduce16_scatter(dsum, tx);
            if (i0 + ii < g.n_own) g.d_own[(i0 + ii) * D + tx] = g.coef * tot;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// The same sweeps with the two contractions on the warp-level tensor cores (mma.sync m16n8k16, fp16 operands, fp32
// accumulate) and the 3-product split of the encoder's GEMMs: x = hi + lo/2048 with hi = fp16(x), lo = fp16((x - hi)*2048);
// a.b ~= hi_a.hi_b + (lo_a.hi_b + hi_a.lo_b)/2048 (22 significand bits per operand).  K = 16 is ONE k16 step: a 16-row x
// 8-column similarity fragment costs 3 MMAs instead of 128 FMAs, and the gradient contraction d_own += P . Y (P = the
// weights of the two fragments just computed, 16 x 16) takes the accumulator fragments AS its A operand -- the m16n8
// accumulator layout is the m16k16 A layout -- so P never leaves registers.  What remains per pair is the exponent
// (one ex2), a few FMAs and the hi/lo split of P.  (A first version on m16n8k8 tf32 MMAs, 3xTF32, was only 1.25x faster
// than the FFMA2 kernel: the legacy tf32 path issues at ~1/16 clk per sub-partition on this chip.)
// A warp owns 16 own rows for the whole sweep; a CTA = 8 warps = 128 own rows; the loop operand arrives in tiles of 128
// columns and is split into fp16 planes ONCE per CTA, in the two layouts the two contractions read: pairs along k
// (B operand of the similarity) and pairs along j (B operand of P . Y); padded rows make both fragment reads conflict-free.
// Accumulation chains are cut per tile (the tensor core's fp32 accumulator does not round to nearest): every tile's
// fragment is added to the running sums with ordinary fp32 adds.  Deterministic: fixed-order shuffles only.
// Problems of one tile (n_loop <= 128) stay on the FFMA2 kernel: there the similarity of a pair and the diagonal term of
// the loss kernel are the same fp32 FMA chain and cancel exactly (a batch of ONE has loss 0 to the last bit).
namespace mm {
constexpr int WARPS = 8;
constexpr int OWN = 16 * WARPS;      // own rows per CTA
constexpr int LOOP = 128;            // loop columns per tile
constexpr int KS = LOOP + 8;         // words per row of the k-pair layout  [8 k-pairs][LOOP columns]
constexpr int JS = LOOP / 2 + 4;     // words per row of the j-pair layout  [16 dims][LOOP/2 column pairs]
constexpr int THREADS = 32 * WARPS;
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    const __half2 h = __halves2half2(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void split_h(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn((x - __half2float(hi)) * LO_SCALE);
}
// Similarity operands (unit vectors, the own rows pre-multiplied by s*log2e): the lo plane is NOT scaled -- |x| <= ~4, so
// lo = x - hi is accurate to fp16's subnormal spacing, 6e-8 ABSOLUTE, which is what a dot product of 16 terms needs --
// and the three products share ONE accumulator (no combine step per element).  The weights P of the gradient contraction
// keep the scaled lo plane: they span many decades and are summed over the whole batch.
__device__ __forceinline__ void split_pair_abs(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}
// (a, b) -> packed hi pair, packed lo pair
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn((a - hf.x) * LO_SCALE, (b - hf.y) * LO_SCALE);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}
// D (16 x 8, fp32) += A (16 x 16 fp16, row fragment) . B (16 x 8 fp16, column fragment)
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool GRAD, bool ARG>
__global__ void __launch_bounds__(THREADS, 2) clip_sweep_mma_kernel(const SweepArgs g) {
    // loop tile, fp16 planes: Kh/Kl[kp][j] = (y[2kp][j], y[2kp+1][j]);  Jh/Jl[d][jp] = (y[d][2jp], y[d][2jp+1])
    __shared__ __align__(16) uint32_t Kh[2][D / 2][KS], Kl[2][D / 2][KS];
    __shared__ __align__(16) uint32_t Jh[2][D][JS], Jl[2][D][JS];
    __shared__ __align__(16) float Ls[2][LOOP];          // 1 / loop_sum of the tile (GRAD)
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int gq = lane >> 2, t = lane & 3;              // fragment coordinates: groupID, threadID_in_group
    const int64_t r0 = (int64_t)blockIdx.x * OWN + warp * 16;
    const int64_t row_a = r0 + gq, row_b = r0 + gq + 8;  // the two own rows of this thread's fragments

    // A operand of the similarity: the warp's 16 own rows x 16 dims (rows beyond n_own: zeros)
    //   a0: (row_a, dims 2t, 2t+1)  a1: (row_b, 2t, 2t+1)  a2: (row_a, 2t+8, 2t+9)  a3: (row_b, 2t+8, 2t+9)
    uint32_t ah[4], al[4];
    {
        const float2 z = make_float2(0.f, 0.f);
        const float2 xa0 = row_a < g.n_own ? __ldg(reinterpret_cast<const float2*>(g.X + row_a * D + 2 * t)) : z;
        const float2 xb0 = row_b < g.n_own ? __ldg(reinterpret_cast<const float2*>(g.X + row_b * D + 2 * t)) : z;
        const float2 xa1 = row_a < g.n_own ? __ldg(reinterpret_cast<const float2*>(g.X + row_a * D + 2 * t + 8)) : z;
        const float2 xb1 = row_b < g.n_own ? __ldg(reinterpret_cast<const float2*>(g.X + row_b * D + 2 * t + 8)) : z;
        split_pair_abs(xa0.x * g.a, xa0.y * g.a, ah[0], al[0]);       // pre-multiplied by s*log2(e): the MMAs return the
        split_pair_abs(xb0.x * g.a, xb0.y * g.a, ah[1], al[1]);       // exponent's argument up to the constant -s*log2(e)
        split_pair_abs(xa1.x * g.a, xa1.y * g.a, ah[2], al[2]);
        split_pair_abs(xb1.x * g.a, xb1.y * g.a, ah[3], al[3]);
    }
    float inv_own[2] = {0.f, 0.f};
    if (GRAD) {
        if (row_a < g.n_own) inv_own[0] = 1.f / __ldg(g.own_sum_in + row_a);
        if (row_b < g.n_own) inv_own[1] = 1.f / __ldg(g.own_sum_in + row_b);
    }

    const int64_t n_tiles = (g.n_loop + LOOP - 1) / LOOP;
    // this thread's share of a tile load: dims (2 kp, 2 kp + 1) x 4 consecutive columns
    const int kp = tid / 32;                             // 0..7
    const int ljq = (tid % 32) * 4;
    auto fetch = [&](int64_t tile, float4 (&v)[2], float& lsum) {
        const int64_t j0 = tile * LOOP;
#pragma unroll
        for (int q = 0; q < 2; ++q)
            v[q] = (j0 + ljq < g.ld_y) ? __ldg(reinterpret_cast<const float4*>(g.Yt + (int64_t)(2 * kp + q) * g.ld_y + j0 + ljq))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        if (GRAD && tid < LOOP) lsum = (j0 + tid < g.n_loop) ? __ldg(g.loop_sum + j0 + tid) : 0.f;
    };
    auto stash = [&](int buf, const float4 (&v)[2], float lsum) {
        const float e[2][4] = {{v[0].x, v[0].y, v[0].z, v[0].w}, {v[1].x, v[1].y, v[1].z, v[1].w}};
        __half h[2][4], l[2][4], la[2][4];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                split_h(e[q][c], h[q][c], l[q][c]);
                la[q][c] = __float2half_rn(e[q][c] - __half2float(h[q][c]));      // un-scaled lo (similarity operand)
            }
        // pairs along k: word (kp, j) = (dim 2kp, dim 2kp+1) of column j
        *reinterpret_cast<uint4*>(&Kh[buf][kp][ljq]) =
            make_uint4(pack_h2(h[0][0], h[1][0]), pack_h2(h[0][1], h[1][1]), pack_h2(h[0][2], h[1][2]), pack_h2(h[0][3], h[1][3]));
        *reinterpret_cast<uint4*>(&Kl[buf][kp][ljq]) =
            make_uint4(pack_h2(la[0][0], la[1][0]), pack_h2(la[0][1], la[1][1]), pack_h2(la[0][2], la[1][2]), pack_h2(la[0][3], la[1][3]));
        if (GRAD) {
            // pairs along j: word (d, jp) = columns (2jp, 2jp+1) of dim d
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                *reinterpret_cast<uint2*>(&Jh[buf][2 * kp + q][ljq / 2]) = make_uint2(pack_h2(h[q][0], h[q][1]), pack_h2(h[q][2], h[q][3]));
                *reinterpret_cast<uint2*>(&Jl[buf][2 * kp + q][ljq / 2]) = make_uint2(pack_h2(l[q][0], l[q][1]), pack_h2(l[q][2], l[q][3]));
            }
            if (tid < LOOP) Ls[buf][tid] = lsum > 0.f ? 1.f / lsum : 0.f;
        }
    };

    float acc[2] = {0.f, 0.f};                           // running row sums (rows gq, gq + 8)
    float best[2] = {-INFINITY, -INFINITY};
    int best_j[2] = {0x7fffffff, 0x7fffffff};
    float dsum[2][4];                                    // running gradient fragments: [d-tile][c0..c3]
#pragma unroll
    for (int dt = 0; dt < 2; ++dt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dsum[dt][e] = 0.f;

    {
        float4 v[2];
        float ls = 0.f;
        fetch(0, v, ls);
        stash(0, v, ls);
    }
    __syncthreads();
    for (int64_t tile = 0; tile < n_tiles; ++tile) {
        const int buf = (int)(tile & 1);
        float4 nv[2];
        float nls = 0.f;
        const bool more = tile + 1 < n_tiles;
        if (more) fetch(tile + 1, nv, nls);              // in flight while this tile is computed

        const int64_t j0 = tile * LOOP;
        float tacc[2] = {0.f, 0.f};
        float dm[2][4], dc[2][4];                        // this tile's gradient fragments: main / correction (x 2048) terms
        if (GRAD) {
#pragma unroll
            for (int dt = 0; dt < 2; ++dt)
#pragma unroll
                for (int e = 0; e < 4; ++e) dm[dt][e] = dc[dt][e] = 0.f;
        }
        auto sweep_tile = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;     // every column of the tile is a real loop row: no predicates
#pragma unroll 2
        for (int np = 0; np < LOOP / 16; ++np) {         // 16 columns = two 8-column similarity fragments
            float ev[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n0 = np * 16 + h * 8;
                const uint32_t bh0 = Kh[buf][t][n0 + gq], bh1 = Kh[buf][t + 4][n0 + gq];
                const uint32_t bl0 = Kl[buf][t][n0 + gq], bl1 = Kl[buf][t + 4][n0 + gq];
                float sv[4] = {-g.a, -g.a, -g.a, -g.a};                 // accumulator starts at -s*log2(e): sv = a * (x.y - 1)
                mma_f16(sv, al, bh0, bh1);
                mma_f16(sv, ah, bl0, bl1);
                mma_f16(sv, ah, bh0, bh1);
                // c0: (row_a, col 2t), c1: (row_a, 2t+1), c2: (row_b, 2t), c3: (row_b, 2t+1)
                const int64_t ja = j0 + n0 + 2 * t;
                const bool ok0 = FULL || ja < g.n_loop, ok1 = FULL || ja + 1 < g.n_loop;
                ev[h][0] = ok0 ? ex2(sv[0]) : 0.f;
                ev[h][1] = ok1 ? ex2(sv[1]) : 0.f;
                ev[h][2] = ok0 ? ex2(sv[2]) : 0.f;
                ev[h][3] = ok1 ? ex2(sv[3]) : 0.f;
                if (!GRAD) {
                    tacc[0] += ev[h][0] + ev[h][1];
                    tacc[1] += ev[h][2] + ev[h][3];
                    if (ARG) {
                    if (ok0 && sv[0] > best[0]) { best[0] = sv[0]; best_j[0] = (int)ja; }
                    if (ok1 && sv[1] > best[0]) { best[0] = sv[1]; best_j[0] = (int)ja + 1; }
                    if (ok0 && sv[2] > best[1]) { best[1] = sv[2]; best_j[1] = (int)ja; }
                    if (ok1 && sv[3] > best[1]) { best[1] = sv[3]; best_j[1] = (int)ja + 1; }
                    }
                } else {
                    const float2 il = *reinterpret_cast<const float2*>(&Ls[buf][n0 + 2 * t]);
                    ev[h][0] *= inv_own[0] + il.x; ev[h][1] *= inv_own[0] + il.y;
                    ev[h][2] *= inv_own[1] + il.x; ev[h][3] *= inv_own[1] + il.y;
                }
            }
            if (GRAD) {
                // P (16 rows x 16 columns) as the A operand: a0 = (row_a, cols 2t, 2t+1) of the first fragment, a1 = row_b,
                // a2 / a3 = the same of the second fragment (columns + 8)
                uint32_t ph[4], pl[4];
                split_pair(ev[0][0], ev[0][1], ph[0], pl[0]);
                split_pair(ev[0][2], ev[0][3], ph[1], pl[1]);
                split_pair(ev[1][0], ev[1][1], ph[2], pl[2]);
                split_pair(ev[1][2], ev[1][3], ph[3], pl[3]);
#pragma unroll
                for (int dt = 0; dt < 2; ++dt) {
                    // B: k = loop columns np*16 + (2t, 2t+1) and + 8, n = embedding dim dt*8 + gq
                    const uint32_t yh0 = Jh[buf][dt * 8 + gq][np * 8 + t], yh1 = Jh[buf][dt * 8 + gq][np * 8 + t + 4];
                    const uint32_t yl0 = Jl[buf][dt * 8 + gq][np * 8 + t], yl1 = Jl[buf][dt * 8 + gq][np * 8 + t + 4];
                    mma_f16(dc[dt], pl, yh0, yh1);
                    mma_f16(dc[dt], ph, yl0, yl1);
                    mma_f16(dm[dt], ph, yh0, yh1);
                }
            }
        }
        };
        if (j0 + LOOP <= g.n_loop) sweep_tile(std::true_type{}); else sweep_tile(std::false_type{});
        if (!GRAD) {
            acc[0] += tacc[0];
            acc[1] += tacc[1];
        } else {
#pragma unroll
            for (int dt = 0; dt < 2; ++dt)
#pragma unroll
                for (int e = 0; e < 4; ++e) dsum[dt][e] += fmaf(dc[dt][e], LO_INV, dm[dt][e]);
        }
        if (more) stash(buf ^ 1, nv, nls);               // the other buffer was last read before the previous barrier
        __syncthreads();
    }

    if (!GRAD) {
        // the 4 threads of a group hold disjoint columns of the same two rows
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float v = acc[h], b = best[h];
            int bj = best_j[h];
#pragma unroll
            for (int off = 1; off <= 2; off <<= 1) {
                v += __shfl_xor_sync(0xffffffffu, v, off);
                const float ob = __shfl_xor_sync(0xffffffffu, b, off);
                const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
                if (ob > b || (ob == b && oj < bj)) { b = ob; bj = oj; }
            }
            const int64_t row = h ? row_b : row_a;
            if (t == 0 && row < g.n_own) {
                g.own_sum[row] = v;
                if (ARG) g.own_arg[row] = bj;
            }
        }
    } else {
#pragma unroll
        for (int dt = 0; dt < 2; ++dt) {
            if (row_a < g.n_own)
                *reinterpret_cast<float2*>(g.d_own + row_a * D + dt * 8 + 2 * t) =
                    make_float2(g.coef * dsum[dt][0], g.coef * dsum[dt][1]);
            if (row_b < g.n_own)
                *reinterpret_cast<float2*>(g.d_own + row_b * D + dt * 8 + 2 * t) =
                    make_float2(g.coef * dsum[dt][2], g.coef * dsum[dt][3]);
        }
    }
}
// CP_CLIP_MMA=0 in the environment keeps every size on the FFMA2 kernel above (A/B runs)
inline bool use_mma(int64_t n_loop) {
    static const bool on = [] { const char* e = getenv("CP_CLIP_MMA"); return !(e && e[0] == '0'); }();
    return on && n_loop > LOOP;
}
}  // namespace mm

// xhat = x / ||x||, inv_norm = 1 / ||x||    (models.py:123,125: no epsilon)
__global__ void __launch_bounds__(256)
clip_normalize_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ xhat, float* __restrict__ inv_norm) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    float4 v[D / 4];
    float n2 = 0.f;
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
        v[q] = __ldg(reinterpret_cast<const float4*>(x + r * D) + q);
        n2 = fmaf(v[q].x, v[q].x, n2); n2 = fmaf(v[q].y, v[q].y, n2);
        n2 = fmaf(v[q].z, v[q].z, n2); n2 = fmaf(v[q].w, v[q].w, n2);
    }
    const float inv = 1.f / sqrtf(n2);
#pragma unroll
    for (int q = 0; q < D / 4; ++q)
        reinterpret_cast<float4*>(xhat + r * D)[q] = make_float4(v[q].x * inv, v[q].y * inv, v[q].z * inv, v[q].w * inv);
    inv_norm[r] = inv;
}

// (n,16) row-major -> (16, ld) k-major
__global__ void __launch_bounds__(256)
clip_transpose_kernel(const float* __restrict__ xhat, int64_t n, int64_t ld, float* __restrict__ xt) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= ld) return;
#pragma unroll
    for (int k = 0; k < D; ++k) xt[(int64_t)k * ld + r] = r < n ? __ldg(xhat + r * D + k) : 0.f;
}

// partial loss over the n samples whose row AND column this rank owns:
//   sum_i [ log rowsum_i + log colsum_i + 2 s - 2 s ehat_i . ghat_i ] / (2 B)       (one CTA, double)
__global__ void __launch_bounds__(1024)
clip_loss_kernel(const float* __restrict__ ehat, const float* __restrict__ ghat, const float* __restrict__ rowsum,
                 const float* __restrict__ colsum, int64_t n, int64_t B, float s, const int32_t* __restrict__ row_arg,
                 int64_t row0, float* __restrict__ loss, int32_t* __restrict__ n_correct) {
    __shared__ double red[32];
    __shared__ int redc[32];
    double acc = 0.0;
    int cor = 0;
    // one double-precision log per TWO samples: log r1 + log c1 + log r2 + log c2 = log(r1 c1 r2 c2) (the sums lie in
    // (exp(-2s), B]: four factors stay far inside the double range); the fp64 log was 90 % of this kernel's 0.54 ms
    double prod = 1.0;
    int in_prod = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < D / 4; ++q) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(ehat + i * D) + q);
            const float4 c = __ldg(reinterpret_cast<const float4*>(ghat + i * D) + q);
            dot = fmaf(a.x, c.x, dot); dot = fmaf(a.y, c.y, dot); dot = fmaf(a.z, c.z, dot); dot = fmaf(a.w, c.w, dot);
        }
        prod *= (double)rowsum[i] * (double)colsum[i];
        if (++in_prod == 2) { acc += log(prod); prod = 1.0; in_prod = 0; }
        acc += 2.0 * (double)s - 2.0 * (double)s * (double)dot;
        if (row_arg) cor += (int64_t)row_arg[i] == row0 + i;
    }
    if (in_prod) acc += log(prod);
    acc = warp_sum(acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cor += __shfl_xor_sync(0xffffffffu, cor, o);
    if (threadIdx.x % 32 == 0) { red[threadIdx.x / 32] = acc; redc[threadIdx.x / 32] = cor; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        int c = 0;
        for (int w = 0; w < 32; ++w) { t += red[w]; c += redc[w]; }
        *loss = (float)(t / (2.0 * (double)B));
        if (n_correct) *n_correct = c;
    }
}

// dhat = d_hat - diag_coef * other;  dx = (dhat - xhat (xhat . dhat)) * inv_norm
__global__ void __launch_bounds__(256)
clip_embed_bwd_kernel(const float* __restrict__ d_hat, const float* __restrict__ xhat, const float* __restrict__ other,
                      const float* __restrict__ inv_norm, int64_t n, float diag_coef, float* __restrict__ dx) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    float dh[D], xh[D];
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(d_hat + r * D) + q);
        const float4 o = __ldg(reinterpret_cast<const float4*>(other + r * D) + q);
        const float4 h = __ldg(reinterpret_cast<const float4*>(xhat + r * D) + q);
        dh[4 * q] = a.x - diag_coef * o.x; dh[4 * q + 1] = a.y - diag_coef * o.y;
        dh[4 * q + 2] = a.z - diag_coef * o.z; dh[4 * q + 3] = a.w - diag_coef * o.w;
        xh[4 * q] = h.x; xh[4 * q + 1] = h.y; xh[4 * q + 2] = h.z; xh[4 * q + 3] = h.w;
    }
#pragma unroll
    for (int k = 0; k < D; ++k) dot = fmaf(xh[k], dh[k], dot);
    const float inv = __ldg(inv_norm + r);
#pragma unroll
    for (int q = 0; q < D / 4; ++q)
        reinterpret_cast<float4*>(dx + r * D)[q] =
            make_float4((dh[4 * q] - xh[4 * q] * dot) * inv, (dh[4 * q + 1] - xh[4 * q + 1] * dot) * inv,
                        (dh[4 * q + 2] - xh[4 * q + 2] * dot) * inv, (dh[4 * q + 3] - xh[4 * q + 3] * dot) * inv);
}

inline unsigned row_grid(int64_t n) { return (unsigned)cp_cdiv(n, 256); }
}  // namespace

extern "C" int cp_clip_normalize(const float* x, int64_t n, float* xhat, float* inv_norm, void* stream) {
    if (!x || !xhat || !inv_norm || n <= 0) return CP_ERR_ARG;
    clip_normalize_kernel<<<row_grid(n), 256, 0, (cudaStream_t)stream>>>(x, n, xhat, inv_norm);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_clip_transpose(const float* xhat, int64_t n, int64_t ld, float* xhat_t, void* stream) {
    if (!xhat || !xhat_t || n <= 0 || ld < n || ld % 4 != 0) return CP_ERR_ARG;
    clip_transpose_kernel<<<row_grid(ld), 256, 0, (cudaStream_t)stream>>>(xhat, n, ld, xhat_t);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_clip_sums(const float* own, int64_t n_own, const float* loop_t, int64_t n_loop, int64_t ld_loop,
                            float scale, float* own_sum, int32_t* own_argmax, void* stream) {
    if (!own || !loop_t || !own_sum || n_own <= 0 || n_loop <= 0 || ld_loop < n_loop || ld_loop % 4 != 0 ||
        !(scale > 0.f) || ((uintptr_t)loop_t % 16) != 0)
        return CP_ERR_ARG;
    SweepArgs g{own, loop_t, n_own, n_loop, ld_loop, scale * LOG2E, own_sum, own_argmax, nullptr, nullptr, 0.f, nullptr};
    if (mm::use_mma(n_loop))
        if (own_argmax) mm::clip_sweep_mma_kernel<false, true><<<(unsigned)cp_cdiv(n_own, mm::OWN), mm::THREADS, 0, (cudaStream_t)stream>>>(g);
        else mm::clip_sweep_mma_kernel<false, false><<<(unsigned)cp_cdiv(n_own, mm::OWN), mm::THREADS, 0, (cudaStream_t)stream>>>(g);
    else
    clip_sweep_kernel<false><<<(unsigned)cp_cdiv(n_own, OWN), THREADS, 0, (cudaStream_t)stream>>>(g);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_clip_grad(const float* own, int64_t n_own, const float* loop_t, int64_t n_loop, int64_t ld_loop,
                            float scale, const float* own_sum, const float* loop_sum, float coef, float* d_own,
                            void* stream) {
    if (!own || !loop_t || !own_sum || !loop_sum || !d_own || n_own <= 0 || n_loop <= 0 || ld_loop < n_loop ||
        ld_loop % 4 != 0 || !(scale > 0.f) || ((uintptr_t)loop_t % 16) != 0)
        return CP_ERR_ARG;
    SweepArgs g{own, loop_t, n_own, n_loop, ld_loop, scale * LOG2E, nullptr, nullptr, own_sum, loop_sum, coef, d_own};
    if (mm::use_mma(n_loop))
        mm::clip_sweep_mma_kernel<true, false><<<(unsigned)cp_cdiv(n_own, mm::OWN), mm::THREADS, 0, (cudaStream_t)stream>>>(g);
    else
    clip_sweep_kernel<true><<<(unsigned)cp_cdiv(n_own, OWN), THREADS, 0, (cudaStream_t)stream>>>(g);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_clip_loss(const float* ehat, const float* ghat, const float* rowsum, const float* colsum, int64_t n,
                            int64_t B, float scale, const int32_t* row_argmax, int64_t row0, float* loss,
                            int32_t* n_correct, void* stream) {
    if (!ehat || !ghat || !rowsum || !colsum || !loss || n <= 0 || B < n || !(scale > 0.f)) return CP_ERR_ARG;
    clip_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ehat, ghat, rowsum, colsum, n, B, scale, row_argmax, row0, loss,
                                                          n_correct);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_clip_embed_backward(const float* d_hat, const float* xhat, const float* other_hat,
                                      const float* inv_norm, int64_t n, float diag_coef, float* dx, void* stream) {
    if (!d_hat || !xhat || !other_hat || !inv_norm || !dx || n <= 0) return CP_ERR_ARG;
    clip_embed_bwd_kernel<<<row_grid(n), 256, 0, (cudaStream_t)stream>>>(d_hat, xhat, other_hat, inv_norm, n, diag_coef, dx);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
