// K1: DB23 batch gather + emg_mean/emg_std normalisation.
// Replaces utils.py:51-64 / load.py:256-273 (one advanced-index launch per ITEM plus a
// default_collate stack) by one coalesced launch per batch, and fuses utils.py:129-130.
// HBM-bound: 8 B index + row read + row write per row; 16-byte vector lanes, rows are 48 B
// (train) or 1200 B (eval) so every row is an integral number of float4.
#include "common.cuh"

// IdxT: 32-bit element indices when the batch allows it (64-bit division is emulated on the GPU).  UNROLL
// independent rows are in flight per thread: index load -> row load is a dependent chain of two HBM latencies,
// and random 48-byte rows give the memory system nothing to prefetch.
#define GATHER_UNROLL 4
template <bool VEC4, typename IdxT>
__global__ void __launch_bounds__(256)
gather_norm_kernel(const float* __restrict__ src, int64_t src_rows, int row_len,
                   const int64_t* __restrict__ idx, int64_t n_rows, float* __restrict__ dst,
                   const float* __restrict__ mean, const float* __restrict__ stdv, int stat_len,
                   int n_ch, int* __restrict__ err_flag) {
    const IdxT vpr = (IdxT)(VEC4 ? row_len / 4 : row_len);   // vectors per row
    const IdxT total = (IdxT)n_rows * vpr;
    const IdxT stride = (IdxT)gridDim.x * blockDim.x;
    // per-channel statistics as float4 when a vector never straddles the channel period
    const bool stat4 = VEC4 && stat_len > 1 && n_ch % 4 == 0;
    for (IdxT v0 = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; v0 < total; v0 += GATHER_UNROLL * stride) {
        IdxT r[GATHER_UNROLL];
        int c0[GATHER_UNROLL];
        int64_t s[GATHER_UNROLL];
#pragma unroll
        for (int u = 0; u < GATHER_UNROLL; ++u) {
            const IdxT v = v0 + (IdxT)u * stride;
            r[u] = v / vpr;
            c0[u] = (int)(v - r[u] * vpr) * (VEC4 ? 4 : 1);
            s[u] = v < total ? __ldg(idx + r[u]) : 0;
            if (s[u] < 0 || s[u] >= src_rows) {
                if (err_flag) *err_flag = 1;
                s[u] = 0;
            }
        }
        if (VEC4) {
            float4 x[GATHER_UNROLL];
#pragma unroll
            for (int u = 0; u < GATHER_UNROLL; ++u)
                x[u] = __ldg(reinterpret_cast<const float4*>(src + s[u] * row_len + c0[u]));
#pragma unroll
            for (int u = 0; u < GATHER_UNROLL; ++u) {
                if (v0 + (IdxT)u * stride >= total) break;
                if (stat_len > 0) {
                    float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
                    if (stat4) {
                        const int ch = c0[u] % n_ch;
                        const float4 m = __ldg(reinterpret_cast<const float4*>(mean + ch));
                        const float4 d = __ldg(reinterpret_cast<const float4*>(stdv + ch));
                        xs[0] = __fdiv_rn(xs[0] - m.x, d.x); xs[1] = __fdiv_rn(xs[1] - m.y, d.y);   // true divide
                        xs[2] = __fdiv_rn(xs[2] - m.z, d.z); xs[3] = __fdiv_rn(xs[3] - m.w, d.w);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ch = stat_len == 1 ? 0 : (c0[u] + j) % n_ch;
                            xs[j] = __fdiv_rn(xs[j] - __ldg(mean + ch), __ldg(stdv + ch));
                        }
                    }
                    x[u] = make_float4(xs[0], xs[1], xs[2], xs[3]);
                }
                *reinterpret_cast<float4*>(dst + (int64_t)r[u] * row_len + c0[u]) = x[u];
            }
        } else {
#pragma unroll
            for (int u = 0; u < GATHER_UNROLL; ++u) {
                if (v0 + (IdxT)u * stride >= total) break;
                float x = __ldg(src + s[u] * row_len + c0[u]);
                if (stat_len > 0) {
                    const int ch = stat_len == 1 ? 0 : c0[u] % n_ch;
                    x = __fdiv_rn(x - __ldg(mean + ch), __ldg(stdv + ch));
                }
                dst[(int64_t)r[u] * row_len + c0[u]] = x;
            }
        }
    }
}

template <bool VEC4, typename IdxT>
static int launch_gather(unsigned blocks, cudaStream_t st, const float* src, int64_t src_rows, int row_len,
                         const int64_t* idx, int64_t n_rows, float* dst, const float* mean, const float* stdv,
                         int stat_len, int n_ch, int* err_flag) {
    gather_norm_kernel<VEC4, IdxT><<<blocks, 256, 0, st>>>(src, src_rows, row_len, idx, n_rows, dst, mean, stdv,
                                                          stat_len, n_ch, err_flag);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_gather_norm(const float* src, int64_t src_rows, int row_len, const int64_t* idx,
                              int64_t n_rows, float* dst, const float* mean, const float* stdv,
                              int stat_len, int n_ch, int* err_flag, void* stream) {
    if (n_rows == 0) return CP_OK;
    if (!src || !idx || !dst || row_len <= 0 || src_rows <= 0 || n_rows < 0) return CP_ERR_ARG;
    if (stat_len != 0 && (!mean || !stdv || n_ch <= 0 || (stat_len != 1 && stat_len != n_ch)))
        return CP_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (row_len % 4 == 0) && (((uintptr_t)src | (uintptr_t)dst) % 16 == 0) &&
                     (stat_len <= 1 || n_ch % 4 != 0 || (((uintptr_t)mean | (uintptr_t)stdv) % 16 == 0));
    const int64_t total = n_rows * (vec ? row_len / 4 : row_len);
    int64_t blocks = cp_cdiv(total, 256 * GATHER_UNROLL);
    const int64_t cap = (int64_t)CP_NUM_SMS * 16;            // 8 resident CTAs/SM x 2 waves
    if (blocks > cap) blocks = cap;
    // 32-bit element indices unless the batch (plus the unroll overshoot) could wrap them
    const bool small = total + (int64_t)GATHER_UNROLL * cap * 256 < (int64_t)1 << 31;
    const unsigned g = (unsigned)blocks;
    if (vec)
        return small ? launch_gather<true, int32_t>(g, st, src, src_rows, row_len, idx, n_rows, dst, mean, stdv, stat_len, n_ch, err_flag)
                     : launch_gather<true, int64_t>(g, st, src, src_rows, row_len, idx, n_rows, dst, mean, stdv, stat_len, n_ch, err_flag);
    return small ? launch_gather<false, int32_t>(g, st, src, src_rows, row_len, idx, n_rows, dst, mean, stdv, stat_len, n_ch, err_flag)
                 : launch_gather<false, int64_t>(g, st, src, src_rows, row_len, idx, n_rows, dst, mean, stdv, stat_len, n_ch, err_flag);
}
