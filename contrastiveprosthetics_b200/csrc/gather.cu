// K1: DB23 batch gather + emg_mean/emg_std normalisation.
// Replaces utils.py:51-64 / load.py:256-273 (one advanced-index launch per ITEM plus a
// default_collate stack) by one coalesced launch per batch, and fuses utils.py:129-130.
// HBM-bound: 8 B index + row read + row write per row; 16-byte vector lanes, rows are 48 B
// (train) or 1200 B (eval) so every row is an integral number of float4.
#include "common.cuh"

template <bool VEC4>
__global__ void __launch_bounds__(256)
gather_norm_kernel(const float* __restrict__ src, int64_t src_rows, int row_len,
                   const int64_t* __restrict__ idx, int64_t n_rows, float* __restrict__ dst,
                   const float* __restrict__ mean, const float* __restrict__ stdv, int stat_len,
                   int n_ch, int* __restrict__ err_flag) {
    const int vpr = VEC4 ? row_len / 4 : row_len;            // vectors per row
    const int64_t total = n_rows * vpr;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total;
         v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = v / vpr;
        const int c0 = (int)(v - r * vpr) * (VEC4 ? 4 : 1);
        int64_t s = __ldg(idx + r);
        if (s < 0 || s >= src_rows) {
            if (err_flag) *err_flag = 1;
            s = 0;
        }
        if (VEC4) {
            float4 x = __ldg(reinterpret_cast<const float4*>(src + s * row_len + c0));
            if (stat_len > 0) {
                float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ch = stat_len == 1 ? 0 : (c0 + j) % n_ch;
                    xs[j] = __fdiv_rn(xs[j] - __ldg(mean + ch), __ldg(stdv + ch));   // true divide
                }
                x = make_float4(xs[0], xs[1], xs[2], xs[3]);
            }
            *reinterpret_cast<float4*>(dst + r * row_len + c0) = x;
        } else {
            float x = __ldg(src + s * row_len + c0);
            if (stat_len > 0) {
                const int ch = stat_len == 1 ? 0 : c0 % n_ch;
                x = __fdiv_rn(x - __ldg(mean + ch), __ldg(stdv + ch));
            }
            dst[r * row_len + c0] = x;
        }
    }
}

extern "C" int cp_gather_norm(const float* src, int64_t src_rows, int row_len, const int64_t* idx,
                              int64_t n_rows, float* dst, const float* mean, const float* stdv,
                              int stat_len, int n_ch, int* err_flag, void* stream) {
    if (n_rows == 0) return CP_OK;
    if (!src || !idx || !dst || row_len <= 0 || src_rows <= 0 || n_rows < 0) return CP_ERR_ARG;
    if (stat_len != 0 && (!mean || !stdv || n_ch <= 0 || (stat_len != 1 && stat_len != n_ch)))
        return CP_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (row_len % 4 == 0) && (((uintptr_t)src | (uintptr_t)dst) % 16 == 0);
    const int64_t total = n_rows * (vec ? row_len / 4 : row_len);
    int64_t blocks = cp_cdiv(total, 256);
    const int64_t cap = (int64_t)CP_NUM_SMS * 16;            // 8 resident CTAs/SM x 2 waves
    if (blocks > cap) blocks = cap;
    if (vec)
        gather_norm_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(src, src_rows, row_len, idx, n_rows,
                                                                  dst, mean, stdv, stat_len, n_ch, err_flag);
    else
        gather_norm_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(src, src_rows, row_len, idx, n_rows,
                                                                   dst, mean, stdv, stat_len, n_ch, err_flag);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
