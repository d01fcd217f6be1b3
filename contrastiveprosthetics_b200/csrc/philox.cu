// Test taps of the dropout mask generator (common.cuh): the raw Philox4x32-10 block function (known-answer vectors
// of Salmon et al., SC'11 / Random123 kat_vectors) and the keep mask exactly as the BatchNorm-apply kernels draw it
// (models.py:282-297 dropout; same seed / step / layer / element mapping as bn_apply_kernel and bn_apply_proj_kernel).
#include "common.cuh"

__global__ void philox_blocks_kernel(const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ key, int64_t n,
                                     uint32_t* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = philox4x32_10(make_uint4(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]),
                                  make_uint2(key[2 * i], key[2 * i + 1]));
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

__global__ void __launch_bounds__(256)
dropout_mask_kernel(uint8_t* __restrict__ keep, int64_t n4, float p, uint64_t seed, unsigned int layer,
                    const unsigned long long* __restrict__ step) {
    if (step) seed += __ldg(step) * 0x9E3779B97F4A7C15ull;
    const unsigned int thr = dropout_threshold(p);
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += (int64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uchar4*>(keep)[v] = dropout_keep4(seed, (uint64_t)v, layer, thr);
}

extern "C" int cp_philox4x32_10(const uint32_t* ctr, const uint32_t* key, int64_t n, uint32_t* out, void* stream) {
    if (n == 0) return CP_OK;
    if (!ctr || !key || !out || n < 0) return CP_ERR_ARG;
    philox_blocks_kernel<<<(unsigned)cp_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(ctr, key, n, out);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, int layer, const uint64_t* step,
                               void* stream) {
    if (n == 0) return CP_OK;
    if (!keep || n < 0 || n % 4 != 0 || p < 0.f || p >= 1.f || ((uintptr_t)keep) % 4 != 0) return CP_ERR_ARG;
    int64_t blocks = cp_cdiv(n / 4, 256);
    if (blocks > CP_NUM_SMS * 16) blocks = CP_NUM_SMS * 16;
    dropout_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(keep, n / 4, p, seed, (unsigned int)layer,
                                                                         (const unsigned long long*)step);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
