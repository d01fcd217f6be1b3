// K5: the regulariser of Model.l2 / EMGNet.l2 / GLOVENet.l2 (models.py:225-228, 344-349, 467-472):
// sum over a list of parameter tensors of the UN-squared Frobenius norm.  The reference issues one
// torch.norm + one add per tensor forward and ~5 elementwise kernels per tensor backward (~100 launches of a
// few microseconds per step); here: two launches forward, one backward, for the whole list.
//   forward : norms[t] = ||W_t||_2 (double accumulation, fixed reduction order -> deterministic),
//             *total   = sum_t norms[t]  (float, tensors in list order like the reference's running sum)
//   backward: dW_t     = coef * W_t / norms[t]      (0 where the norm is 0: torch's masked_fill of norm's backward)
#include "common.cuh"

#define L2_MAX_TENSORS 32
#define L2_CHUNKS 16                                   // CTAs per tensor
struct L2List {
    const float* w[L2_MAX_TENSORS];
    float* g[L2_MAX_TENSORS];
    int64_t n[L2_MAX_TENSORS];
    int count;
};

__global__ void __launch_bounds__(256)
l2_partial_kernel(const L2List L, double* __restrict__ partial /*[count][L2_CHUNKS]*/) {
    __shared__ double red[8];
    const int t = blockIdx.y, c = blockIdx.x;
    const float* __restrict__ w = L.w[t];
    const int64_t n = L.n[t];
    const int64_t per = (n + L2_CHUNKS - 1) / L2_CHUNKS;
    const int64_t lo = c * per, hi = min(n, lo + per);
    const double tot = cta_sum_squares_256(w, lo, hi, red);
    if (threadIdx.x == 0) partial[t * L2_CHUNKS + c] = tot;
}

__global__ void l2_finish_kernel(const double* __restrict__ partial, int count, float* __restrict__ norms,
                                 float* __restrict__ total) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float acc = 0.f;
    for (int t = 0; t < count; ++t) {
        double s = 0.0;
        for (int c = 0; c < L2_CHUNKS; ++c) s += partial[t * L2_CHUNKS + c];
        const float nt = (float)sqrt(s);
        norms[t] = nt;
        acc += nt;                                      // float running sum in list order (models.py:347-348)
    }
    *total = acc;
}

__global__ void __launch_bounds__(256)
l2_backward_kernel(const L2List L, const float* __restrict__ norms, const float* __restrict__ g_total, float coef) {
    const int t = blockIdx.y;
    const float* __restrict__ w = L.w[t];
    float* __restrict__ g = L.g[t];
    const int64_t n = L.n[t];
    const float nt = __ldg(norms + t);
    const float k = nt > 0.f ? coef * __ldg(g_total) / nt : 0.f;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) g[i] = k * __ldg(w + i);
}

static int fill(L2List& L, const float* const* tensors, float* const* grads, const int64_t* sizes, int count) {
    if (!tensors || !sizes || count <= 0 || count > L2_MAX_TENSORS) return CP_ERR_ARG;
    L.count = count;
    for (int t = 0; t < count; ++t) {
        if (!tensors[t] || sizes[t] <= 0 || (grads && !grads[t])) return CP_ERR_ARG;
        L.w[t] = tensors[t];
        L.g[t] = grads ? grads[t] : nullptr;
        L.n[t] = sizes[t];
    }
    return CP_OK;
}

extern "C" size_t cp_l2_workspace_bytes(int n_tensors) {
    if (n_tensors <= 0 || n_tensors > L2_MAX_TENSORS) return 0;
    return sizeof(double) * n_tensors * L2_CHUNKS;
}

extern "C" int cp_l2_forward(const float* const* tensors, const int64_t* sizes, int n_tensors, float* norms,
                             float* total, void* workspace, size_t workspace_bytes, void* stream) {
    L2List L;
    if (int rc = fill(L, tensors, nullptr, sizes, n_tensors)) return rc;
    if (!norms || !total || !workspace || workspace_bytes < cp_l2_workspace_bytes(n_tensors)) return CP_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    double* partial = reinterpret_cast<double*>(workspace);
    l2_partial_kernel<<<dim3(L2_CHUNKS, n_tensors), 256, 0, st>>>(L, partial);
    CP_CHECK_LAUNCH();
    l2_finish_kernel<<<1, 32, 0, st>>>(partial, n_tensors, norms, total);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_l2_backward(const float* const* tensors, const int64_t* sizes, int n_tensors, const float* norms,
                              const float* g_total, float coef, float* const* grads, void* stream) {
    L2List L;
    if (int rc = fill(L, tensors, grads, sizes, n_tensors)) return rc;
    if (!norms || !g_total) return CP_ERR_ARG;
    l2_backward_kernel<<<dim3(L2_CHUNKS * 2, n_tensors), 256, 0, (cudaStream_t)stream>>>(L, norms, g_total, coef);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
