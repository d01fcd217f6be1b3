// K6: the two ends of the train step that the reference leaves to torch's autograd / optimiser machinery
// (train.py:95-108: `loss = model.loss(...) + model.l2()`, `loss.backward()`, `optimizer_emg.step()`,
// `optimizer_glove.step()`, with the two Adams of train.py:72-73).  Through autograd those cost ~35 launches of a few
// microseconds per step (norm / mul / add / fill kernels, one gradient accumulation per regularised tensor, five
// multi-tensor Adam launches) -- a quarter of the nodes of the batch_size-8 step (go.sh:6).  Here:
//   cp_step_prologue : ONE launch: ||W_t||_2 of every regularised tensor (same arithmetic as cp_l2_forward: double
//                      partial sums in a fixed order, last-arriving CTA finishes) + the step counters advance
//                      (dropout key, Adam's t).
//   cp_adam_step     : ONE launch for every parameter of BOTH optimisers: g = dL/dW (+ reg * W / ||W||, the gradient
//                      of Model.l2, models.py:225-228, 344-349) -> Adam (torch.optim.Adam defaults: no weight decay,
//                      no amsgrad; the arithmetic of torch's single-kernel implementation: moments in double
//                      precision from fp32 state, bias corrections from pow() in double).
// Gradients, exp_avg and exp_avg_sq are FLAT buffers (tensor t at offsets[t]), so a sample-sharded job all-reduces
// the gradient bucket in place, with no pack / unpack.
#include "common.cuh"

#define ST_MAX CP_STEP_MAX_TENSORS
#define ST_CHUNKS 16                 // CTAs per tensor of the norm pass (== L2_CHUNKS of l2.cu: same partial sums)
#define ST_BLOCK_ELEMS 1024          // parameters per CTA of the Adam pass

struct NormList {
    const float* w[ST_MAX];
    int64_t n[ST_MAX];
    int count;
};

__global__ void __launch_bounds__(256)
step_prologue_kernel(const NormList L, double* __restrict__ partial, unsigned int* __restrict__ ticket,
                     float* __restrict__ norms, int64_t* __restrict__ counters, int n_counters) {
    __shared__ double red[8];
    __shared__ bool last;
    const int t = blockIdx.y, c = blockIdx.x;
    const float* __restrict__ w = L.w[t];
    const int64_t n = L.n[t];
    const int64_t per = (n + ST_CHUNKS - 1) / ST_CHUNKS;
    const int64_t lo = c * per, hi = min(n, lo + per);
    const double tot = cta_sum_squares_256(w, lo, hi, red);
    if (threadIdx.x == 0) {
        partial[t * ST_CHUNKS + c] = tot;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int k = threadIdx.x; k < L.count; k += 256) {
        double a = 0.0;
        for (int j = 0; j < ST_CHUNKS; ++j) a += __ldcg(partial + k * ST_CHUNKS + j);
        norms[k] = (float)sqrt(a);
    }
    if (threadIdx.x == 0) *ticket = 0;                     // ready for the next launch (graph replays included)
    if (threadIdx.x < n_counters) counters[threadIdx.x] += 1;
}

extern "C" size_t cp_step_workspace_bytes(int n_tensors) {
    if (n_tensors <= 0 || n_tensors > ST_MAX) return 0;
    return 256 + sizeof(double) * n_tensors * ST_CHUNKS;
}

extern "C" int cp_step_prologue(const float* const* tensors, const int64_t* sizes, int n_tensors, float* norms,
                                int64_t* counters, int n_counters, void* workspace, size_t workspace_bytes,
                                void* stream) {
    if (!tensors || !sizes || n_tensors <= 0 || n_tensors > ST_MAX || !norms || !workspace ||
        workspace_bytes < cp_step_workspace_bytes(n_tensors) || n_counters < 0 || n_counters > 32 ||
        (n_counters > 0 && !counters))
        return CP_ERR_ARG;
    NormList L;
    L.count = n_tensors;
    for (int t = 0; t < n_tensors; ++t) {
        if (!tensors[t] || sizes[t] <= 0) return CP_ERR_ARG;
        L.w[t] = tensors[t];
        L.n[t] = sizes[t];
    }
    unsigned int* ticket = reinterpret_cast<unsigned int*>(workspace);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256);
    step_prologue_kernel<<<dim3(ST_CHUNKS, n_tensors), 256, 0, (cudaStream_t)stream>>>(L, partial, ticket, norms, counters,
                                                                                     n_counters);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

struct AdamList {
    float* p[ST_MAX];
    int64_t off[ST_MAX];
    int64_t n[ST_MAX];
    int blk0[ST_MAX + 1];
    int lr_index[ST_MAX];
    int norm_index[ST_MAX];
    float reg[ST_MAX];
    int count;
};

struct AdamCoef {
    double beta1, beta2, eps;
    float step_size, bc2_sqrt, k;
    bool regularised;
};
// one parameter: the regulariser's gradient with its own two roundings (what autograd adds to .grad), then Adam in the
// arithmetic of torch's single-kernel implementation (moments through double, the rest in float)
__device__ __forceinline__ void adam_update(const AdamCoef& c, float& w, float g, float& m, float& v) {
    if (c.regularised) g = __fadd_rn(g, __fmul_rn(c.k, w));
    m = (float)(c.beta1 * (double)m + (1.0 - c.beta1) * (double)g);
    v = (float)(c.beta2 * (double)v + (1.0 - c.beta2) * (double)g * (double)g);
    const float denom = (float)((double)(sqrtf(v) / c.bc2_sqrt) + c.eps);
    w -= c.step_size * m / denom;
}

__global__ void __launch_bounds__(256)
adam_step_kernel(const AdamList L, const float* __restrict__ grads, float* __restrict__ exp_avg,
                 float* __restrict__ exp_avg_sq, const double* __restrict__ lr, const float* __restrict__ norms,
                 const int64_t* __restrict__ step, const double beta1, const double beta2, const double eps) {
    __shared__ int s_t;
    __shared__ float s_step_size, s_bc2_sqrt, s_k;
    if (threadIdx.x == 0) {
        int t = 0;
        while (t + 1 < L.count && (int)blockIdx.x >= L.blk0[t + 1]) ++t;
        s_t = t;
        const double st = (double)*step;
        const float bc1 = (float)(1.0 - pow(beta1, st));
        s_bc2_sqrt = (float)sqrt(1.0 - pow(beta2, st));
        s_step_size = (float)(lr[L.lr_index[t]] / (double)bc1);
        float k = 0.f;
        if (L.norm_index[t] >= 0) {                        // d(reg * ||W||)/dW = reg * W / ||W||  (0 where the norm is 0)
            const float nt = norms[L.norm_index[t]];
            k = nt > 0.f ? L.reg[t] / nt : 0.f;
        }
        s_k = k;
    }
    __syncthreads();
    const int t = s_t;
    const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt, k = s_k;
    const bool regularised = L.norm_index[t] >= 0;
    float* __restrict__ p = L.p[t];
    const int64_t off = L.off[t];
    const int64_t lo = (int64_t)((int)blockIdx.x - L.blk0[t]) * ST_BLOCK_ELEMS;
    const int64_t hi = min(L.n[t], lo + ST_BLOCK_ELEMS);
    const AdamCoef c{beta1, beta2, eps, step_size, bc2_sqrt, k, regularised};
    if (hi - lo == ST_BLOCK_ELEMS && (off & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        // full block, 16-byte aligned: one float4 of every array per thread, all loads in flight at once
        const int64_t i = lo + 4 * threadIdx.x;
        float4 w = *reinterpret_cast<const float4*>(p + i);
        const float4 g = *reinterpret_cast<const float4*>(grads + off + i);
        float4 m = *reinterpret_cast<const float4*>(exp_avg + off + i);
        float4 v = *reinterpret_cast<const float4*>(exp_avg_sq + off + i);
        adam_update(c, w.x, g.x, m.x, v.x);
        adam_update(c, w.y, g.y, m.y, v.y);
        adam_update(c, w.z, g.z, m.z, v.z);
        adam_update(c, w.w, g.w, m.w, v.w);
        *reinterpret_cast<float4*>(p + i) = w;
        *reinterpret_cast<float4*>(exp_avg + off + i) = m;
        *reinterpret_cast<float4*>(exp_avg_sq + off + i) = v;
        return;
    }
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
        float w = p[i], m = exp_avg[off + i], v = exp_avg_sq[off + i];
        adam_update(c, w, grads[off + i], m, v);
        exp_avg[off + i] = m;
        exp_avg_sq[off + i] = v;
        p[i] = w;
    }
}

extern "C" int cp_adam_step(float* const* params, const int64_t* sizes, const int64_t* offsets, int n_tensors,
                            const float* grads, float* exp_avg, float* exp_avg_sq, const double* lr,
                            const int32_t* lr_index, const float* reg, const int32_t* norm_index, const float* norms,
                            const int64_t* step, double beta1, double beta2, double eps, void* stream) {
    if (!params || !sizes || !offsets || n_tensors <= 0 || n_tensors > ST_MAX || !grads || !exp_avg || !exp_avg_sq ||
        !lr || !lr_index || !reg || !norm_index || !step)
        return CP_ERR_ARG;
    AdamList L;
    L.count = n_tensors;
    int64_t blocks = 0;
    for (int t = 0; t < n_tensors; ++t) {
        if (!params[t] || sizes[t] <= 0 || offsets[t] < 0 || lr_index[t] < 0 || (norm_index[t] >= 0 && !norms))
            return CP_ERR_ARG;
        L.p[t] = params[t];
        L.n[t] = sizes[t];
        L.off[t] = offsets[t];
        L.lr_index[t] = lr_index[t];
        L.norm_index[t] = norm_index[t];
        L.reg[t] = reg[t];
        L.blk0[t] = (int)blocks;
        blocks += cp_cdiv(sizes[t], ST_BLOCK_ELEMS);
        if (blocks > 0x7fffffff) return CP_ERR_ARG;
    }
    L.blk0[n_tensors] = (int)blocks;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(L, grads, exp_avg, exp_avg_sq, lr, norms, step,
                                                                       beta1, beta2, eps);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
