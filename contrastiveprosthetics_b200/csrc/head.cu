// K3: fused contrastive head, forward + backward in one pass.
//
// Replaces, per training step, Model.forward's head (models.py:121-130: two L2 normalisations,
// one_hot -> Linear(41,16), transpose, bmm) and Model.loss (models.py:198-208 -> 132-173: a Python
// loop over the B groups with 2B cross_entropy + B softmax/argmax launches and B host syncs).
// One persistent launch: each CTA walks groups (grid-stride), keeps the 41x41 similarity tile
// in shared memory (logits are never written to HBM unless the caller asks for them), and emits
// loss, argmax, correct counts and the gradients w.r.t. the un-normalised embeddings and the class
// table.  Accumulation across groups is deterministic: per-CTA partials + a one-CTA finalize.
//
// Bytes per group: 41*16*4 in + the same out = 5.2 KB -> HBM/latency bound.
#include "common.cuh"

#define T CP_TASKS
#define D CP_EMB_DIM
#define HEAD_THREADS 64
#define EPT ((T * D + HEAD_THREADS - 1) / HEAD_THREADS)     // class-table grad elements per thread

struct HeadPartial {
    float dC[T * D];
    double loss;
};

template <bool FROM_LOGITS>
__global__ void __launch_bounds__(HEAD_THREADS)
head_kernel(const float* __restrict__ emb, int64_t G, int W, const float* __restrict__ table_w,
            const float* __restrict__ table_b, const float* __restrict__ logits_in,
            float* __restrict__ d_emb, float* __restrict__ d_logits, int32_t* __restrict__ pred,
            int32_t* __restrict__ n_correct, float* __restrict__ logits_out,
            HeadPartial* __restrict__ partial, int need_grad) {
    __shared__ float E[T][D + 1];
    __shared__ float DE[T][D + 1];
    __shared__ float C[T][D + 1];
    __shared__ float S[T][T + 1];
    __shared__ float invn[T], lse_r[T], lse_c[T];
    __shared__ int s_correct;
    __shared__ double s_loss[HEAD_THREADS];

    const int tid = threadIdx.x;
    const float coef = (float)(1.0 / (2.0 * (double)G * (double)T));
    float dC_acc[EPT];
#pragma unroll
    for (int k = 0; k < EPT; ++k) dC_acc[k] = 0.f;
    double loss_acc = 0.0;

    if (!FROM_LOGITS) {
        // class table rows c_j = W[:, j] + b (GLOVENet default branch, models.py:457-458), normalised
        for (int e = tid; e < T * D; e += HEAD_THREADS) {
            const int j = e / D, d = e % D;
            C[j][d] = __ldg(table_w + d * T + j) + __ldg(table_b + d);
        }
        __syncthreads();
        if (tid < T) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < D; ++d) s += C[tid][d] * C[tid][d];
            const float n = sqrtf(s);
#pragma unroll
            for (int d = 0; d < D; ++d) C[tid][d] = C[tid][d] / n;
        }
        __syncthreads();
    }

    for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
        if (tid == 0) s_correct = 0;
        if (!FROM_LOGITS) {
            const int64_t b = g / W;
            const int w = (int)(g - b * W);
            for (int e = tid; e < T * D; e += HEAD_THREADS) {
                const int i = e / D, d = e % D;
                E[i][d] = __ldg(emb + ((b * T + i) * W + w) * D + d);
            }
            __syncthreads();
            if (tid < T) {
                float s = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) s += E[tid][d] * E[tid][d];
                const float n = sqrtf(s);
                invn[tid] = 1.0f / n;
#pragma unroll
                for (int d = 0; d < D; ++d) E[tid][d] = E[tid][d] / n;
            }
            __syncthreads();
            for (int e = tid; e < T * T; e += HEAD_THREADS) {
                const int i = e / T, j = e % T;
                float s = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) s = fmaf(E[i][d], C[j][d], s);
                S[i][j] = s;
                if (logits_out) logits_out[g * (T * T) + e] = s;
            }
        } else {
            for (int e = tid; e < T * T; e += HEAD_THREADS) S[e / T][e % T] = __ldg(logits_in + g * (T * T) + e);
        }
        __syncthreads();

        if (tid < T) {
            // row tid: log-sum-exp + first-max argmax (F.softmax(..).argmax(-1), models.py:148)
            float m = S[tid][0];
            int am = 0;
            for (int j = 1; j < T; ++j)
                if (S[tid][j] > m) { m = S[tid][j]; am = j; }
            float se = 0.f;
            for (int j = 0; j < T; ++j) se += expf(S[tid][j] - m);
            lse_r[tid] = m + logf(se);
            // column tid
            float mc = S[0][tid];
            for (int i = 1; i < T; ++i) mc = fmaxf(mc, S[i][tid]);
            float sc = 0.f;
            for (int i = 0; i < T; ++i) sc += expf(S[i][tid] - mc);
            lse_c[tid] = mc + logf(sc);
            loss_acc += (double)(lse_r[tid] - S[tid][tid]) + (double)(lse_c[tid] - S[tid][tid]);
            if (pred) pred[g * T + tid] = am;
            if (am == tid) atomicAdd(&s_correct, 1);
        }
        __syncthreads();
        if (tid == 0 && n_correct) n_correct[g] = s_correct;

        if (need_grad) {
            // dS = coef * (softmax_row + softmax_col - 2 I)
            for (int e = tid; e < T * T; e += HEAD_THREADS) {
                const int i = e / T, j = e % T;
                const float s = S[i][j];
                float v = expf(s - lse_r[i]) + expf(s - lse_c[j]);
                if (i == j) v -= 2.0f;
                v *= coef;
                if (FROM_LOGITS) d_logits[g * (T * T) + e] = v;
                S[i][j] = v;
            }
            if (!FROM_LOGITS) {
                __syncthreads();
                for (int e = tid; e < T * D; e += HEAD_THREADS) {
                    const int i = e / D, d = e % D;
                    float s = 0.f;
                    for (int j = 0; j < T; ++j) s = fmaf(S[i][j], C[j][d], s);
                    DE[i][d] = s;
                }
#pragma unroll
                for (int k = 0; k < EPT; ++k) {
                    const int e = tid + k * HEAD_THREADS;
                    if (e < T * D) {
                        const int j = e / D, d = e % D;
                        float s = 0.f;
                        for (int i = 0; i < T; ++i) s = fmaf(S[i][j], E[i][d], s);
                        dC_acc[k] += s;
                    }
                }
                __syncthreads();
                if (d_emb) {
                    const int64_t b = g / W;
                    const int w = (int)(g - b * W);
                    // through x / ||x||:  dx = (dy - y (y . dy)) / ||x||
                    for (int e = tid; e < T * D; e += HEAD_THREADS) {
                        const int i = e / D, d = e % D;
                        float dot = 0.f;
#pragma unroll
                        for (int q = 0; q < D; ++q) dot = fmaf(E[i][q], DE[i][q], dot);
                        d_emb[((b * T + i) * W + w) * D + d] = (DE[i][d] - E[i][d] * dot) * invn[i];
                    }
                }
            }
        }
        __syncthreads();
    }

    s_loss[tid] = loss_acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int k = 0; k < HEAD_THREADS; ++k) s += s_loss[k];
        partial[blockIdx.x].loss = s;
    }
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
        const int e = tid + k * HEAD_THREADS;
        if (e < T * D) partial[blockIdx.x].dC[e] = dC_acc[k];
    }
}

// Per-CTA partials -> HEAD_SLICES slice sums (double), one CTA per slice; then one CTA finishes the loss and the
// class-table gradient from the slices.  (A single CTA walking every partial -- 2.6 KB apart -- took 36 us at 592
// partials and grew with the grid, which capped head_kernel at 8 warps per SM.)
#define HEAD_SLICES 32
struct HeadSlice {
    double dC[T * D];
    double loss;
};
__global__ void __launch_bounds__(704)
head_slice_kernel(const HeadPartial* __restrict__ partial, int n_part, HeadSlice* __restrict__ slices) {
    const int per = (n_part + HEAD_SLICES - 1) / HEAD_SLICES;
    const int p0 = blockIdx.x * per, p1 = min(n_part, p0 + per);
    const int tid = threadIdx.x;
    if (tid < T * D) {
        // 4 independent chains in a fixed order (deterministic)
        double a[4] = {0, 0, 0, 0};
        int p = p0;
        for (; p + 4 <= p1; p += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] += (double)partial[p + u].dC[tid];
        }
        for (; p < p1; ++p) a[0] += (double)partial[p].dC[tid];
        slices[blockIdx.x].dC[tid] = (a[0] + a[1]) + (a[2] + a[3]);
    }
    if (tid == T * D) {
        double s = 0.0;
        for (int p = p0; p < p1; ++p) s += partial[p].loss;
        slices[blockIdx.x].loss = s;
    }
}

__global__ void __launch_bounds__(704)
head_finalize_kernel(const HeadSlice* __restrict__ slices, int64_t G,
                     const float* __restrict__ table_w, const float* __restrict__ table_b,
                     float* __restrict__ loss, float* __restrict__ d_table_w,
                     float* __restrict__ d_table_b) {
    __shared__ double dCh[T][D];      // gradient w.r.t. the normalised table
    __shared__ float dc[T][D];        // gradient w.r.t. the raw table rows
    const int tid = threadIdx.x;
    if (tid < T * D) {
        double a = 0.0;
        for (int p = 0; p < HEAD_SLICES; ++p) a += slices[p].dC[tid];
        dCh[tid / D][tid % D] = a;
    }
    if (tid == T * D && loss) {
        double s = 0.0;
        for (int p = 0; p < HEAD_SLICES; ++p) s += slices[p].loss;
        *loss = (float)(s / (2.0 * (double)G * (double)T));
    }
    __syncthreads();
    if (!d_table_w && !d_table_b) return;
    if (tid < T) {
        const int j = tid;
        double c[D], n2 = 0.0;
        for (int d = 0; d < D; ++d) {
            c[d] = (double)(table_w[d * T + j] + table_b[d]);
            n2 += c[d] * c[d];
        }
        const double n = sqrt(n2);
        double dot = 0.0;
        for (int d = 0; d < D; ++d) dot += (c[d] / n) * dCh[j][d];
        for (int d = 0; d < D; ++d) dc[j][d] = (float)((dCh[j][d] - (c[d] / n) * dot) / n);
    }
    __syncthreads();
    if (tid < T * D && d_table_w) {
        const int d = tid / T, j = tid % T;
        d_table_w[d * T + j] = dc[j][d];
    }
    if (tid < D && d_table_b) {
        float s = 0.f;
        for (int j = 0; j < T; ++j) s += dc[j][tid];
        d_table_b[tid] = s;
    }
}

static int head_blocks(int64_t G) {
    const int64_t cap = (int64_t)CP_NUM_SMS * 16;      // 64-thread CTAs: up to 32 warps per SM (the partials are reduced
                                                        // by HEAD_SLICES CTAs, so the grid no longer costs the finalize)
    return (int)(G < cap ? G : cap);
}
static size_t head_partial_bytes(int64_t G) { return cp_align(sizeof(HeadPartial) * (size_t)head_blocks(G < 1 ? 1 : G)); }

extern "C" size_t cp_head_workspace_bytes(int64_t n_groups) {
    return head_partial_bytes(n_groups) + cp_align(sizeof(HeadSlice) * HEAD_SLICES);
}

extern "C" int cp_head_forward_backward(const float* emb, int64_t B, int W, const float* table_w,
                                        const float* table_b, float* loss, float* d_emb,
                                        float* d_table_w, float* d_table_b, int32_t* pred,
                                        int32_t* n_correct, float* logits, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    if (!emb || !table_w || !table_b || !workspace || B <= 0 || W <= 0) return CP_ERR_ARG;
    const int64_t G = B * W;
    if (workspace_bytes < cp_head_workspace_bytes(G)) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = head_blocks(G);
    const int need_grad = (d_emb || d_table_w || d_table_b) ? 1 : 0;
    HeadPartial* part = reinterpret_cast<HeadPartial*>(workspace);
    head_kernel<false><<<nb, HEAD_THREADS, 0, st>>>(emb, G, W, table_w, table_b, nullptr, d_emb, nullptr,
                                                    pred, n_correct, logits, part, need_grad);
    CP_CHECK_LAUNCH();
    HeadSlice* slices = reinterpret_cast<HeadSlice*>(reinterpret_cast<char*>(workspace) + head_partial_bytes(G));
    head_slice_kernel<<<HEAD_SLICES, 704, 0, st>>>(part, nb, slices);
    CP_CHECK_LAUNCH();
    head_finalize_kernel<<<1, 704, 0, st>>>(slices, G, table_w, table_b, loss, d_table_w, d_table_b);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_logits_loss(const float* logits, int64_t G, float* loss, float* d_logits,
                              int32_t* pred, int32_t* n_correct, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!logits || !workspace || G <= 0) return CP_ERR_ARG;
    if (workspace_bytes < cp_head_workspace_bytes(G)) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = head_blocks(G);
    HeadPartial* part = reinterpret_cast<HeadPartial*>(workspace);
    head_kernel<true><<<nb, HEAD_THREADS, 0, st>>>(nullptr, G, 1, nullptr, nullptr, logits, nullptr, d_logits,
                                                   pred, n_correct, nullptr, part, d_logits ? 1 : 0);
    CP_CHECK_LAUNCH();
    HeadSlice* slices = reinterpret_cast<HeadSlice*>(reinterpret_cast<char*>(workspace) + head_partial_bytes(G));
    head_slice_kernel<<<HEAD_SLICES, 704, 0, st>>>(part, nb, slices);
    CP_CHECK_LAUNCH();
    head_finalize_kernel<<<1, 704, 0, st>>>(slices, G, nullptr, nullptr, loss, nullptr, nullptr);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
