// K4 / K4': windowed majority vote and class-subset evaluator.  Integer / indexing kernels.
//
// K4  replaces the per-group Python loop of 249 `pred[:win].mode(0)` launches + host syncs
//     (models.py:151-163) by one launch: one CTA per item, one thread per class row, an
//     incremental prefix-mode with torch's CPU tie rule (smallest label among the most frequent).
// K4' implements the README-only subset evaluator (README.md:11,15).  Instead of re-running a
//     masked argmax over |S|^2 logits per (trial, window), every logit row is ranked ONCE
//     (cp_rank_rows, uint8 order); the restricted argmax of any subset is then "first label of
//     the ranked row that is in the subset" -- ~41/|S| byte probes.  Logits are read from HBM
//     once for all trials (164 B/window).
#include <cstdlib>
#include "common.cuh"

#define T CP_TASKS
#define MAXW 32

// ------------------------------------------------------------------------------------- K4 vote
__global__ void __launch_bounds__(64)
vote_kernel(const int32_t* __restrict__ pred, int W, int n_votes, int32_t* __restrict__ votes,
            int64_t* __restrict__ y_pred) {
    __shared__ uint16_t cnt[T][T + 1];            // 16-bit: a row may collect all W <= 256 votes on one label
    __shared__ int correct_at[MAXW * 8];
    const int64_t b = blockIdx.x;
    const int i = threadIdx.x;
    for (int k = threadIdx.x; k < W; k += blockDim.x) correct_at[k] = 0;
    if (i < T)
        for (int j = 0; j < T; ++j) cnt[i][j] = 0;
    __syncthreads();
    int best = 0, best_c = 0;
    if (i < T) {
        for (int w = 0; w < W; ++w) {
            const int l = pred[(b * W + w) * T + i];
            if ((unsigned)l >= (unsigned)T) {        // not a class label: no vote (the prefix mode is unchanged)
                if (best == i && best_c > 0) atomicAdd(&correct_at[w], 1);
                continue;
            }
            const int c = ++cnt[i][l];
            // prefix mode, ties -> smallest label: the incremented label wins iff it now has
            // strictly more votes, or as many votes and a smaller label
            if (c > best_c || (c == best_c && l < best)) { best = l; best_c = c; }
            if (best == i) atomicAdd(&correct_at[w], 1);
        }
        y_pred[b * T + i] = best;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < n_votes; v += blockDim.x) {
        const int w = (v + 1 < W ? v + 1 : W) - 1;           // pred[:win] clamps at W rows
        votes[b * n_votes + v] = correct_at[w];
    }
}

extern "C" int cp_vote_eval(const int32_t* pred, int64_t B, int W, int n_votes, int32_t* votes,
                            int64_t* y_pred, void* stream) {
    if (B == 0) return CP_OK;
    if (!pred || !votes || !y_pred || B < 0 || W <= 0 || W > MAXW * 8 || n_votes <= 0) return CP_ERR_ARG;
    vote_kernel<<<(unsigned)B, 64, 0, (cudaStream_t)stream>>>(pred, W, n_votes, votes, y_pred);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// ------------------------------------------------------------------------------- K4' rank rows
// One thread per row: the row's 41 logits live in registers and every unordered pair is compared ONCE,
//   k beats j (j < k)  <=>  v_k > v_j        (ties: the smaller label ranks first)
// crediting rank_j or rank_k -- 820 compares per row instead of 41 x 41 -- then order[rank_j] = j.
// Rows are staged through shared memory so that global loads / stores stay coalesced (row stride 41 floats is
// odd: conflict-free when each thread reads its own row).  ALU-bound: 205 B of traffic per ~2.5 k instructions.
#define RR_ROWS 128
__global__ void __launch_bounds__(RR_ROWS)
rank_rows_kernel(const float* __restrict__ logits, int64_t n_rows, uint8_t* __restrict__ order) {
    __shared__ float v[RR_ROWS * T];
    __shared__ __align__(4) uint8_t ord[RR_ROWS * T];
    const int64_t r0 = (int64_t)blockIdx.x * RR_ROWS;
    const int nr = (int)min((int64_t)RR_ROWS, n_rows - r0);
    for (int e = threadIdx.x; e < nr * T; e += RR_ROWS) v[e] = __ldg(logits + r0 * T + e);
    __syncthreads();
    if (threadIdx.x < nr) {
        float x[T];
        int rank[T];
#pragma unroll
        for (int j = 0; j < T; ++j) {
            x[j] = v[threadIdx.x * T + j];
            rank[j] = 0;
        }
#pragma unroll
        for (int j = 0; j < T; ++j) {
#pragma unroll
            for (int k = j + 1; k < T; ++k) {
                const int beats = x[k] > x[j];
                rank[j] += beats;
                rank[k] += 1 - beats;
            }
        }
#pragma unroll
        for (int j = 0; j < T; ++j) ord[threadIdx.x * T + rank[j]] = (uint8_t)j;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nr * T; e += RR_ROWS) order[r0 * T + e] = ord[e];
}

extern "C" int cp_rank_rows(const float* logits, int64_t n_rows, uint8_t* order, void* stream) {
    if (n_rows == 0) return CP_OK;
    if (!logits || !order || n_rows < 0) return CP_ERR_ARG;
    rank_rows_kernel<<<(unsigned)cp_cdiv(n_rows, RR_ROWS), RR_ROWS, 0, (cudaStream_t)stream>>>(logits, n_rows, order);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// ----------------------------------------------------------------------------- K4' subset eval
// grid = (item b, trial chunk).  The CTA stages the ranked rows of its item (W*41*41 bytes = 42 KB
// at W=25) in shared memory once; each thread owns one trial of the chunk and walks the rows of
// its subset.  Per-thread label counters live in shared memory (byte lanes, column = thread).
#define SE_THREADS 128
__global__ void __launch_bounds__(SE_THREADS)
subset_eval_kernel(const uint8_t* __restrict__ order, int W, const uint8_t* __restrict__ masks,
                   int64_t n_trials, unsigned long long* __restrict__ correct) {
    extern __shared__ uint8_t smem[];
    uint8_t* ord = smem;                                   // [W][T][T]
    uint8_t* cnt = smem + ((W * T * T + 15) / 16) * 16;    // [T][SE_THREADS]
    const int64_t b = blockIdx.x;
    const uint8_t* src = order + b * (int64_t)W * T * T;
    for (int e = threadIdx.x; e < W * T * T; e += SE_THREADS) ord[e] = __ldg(src + e);
    for (int e = threadIdx.x; e < T * SE_THREADS; e += SE_THREADS) cnt[e] = 0;
    __syncthreads();
    const int tid = threadIdx.x;
    for (int64_t t = (int64_t)blockIdx.y * SE_THREADS + tid; t < n_trials;
         t += (int64_t)gridDim.y * SE_THREADS) {
        unsigned long long m = 0;
        for (int j = 0; j < T; ++j) m |= (unsigned long long)(__ldg(masks + t * T + j) != 0) << j;
        int n_ok = 0;
        for (int i = 0; i < T; ++i) {
            if (!((m >> i) & 1ull)) continue;
            int best = 0, best_c = 0;
            uint8_t seen[MAXW];
            for (int w = 0; w < W; ++w) {
                const uint8_t* row = ord + (w * T + i) * T;
                int k = 0;
                int l = row[0];
                while (!((m >> l) & 1ull)) l = row[++k];     // subset contains i, so this terminates
                seen[w] = (uint8_t)l;
                const int c = ++cnt[l * SE_THREADS + tid];
                if (c > best_c || (c == best_c && l < best)) { best = l; best_c = c; }
            }
            for (int w = 0; w < W; ++w) cnt[seen[w] * SE_THREADS + tid] = 0;
            n_ok += (best == i);
        }
        if (n_ok) atomicAdd(correct + t, (unsigned long long)n_ok);
    }
}

// Round 2: one WARP per (item, trial), one lane per window.  The thread-per-trial kernel above diverges on everything
// (every trial of a warp has its own subset: the row loop runs over the union of 32 masks, the probe walks serialise)
// and keeps its vote counters in shared memory; here control flow is uniform -- the warp walks the rows of ITS subset --
// each lane probes the ranked list of its window, and the windowed majority vote is two warp instructions:
// match.any groups the lanes by predicted label (popcount = votes), redux.max picks (votes, smallest label).
// Shared memory holds only the item's ranked rows, windows padded to 4 x 421 bytes so that the 25 lanes hit 25 banks.
#define SE2_WARPS 8
#define SE2_WSTRIDE (4 * ((T * T + 3) / 4 + (((T * T + 3) / 4) % 2 == 0 ? 1 : 0)))      // 1684: odd number of words
__global__ void __launch_bounds__(SE2_WARPS * 32)
subset_eval_warp_kernel(const uint8_t* __restrict__ order, int W, const uint8_t* __restrict__ masks,
                        int64_t n_trials, unsigned long long* __restrict__ correct) {
    extern __shared__ __align__(8) uint8_t se2_smem[];
    // above[w][i]: the labels ranked above label i in row (w, i) as a bit set -- window w predicts i for subset S  <=>
    // above[w][i] & S == 0, one AND instead of a probe walk; the vote count of i decides most rows by itself (below)
    unsigned long long* above = reinterpret_cast<unsigned long long*>(se2_smem);          // [W][T]
    uint8_t* ord = se2_smem + (size_t)W * T * 8;           // [W][SE2_WSTRIDE]: window w, row i, rank k at w*stride + i*T + k
    const int64_t b = blockIdx.x;
    const uint8_t* src = order + b * (int64_t)W * T * T;
    for (int w = 0; w < W; ++w)
        for (int e = threadIdx.x; e < T * T; e += SE2_WARPS * 32) ord[w * SE2_WSTRIDE + e] = __ldg(src + w * T * T + e);
    __syncthreads();
    for (int e = threadIdx.x; e < W * T; e += SE2_WARPS * 32) {
        const int w = e / T, i = e % T;
        const uint8_t* row = ord + w * SE2_WSTRIDE + i * T;
        unsigned long long ab = 0;
        for (int k = 0; k < T && row[k] != i; ++k) ab |= 1ull << row[k];
        above[e] = ab;
    }
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const bool active = lane < W;
    const unsigned wmask = W >= 32 ? 0xffffffffu : ((1u << W) - 1u);
    const uint8_t* base = ord + (active ? lane : 0) * SE2_WSTRIDE;
    for (int64_t t = (int64_t)blockIdx.y * SE2_WARPS + warp; t < n_trials; t += (int64_t)gridDim.y * SE2_WARPS) {
        const unsigned lo = __ballot_sync(0xffffffffu, __ldg(masks + t * T + lane) != 0);
        const unsigned hi = __ballot_sync(0xffffffffu, lane < T - 32 && __ldg(masks + t * T + 32 + (lane < T - 32 ? lane : 0)) != 0);
        const unsigned long long m = (unsigned long long)lo | ((unsigned long long)hi << 32);
        int n_ok = 0;
        const int n_s = __popcll(m);
        for (unsigned long long mm = m; mm; mm &= mm - 1) {
            const int i = __ffsll((long long)mm) - 1;
            // votes for i itself.  More than half of the windows: i wins whatever the others got.  Fewer than the mean
            // vote W / |S|: some other label has more, i loses.  Only the rows in between need the full vote.
            const bool mine = active && !(above[lane * T + i] & m);
            const int c_i = __popc(__ballot_sync(0xffffffffu, mine));
            if (2 * c_i > W) { ++n_ok; continue; }
            if (c_i * n_s < W) continue;
            int l = 63;                                      // idle lanes: a label no subset contains
            if (active) {
                // first ranked label inside the subset, four candidates per step (independent byte loads; the subset
                // contains i, so the walk ends by rank 40 -- the look-ahead reads at most 3 bytes into the window's padding)
                const uint8_t* row = base + i * T;
                for (int k = 0;; k += 4) {
                    const int l0 = row[k], l1 = row[k + 1], l2 = row[k + 2], l3 = row[k + 3];
                    const bool h0 = (m >> l0) & 1ull, h1 = (m >> (l1 & 63)) & 1ull, h2 = (m >> (l2 & 63)) & 1ull,
                               h3 = (m >> (l3 & 63)) & 1ull;
                    if (h0 | h1 | h2 | h3) {
                        l = h0 ? l0 : (h1 ? l1 : (h2 ? l2 : l3));
                        break;
                    }
                }
            }
            const unsigned peers = __match_any_sync(0xffffffffu, l);
            // votes of the lane's label, ties to the SMALLER label (prefix-mode rule of models.py:154)
            // (a loop-free histogram -- seven redux.add over packed 5-bit counters -- measured 14 % slower than match.any)
            unsigned key = active ? (((unsigned)__popc(peers & wmask) << 6) | (unsigned)(63 - l)) : 0u;
            key = __reduce_max_sync(0xffffffffu, key);
            n_ok += (63 - (int)(key & 63u)) == i;
        }
        if (lane == 0 && n_ok) atomicAdd(correct + t, (unsigned long long)n_ok);
    }
}

__global__ void subset_total_kernel(const uint8_t* __restrict__ masks, int64_t n_trials, int64_t B,
                                    int64_t* __restrict__ total) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_trials) return;
    int s = 0;
    for (int j = 0; j < T; ++j) s += masks[t * T + j] != 0;
    total[t] = B * s;
}

extern "C" int cp_subset_eval(const uint8_t* order, int64_t B, int W, const uint8_t* masks,
                              int64_t n_trials, int64_t* correct, int64_t* total, void* stream) {
    if (n_trials == 0) return CP_OK;
    if (!order || !masks || !correct || !total || B < 0 || W <= 0 || W > MAXW || n_trials < 0)
        return CP_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    CP_CUDA(cudaMemsetAsync(correct, 0, sizeof(int64_t) * n_trials, st));
    subset_total_kernel<<<(unsigned)cp_cdiv(n_trials, 256), 256, 0, st>>>(masks, n_trials, B, total);
    CP_CHECK_LAUNCH();
    if (B == 0) return CP_OK;
    static const bool warp_kernel = [] { const char* e = getenv("CP_SUBSET_WARP"); return !(e && e[0] == '0'); }();
    if (warp_kernel) {
        const size_t smem2 = (size_t)W * T * 8 + (size_t)W * SE2_WSTRIDE;
        CP_CUDA(cudaFuncSetAttribute(subset_eval_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        // trial chunks: ~4 resident CTAs on every SM, but at least ~8 trials per warp so that staging the item's rows pays
        int64_t chunks2 = cp_cdiv((int64_t)CP_NUM_SMS * 4, B);
        const int64_t most = cp_cdiv(n_trials, (int64_t)SE2_WARPS * 8);
        if (chunks2 > most) chunks2 = most;
        if (chunks2 < 1) chunks2 = 1;
        subset_eval_warp_kernel<<<dim3((unsigned)B, (unsigned)chunks2), SE2_WARPS * 32, smem2, st>>>(
            order, W, masks, n_trials, reinterpret_cast<unsigned long long*>(correct));
        CP_CHECK_LAUNCH();
        return CP_OK;
    }
    const size_t smem = ((size_t)(W * T * T + 15) / 16) * 16 + (size_t)T * SE_THREADS;
    CP_CUDA(cudaFuncSetAttribute(subset_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // trial chunks: enough CTAs for >= 2 waves of 148 SMs x 4 resident CTAs when B is small
    int64_t chunks = cp_cdiv(n_trials, SE_THREADS);
    int64_t want = cp_cdiv((int64_t)CP_NUM_SMS * 8, B);
    if (chunks > want) chunks = want < 1 ? 1 : want;
    dim3 grid((unsigned)B, (unsigned)chunks);
    subset_eval_kernel<<<grid, SE_THREADS, smem, st>>>(order, W, masks, n_trials,
                                                       reinterpret_cast<unsigned long long*>(correct));
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// ------------------------------------------------------------------------- confusion matrix
// counts[t, p] = #{k : y_true[k] == t and y_pred[k] == p}   (results.py:58: sklearn's confusion_matrix on the
// voted decisions, labels 0..C-1).  Per-CTA shared-memory histogram (C <= 64: 16 KB of int32), one 64-bit
// atomic per non-zero cell per CTA; integer adds, so the result does not depend on the order.
#define CM_MAX_C 64
__global__ void __launch_bounds__(256)
confusion_kernel(const int64_t* __restrict__ y_true, const int64_t* __restrict__ y_pred, int64_t n, int C,
                 unsigned long long* __restrict__ counts, int* __restrict__ err_flag) {
    __shared__ unsigned int h[CM_MAX_C * CM_MAX_C];
    for (int e = threadIdx.x; e < C * C; e += blockDim.x) h[e] = 0;
    __syncthreads();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = __ldg(y_true + k), p = __ldg(y_pred + k);
        if (t < 0 || t >= C || p < 0 || p >= C) {
            if (err_flag) *err_flag = 1;
            continue;
        }
        atomicAdd(&h[(int)t * C + (int)p], 1u);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < C * C; e += blockDim.x)
        if (h[e]) atomicAdd(counts + e, (unsigned long long)h[e]);
}

extern "C" int cp_confusion_matrix(const int64_t* y_true, const int64_t* y_pred, int64_t n, int n_classes,
                                   int64_t* counts, int* err_flag, void* stream) {
    if (!counts || n < 0 || n_classes <= 0 || n_classes > CM_MAX_C || (n > 0 && (!y_true || !y_pred)))
        return CP_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    CP_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * n_classes * n_classes, st));
    if (n == 0) return CP_OK;
    int64_t blocks = cp_cdiv(n, 256 * 8);
    if (blocks > CP_NUM_SMS * 4) blocks = CP_NUM_SMS * 4;
    confusion_kernel<<<(unsigned)blocks, 256, 0, st>>>(y_true, y_pred, n, n_classes,
                                                       reinterpret_cast<unsigned long long*>(counts), err_flag);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
