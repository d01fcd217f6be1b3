// K2: EMG encoder forward / backward orchestration (EMGNet, models.py:230-342).
//
//   x (n,12) -> conv k3 (1->64) -> ReLU -> BN2d -> conv k3 (64->64) -> ReLU -> BN2d -> flatten
//            -> 7 x [Linear -> ReLU -> BN1d (-> Dropout after blocks 4..7)] -> Linear 512->16
//
// HBM layout (fp32, all inside the caller's workspace):
//   conv stages   [n*12, 64]   position-major, channel-contiguous: the k=3 convolution becomes an
//                              implicit GEMM whose A row (n,p) is the 192 contiguous floats that
//                              start at row (n,p-1); the flatten is index p*64+c, so fc1 runs on a
//                              column-permuted copy of its weight (prep_weights_kernel).
//   linear stages [n, 512]
//   per stage: Y (post-ReLU, pre-BN; kept for backward) and A (post-BN/dropout = next GEMM input).
//   The conv1 stage has no Y: its 3-FMA activation is recomputed from the 48-byte window wherever it is needed.
// BatchNorm statistics are produced by the GEMM / conv epilogues as per-CTA partial column sums
// (no second pass over the activation) and finalised in double precision.
//
// Engines (cp_encoder_opts.engine): CP_ENGINE_SIMT runs every GEMM as fp32 FFMA; CP_ENGINE_TC_FP16 is CP_ENGINE_TC with
// ONE tensor-core product on the hi planes only (TF32-class accuracy; the reduced-precision path); CP_ENGINE_TC runs conv2 and
// the seven linear layers on tcgen05 with the 3-product fp16 split of gemm_tc.cuh: the BN-apply and
// BN-backward kernels then write their outputs as two fp16 planes (hi, lo) -- both inside the stage's fp32
// slot, hi in the first half and lo in the second -- that the TMA-fed GEMMs consume directly, and the weights
// are split (and transposed for the data gradient) once per call.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "encoder_kernels.cuh"
#include "proj_fused.cuh"

namespace {

constexpr int F_CONV = 64;
constexpr int F_FC = 512;
constexpr int K_FC1 = 768;
constexpr int WGRAD_SPLITS_MAX = 32;                       // at the largest (512x768) weight
constexpr size_t WPART_ELEMS = (size_t)WGRAD_SPLITS_MAX * F_FC * K_FC1;
constexpr int WS_ZERO_WORDS = 64 + 16 + 8 + 8 + 8 + 8;    // tickets, abound, wmax, l1max, rowl1, bmax: zeroed per forward call
constexpr int N_FOLD = 3;                                 // linear layers 2..4 take their input BatchNorm folded into W

struct Ws {
    float *X0, *Y1, *A1, *Y2, *A2;
    float *Y[CP_N_FC], *A[CP_N_FC];
    uint8_t* keep[4];
    float *G0, *G1;
    float *mean[CP_N_BN], *istd[CP_N_BN], *scale[CP_N_BN], *shift[CP_N_BN];
    float *pa, *pb;            // per-CTA column partials
    float *m1, *m2;
    double* totals;            // [2*512 + 1] SyncBN: rank-local column totals + row count, all-reduced in place
    double* rscratch;          // [RP_SLABS][2][512] slab sums of the two-level partial reductions
    unsigned int* tickets;     // [64] last-CTA tickets, then (zeroed together, once per forward call):
    unsigned int* abound;      // [16] bound on |BatchNorm output| per stage (bit pattern) -> scale of the A planes
    unsigned int* wmax;        // [8] max |W| of the 7 linear layers and conv2 (bit pattern) -> scale of the W planes
    float* ascale_inv;         // [16] 1 / (power-of-two scale of the stage's A planes)
    float* wscale_inv;         // [8]  1 / (power-of-two scale of the layer's W planes; 7 = conv2)
    float *wpart;              // split-K weight-gradient partials
    float *Wc2, *Wc2d, *W1p;
    float *c1w, *c1b;          // copy of the conv1 parameters (parity tap)
    float *ppart;              // projection / conv1 weight-gradient partials
    // tensor-core engine: an activation / gradient slot holds both fp16 planes (hi_of / lo_of); split weights
    float *Wc2_lo, *Wc2d_lo;
    float *G1b;               // second pre-activation-gradient buffer (weight-gradient GEMMs run on a side stream)
    plane_t *Wh[CP_N_FC], *Wl[CP_N_FC], *Wth[CP_N_FC], *Wtl[CP_N_FC];
    unsigned int* gmax;        // [16] max |g'| per backward stage (bit pattern) + [16] gamma == 0 flags per stage,
                               // zeroed per backward call
    float* gscale_inv;         // [16] 1 / (power-of-two scale of the stage's G1 planes)
    double* cancel;            // [16][WS_CANCEL_PARTS][2] cancellation estimates of the reduce-free BN backward
    unsigned int* l1max;       // [8] max column L1 norm of the linear weights (bit pattern; zeroed with tickets, forward)
    unsigned int* g1max;       // [16] max |pre-activation gradient| per stage (bit pattern; zeroed with gmax, backward)
    float* coef;               // [3][512] c1, c2, c3 of the fused BN-backward epilogue (one stage at a time)
    float* gz_bound;           // [16] bound on |gz| per stage
    // BatchNorm of linear blocks 1..3 folded into the weights of layers 2..4 (fold_bn_weights_kernel): per forward call
    unsigned int *rowl1, *bmax;  // [8] max row L1 norm / max |bias| of the linear layers (bit patterns, zeroed with tickets)
    plane_t *Wfh[N_FOLD], *Wfl[N_FOLD];   // planes of W_{l+1} diag(scale_l)
    float* bias_f[N_FOLD];       // b_{l+1} + W_{l+1} shift_l
    float* wfscale_inv;          // [N_FOLD] 1 / (power-of-two scale of those planes)
    size_t bytes;
};

// the two fp16 planes of a tensor-core operand live in the stage's fp32 slot of `elems` floats
inline const plane_t* hi_of(const float* slot) { return reinterpret_cast<const plane_t*>(slot); }
inline const plane_t* lo_of(const float* slot, size_t elems) { return reinterpret_cast<const plane_t*>(slot) + elems; }

struct Carver {
    char* base;
    size_t off = 0;
    template <typename Tp> Tp* take(size_t n) {
        Tp* p = reinterpret_cast<Tp*>(base + off);
        off += cp_align(n * sizeof(Tp));
        return p;
    }
};

int64_t partial_rows(int64_t n) {
    // the largest number of per-CTA partial rows any stage produces
    return cp_cdiv(n * 12, 128) + cp_cdiv(n, 128) + 8;
}

Ws carve(void* base, int64_t n, const cp_encoder_opts* o) {
    Ws w;
    Carver c{reinterpret_cast<char*>(base)};
    const bool save = o->save_for_backward != 0;
    const size_t conv_elems = (size_t)n * 12 * F_CONV, fc_elems = (size_t)n * F_FC;
    w.X0 = c.take<float>((size_t)n * 12);
    if (save) {
        w.Y1 = nullptr; w.A1 = c.take<float>(conv_elems);
        w.Y2 = c.take<float>(conv_elems); w.A2 = c.take<float>(conv_elems);
        for (int l = 0; l < CP_N_FC; ++l) {
            w.Y[l] = c.take<float>(fc_elems);
            // the last block's output only feeds the projection: fused, never stored (proj_fused.cuh)
            w.A[l] = l + 1 < CP_N_FC ? c.take<float>(fc_elems) : nullptr;
        }
        w.G0 = c.take<float>(conv_elems);
        w.G1 = c.take<float>(conv_elems);
    } else {
        // inference: every stage reuses one Y and two alternating A buffers
        float* y = c.take<float>(conv_elems);
        float* a0 = c.take<float>(conv_elems);
        float* a1 = c.take<float>(conv_elems);
        w.Y1 = nullptr; w.Y2 = y; w.A1 = a0; w.A2 = a1;
        for (int l = 0; l < CP_N_FC; ++l) { w.Y[l] = y; w.A[l] = (l & 1) ? a1 : a0; }
        w.G0 = w.G1 = nullptr;
    }
    for (int d = 0; d < 4; ++d) w.keep[d] = (o->dropout_p > 0.f) ? c.take<uint8_t>(fc_elems) : nullptr;
    for (int l = 0; l < CP_N_BN; ++l) {
        w.mean[l] = c.take<float>(F_FC); w.istd[l] = c.take<float>(F_FC);
        w.scale[l] = c.take<float>(F_FC); w.shift[l] = c.take<float>(F_FC);
    }
    const size_t pr = (size_t)partial_rows(n);
    w.pa = c.take<float>(pr * F_FC);
    w.pb = c.take<float>(pr * F_FC);
    w.m1 = c.take<float>(F_FC);
    w.m2 = c.take<float>(F_FC);
    w.totals = c.take<double>(2 * F_FC + 8);
    w.rscratch = c.take<double>((size_t)RP_SLABS * 2 * F_FC);
    w.tickets = c.take<unsigned int>(WS_ZERO_WORDS);
    w.abound = w.tickets + 64;
    w.wmax = w.abound + 16;
    w.l1max = w.wmax + 8;
    w.rowl1 = w.l1max + 8;
    w.bmax = w.rowl1 + 8;
    w.ascale_inv = c.take<float>(16);
    w.wscale_inv = c.take<float>(8);
    w.wpart = save ? c.take<float>(WPART_ELEMS) : nullptr;
    w.Wc2 = c.take<float>(64 * 192);
    w.Wc2d = c.take<float>(64 * 192);
    w.W1p = c.take<float>(F_FC * K_FC1);
    w.c1w = c.take<float>(64 * 9);
    w.c1b = c.take<float>(64);
    w.ppart = save ? c.take<float>((size_t)pf::grid_for(n) * CP_EMB_DIM * 512 +
                                   (size_t)cp_cdiv(n, C1_WIN) * 3 * 64)
                   : nullptr;
    w.Wc2_lo = w.Wc2d_lo = w.G1b = nullptr;
    for (int l = 0; l < CP_N_FC; ++l) w.Wh[l] = w.Wl[l] = w.Wth[l] = w.Wtl[l] = nullptr;
    w.gmax = c.take<unsigned int>(48);
    w.g1max = w.gmax + 32;
    w.coef = c.take<float>(3 * F_FC);
    w.gz_bound = c.take<float>(16);
    w.gscale_inv = c.take<float>(16);
    w.cancel = c.take<double>(16 * WS_CANCEL_PARTS * 2);
    if (o->engine != CP_ENGINE_SIMT) {
        w.Wc2_lo = c.take<float>(64 * 192);
        w.Wc2d_lo = c.take<float>(64 * 192);
        if (save) w.G1b = c.take<float>(conv_elems);
        for (int l = 0; l < CP_N_FC; ++l) {
            const size_t we = (size_t)F_FC * (l == 0 ? K_FC1 : F_FC);
            w.Wh[l] = c.take<plane_t>(we); w.Wl[l] = c.take<plane_t>(we);
            w.Wth[l] = c.take<plane_t>(we); w.Wtl[l] = c.take<plane_t>(we);
        }
    }
    for (int f = 0; f < N_FOLD; ++f) {
        const bool on = o->engine != CP_ENGINE_SIMT;
        w.Wfh[f] = on ? c.take<plane_t>((size_t)F_FC * F_FC) : nullptr;
        w.Wfl[f] = on ? c.take<plane_t>((size_t)F_FC * F_FC) : nullptr;
        w.bias_f[f] = on ? c.take<float>(F_FC) : nullptr;
    }
    w.wfscale_inv = c.take<float>(8);
    w.bytes = c.off;
    return w;
}

inline unsigned ew_grid(int64_t n_vec) {
    int64_t b = cp_cdiv(n_vec, 256);
    const int64_t cap = (int64_t)CP_NUM_SMS * 16;
    return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// C = act(A . B^T + bias) (+ column statistics partials)
template <int BM, int BN, int BMODE, bool ACONV>
int launch_nt(const float* A, int64_t M, int K, int lda, const float* B, int N, int ldb, const float* bias,
              float* C, int ldc, float* psum, float* psq, int relu, cudaStream_t st) {
    if (K % GEMM_BK != 0 || N % BN != 0) return CP_ERR_ARG;
    GemmNT g{A, M, K, lda, B, N, ldb, bias, C, ldc, psum, psq, relu};
    const int64_t tiles = cp_cdiv(M, BM) * (N / BN);
    gemm_nt_kernel<BM, BN, BMODE, ACONV><<<(unsigned)tiles, 256, 0, st>>>(g);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// dW[Mo,No] = G^T . A, split over row slabs; out written through wgrad_reduce (mode = re-layout)
template <int BM, int BN, bool ACONV>
int launch_wgrad(const float* G, int ldg, int Mo, const float* A, int lda, int No, int64_t R, float* wpart,
                 float* out, int mode, cudaStream_t st, const float* G_lo = nullptr, const float* A_lo = nullptr,
                 size_t wpart_elems = WPART_ELEMS) {
    if (Mo % BM != 0 || No % BN != 0) return CP_ERR_ARG;
    const int tiles = (Mo / BM) * (No / BN);
    constexpr int resident = BM * BN >= 128 * 128 ? 2 : 4;      // CTAs per SM the tile shape allows
    int S = (resident * CP_NUM_SMS + tiles - 1) / tiles;        // one full wave of 148 SMs
    const int64_t max_s = cp_cdiv(R, 8 * GEMM_BK);
    if (S > max_s) S = (int)max_s;
    const int64_t cap = (int64_t)(wpart_elems / ((size_t)Mo * No));
    if (S > cap) S = (int)cap;
    if (S < 1) S = 1;
    int64_t rps = cp_cdiv(cp_cdiv(R, S), GEMM_BK) * GEMM_BK;
    S = (int)cp_cdiv(R, rps);
    GemmTN g{G, ldg, Mo, A, lda, No, R, rps, wpart, G_lo, A_lo};
    dim3 grid(No / BN, Mo / BM, S);
    gemm_tn_kernel<BM, BN, ACONV><<<grid, 256, 0, st>>>(g);
    CP_CHECK_LAUNCH();
    wgrad_reduce_kernel<<<(unsigned)cp_cdiv((int64_t)Mo * No, 256), 256, 0, st>>>(wpart, S, Mo, No, out, mode);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// SyncBN: sum the 2F+1 doubles at w.totals over all ranks through the caller's collective
int sync_totals(const Ws& w, int F, const cp_encoder_opts* o, cudaStream_t st) {
    const int rc = o->allreduce(o->allreduce_user, w.totals, (size_t)(2 * F + 1), (void*)st);
    return rc == 0 ? CP_OK : CP_ERR_COLLECTIVE;
}

int bn_finalize(const Ws& w, int l, int F, int P, int64_t R, const cp_encoder_tensors* p,
                const cp_encoder_opts* o, cudaStream_t st) {
    if (o->bn_mode != CP_BN_BATCH && (!p->bn_rm[l] || !p->bn_rv[l])) return CP_ERR_ARG;
    if (o->allreduce && o->bn_mode != CP_BN_RUNNING) {
        bn_finalize_kernel<<<dim3(F / 32, rp_slabs(P)), 1024, 0, st>>>(
            w.pa, w.pb, P, F, R, nullptr, nullptr, nullptr, nullptr, o->bn_mode, 0.f, 0.f, nullptr, nullptr, nullptr,
            nullptr, w.rscratch, w.tickets, w.totals);
        CP_CHECK_LAUNCH();
        if (int rc = sync_totals(w, F, o, st)) return rc;
        bn_finalize_totals_kernel<<<(F + 511) / 512, 512, 0, st>>>(w.totals, F, p->bn_w[l], p->bn_b[l], p->bn_rm[l],
                                                                   p->bn_rv[l], o->bn_mode, o->bn_momentum, o->bn_eps,
                                                                   w.mean[l], w.istd[l], w.scale[l], w.shift[l],
                                                                   w.abound + l);
        CP_CHECK_LAUNCH();
        return CP_OK;
    }
    bn_finalize_kernel<<<dim3(F / 32, o->bn_mode == CP_BN_RUNNING ? 1 : rp_slabs(P)), 1024, 0, st>>>(
        w.pa, w.pb, P, F, R, p->bn_w[l], p->bn_b[l], p->bn_rm[l], p->bn_rv[l], o->bn_mode, o->bn_momentum, o->bn_eps,
        w.mean[l], w.istd[l], w.scale[l], w.shift[l], w.rscratch, w.tickets, nullptr, w.abound + l);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// planes: write the (hi, lo) fp16 planes into the slot `a` instead of the fp32 value
template <int F>
int bn_apply(const float* y, float* a, bool planes, int64_t R, const Ws& w, int l, uint8_t* keep,
             float inv_keep, cudaStream_t st, float gen_p = 0.f, uint64_t seed = 0, uint64_t layer = 0,
             const unsigned long long* seed_offset = nullptr, int fast = 0) {
    // (the single-product engine never reads the lo plane; it is still written -- an `if` in the store path cost the
    //  parity engine's BN kernels 10 %)
    float* a_lo = planes ? reinterpret_cast<float*>(reinterpret_cast<plane_t*>(a) + (size_t)R * F) : nullptr;
    if (planes)
        bn_apply_kernel<F, true><<<ew_grid(R * (F / 4)), 256, 0, st>>>(y, a, a_lo, R, w.scale[l], w.shift[l], keep,
                                                                     inv_keep, gen_p, seed, layer, seed_offset,
                                                                     w.abound + l, w.ascale_inv + l);
    else
        bn_apply_kernel<F, false><<<ew_grid(R * (F / 4)), 256, 0, st>>>(y, a, nullptr, R, w.scale[l], w.shift[l], keep,
                                                                      inv_keep, gen_p, seed, layer, seed_offset);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// BN backward + ReLU backward of stage l:  g (grad w.r.t. stage output A) -> gz (grad w.r.t. the
// pre-activation), d_gamma, d_beta, d_bias
// BN backward sums of stage l:  m1 = mean g', m2 = mean g' xh, d_gamma, d_beta.
// stats_ready: w.m1 / w.m2 / d_gamma / d_beta were already produced from the next layer's parameter gradients
// (bn_bwd_stats_from_wgrad_kernel): the two kernels still launch, but return at once unless that derivation asked for
// the exact pass (gamma == 0: flag w.gmax[16 + l]; bad cancellation: w.cancel, see stage_needs_exact)
template <int F>
int bn_bwd_sums(const float* g, const float* y, bool planes, int64_t R, const Ws& w, int l, const uint8_t* keep,
                float inv_keep, float* d_gamma, float* d_beta, cudaStream_t st, const cp_encoder_opts* o,
                bool stats_ready) {
    const int P = (int)cp_cdiv(R, ColMap<F>::ROWS);
    const unsigned int* run_flag = stats_ready ? w.gmax + 16 + l : nullptr;
    const bool sync = o->allreduce != nullptr;
    if (stats_ready && !sync) {
        // the sums were derived from the next layer's dW: reduce + finalize as ONE conditional launch (normally empty)
        bn_bwd_reduce_finalize_kernel<F><<<P, 256, 0, st>>>(g, y, R, keep, inv_keep, w.mean[l], w.istd[l], w.pa, w.pb,
                                                            planes ? w.gmax + l : nullptr, run_flag, w.m1, w.m2, d_gamma,
                                                            d_beta, w.tickets + 48 + l);
        CP_CHECK_LAUNCH();
        return CP_OK;
    }
    bn_bwd_reduce_kernel<F><<<P, 256, 0, st>>>(g, y, R, keep, inv_keep, w.mean[l], w.istd[l], w.pa, w.pb, nullptr,
                                               planes ? w.gmax + l : nullptr, run_flag);
    CP_CHECK_LAUNCH();
    bn_bwd_finalize_kernel<<<dim3(F / 32, rp_slabs(P)), 1024, 0, st>>>(w.pa, w.pb, P, F, R, w.m1, w.m2, d_gamma, d_beta,
                                                                    w.rscratch, w.tickets, sync ? w.totals : nullptr,
                                                                    run_flag);
    CP_CHECK_LAUNCH();
    if (sync) {
        if (int rc = sync_totals(w, F, o, st)) return rc;
        bn_bwd_means_totals_kernel<<<(F + 511) / 512, 512, 0, st>>>(w.totals, F, w.m1, w.m2);
        CP_CHECK_LAUNCH();
    }
    return CP_OK;
}

// BN backward + ReLU backward of stage l:  g (grad w.r.t. stage output A) -> gz (grad w.r.t. the
// pre-activation), d_gamma, d_beta, d_bias
template <int F>
int bn_backward(const float* g, const float* y, float* gz, bool planes, int64_t R, const Ws& w, int l,
                const uint8_t* keep, float inv_keep, const float* gamma, float* d_gamma, float* d_beta,
                float* d_bias, cudaStream_t st, const cp_encoder_opts* o, bool stats_ready = false) {
    // small batches (the reference's batch_size 8 = 328 windows): 64 sequential rows per thread on 3 CTAs took 26 us;
    // 8 rows per thread on 16x the CTAs (the partial rows of w.pa stay inside partial_rows(n): 16 >= 128 / 13)
    const int rpc = (F == 512 && R < (int64_t)ColMap<F>::ROWS * 2 * CP_NUM_SMS) ? 16 : ColMap<F>::ROWS;
    const int P = (int)cp_cdiv(R, rpc);
    float* gz_lo = planes ? reinterpret_cast<float*>(reinterpret_cast<plane_t*>(gz) + (size_t)R * F) : nullptr;
    if (int rc = bn_bwd_sums<F>(g, y, planes, R, w, l, keep, inv_keep, d_gamma, d_beta, st, o, stats_ready)) return rc;
    if (planes)
        bn_bwd_apply_kernel<F, true><<<P, 256, 0, st>>>(g, y, R, keep, inv_keep, w.mean[l], w.istd[l], gamma, w.m1,
                                                        w.m2, gz, gz_lo, w.pa, nullptr, w.gmax + l, w.gscale_inv + l,
                                                        w.g1max + l, rpc);
    else
        bn_bwd_apply_kernel<F, false><<<P, 256, 0, st>>>(g, y, R, keep, inv_keep, w.mean[l], w.istd[l], gamma, w.m1,
                                                         w.m2, gz, nullptr, w.pa, nullptr, nullptr, nullptr, nullptr, rpc);
    CP_CHECK_LAUNCH();
    colsum_finalize_kernel<<<F / 32, 1024, 0, st>>>(w.pa, P, F, d_bias, 0);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

#define CP_TRY(expr)                 \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != CP_OK) return rc__; \
    } while (0)

// Side stream for the weight-gradient GEMMs of the tensor-core engine: they are off the critical
// path (dgrad -> BN backward -> dgrad ...) and, being tensor-pipe bound, overlap with the HBM-bound
// BN-backward kernels of the next layer.  Fork/join is by events, so the caller's stream semantics
// (and CUDA-graph capture) are preserved.  One per DEVICE, created on first use; `lock` serialises the host-side
// enqueue of cp_encoder_backward calls on that device (two host threads recording / waiting on the same events
// would otherwise pick up each other's records) -- the GPU work of different caller streams still overlaps.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t ready[2] = {nullptr, nullptr};     // main -> side: G1 buffer b has been written
    cudaEvent_t done[2] = {nullptr, nullptr};      // side -> main: G1 buffer b has been consumed
    bool ok = false;
    std::mutex lock;
    int init() {                                   // call with `lock` held
        if (ok) return CP_OK;
        CP_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CP_CUDA(cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming));
            CP_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
        }
        ok = true;
        return CP_OK;
    }
};
// CP_FUSE_BNBWD=0 in the environment keeps the un-fused BN backward (A/B measurements)
const bool g_fuse_bnbwd = []() { const char* e = getenv("CP_FUSE_BNBWD"); return !(e && e[0] == '0'); }();
constexpr int CP_MAX_DEVICES = 64;
SideStream g_side_of[CP_MAX_DEVICES];
SideStream* side_of_current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= CP_MAX_DEVICES) return nullptr;
    return &g_side_of[d];
}

// Backward of the last linear block from d_emb: projection weight gradient, BN backward, ReLU backward and the
// bias gradient in two passes over (Y7, mask, d_emb) -- see proj_fused.cuh.  gz: pre-activation gradient
// (fp16 planes when `planes`).
int last_block_backward(const float* d_emb, float* gz, bool planes, int64_t n, const Ws& w, const uint8_t* keep,
                        float inv_keep, const cp_encoder_tensors* p, const cp_encoder_tensors* gr, cudaStream_t st,
                        const cp_encoder_opts* o) {
    constexpr int LL = CP_N_FC - 1, S = CP_N_BN - 1;           // linear layer / BN stage of the last block
    if (((uintptr_t)d_emb) % 16 != 0) return CP_ERR_ARG;
    const int G = pf::grid_for(n);
    CP_ONCE_PER_DEVICE({
        CP_TRY(pf::set_smem(pf::proj_bwd_reduce_kernel<false>));
        CP_TRY(pf::set_smem(pf::proj_bwd_reduce_kernel<true>));
        CP_TRY(pf::set_smem(pf::proj_bwd_apply_kernel<false, false>));
        CP_TRY(pf::set_smem(pf::proj_bwd_apply_kernel<false, true>));
        CP_TRY(pf::set_smem(pf::proj_bwd_apply_kernel<true, false>));
        CP_TRY(pf::set_smem(pf::proj_bwd_apply_kernel<true, true>));
    });
    unsigned int* gmax = planes ? w.gmax + S : nullptr;
    if (keep)
        pf::proj_bwd_reduce_kernel<true><<<G, 256, pf::SMEM, st>>>(w.Y[LL], keep, d_emb, n, inv_keep, w.scale[S], w.shift[S],
                                                                   w.mean[S], w.istd[S], p->proj_w, w.pa, w.pb, w.ppart, gmax);
    else
        pf::proj_bwd_reduce_kernel<false><<<G, 256, pf::SMEM, st>>>(w.Y[LL], nullptr, d_emb, n, 1.f, w.scale[S], w.shift[S],
                                                                    w.mean[S], w.istd[S], p->proj_w, w.pa, w.pb, w.ppart, gmax);
    CP_CHECK_LAUNCH();
    colsum_finalize_kernel<<<CP_EMB_DIM * F_FC / 32, 1024, 0, st>>>(w.ppart, G, CP_EMB_DIM * F_FC, gr->proj_w, 0);
    CP_CHECK_LAUNCH();
    const bool sync = o->allreduce != nullptr;
    bn_bwd_finalize_kernel<<<dim3(F_FC / 32, rp_slabs(G)), 1024, 0, st>>>(w.pa, w.pb, G, F_FC, n, w.m1, w.m2, gr->bn_w[S],
                                                                       gr->bn_b[S], w.rscratch, w.tickets,
                                                                       sync ? w.totals : nullptr);
    CP_CHECK_LAUNCH();
    if (sync) {
        CP_TRY(sync_totals(w, F_FC, o, st));
        bn_bwd_means_totals_kernel<<<1, 512, 0, st>>>(w.totals, F_FC, w.m1, w.m2);
        CP_CHECK_LAUNCH();
    }
    float* gz_lo = planes ? reinterpret_cast<float*>(reinterpret_cast<plane_t*>(gz) + (size_t)n * F_FC) : nullptr;
#define CP_PF_APPLY(K, SP)                                                                                              \
    pf::proj_bwd_apply_kernel<K, SP><<<G, 256, pf::SMEM, st>>>(w.Y[LL], keep, d_emb, n, inv_keep, w.mean[S], w.istd[S],  \
                                                               p->bn_w[S], w.m1, w.m2, p->proj_w, gz, gz_lo, w.pa, gmax, \
                                                               w.gscale_inv + S, planes ? w.g1max + S : nullptr)
    if (keep) { if (planes) CP_PF_APPLY(true, true); else CP_PF_APPLY(true, false); }
    else { if (planes) CP_PF_APPLY(false, true); else CP_PF_APPLY(false, false); }
#undef CP_PF_APPLY
    CP_CHECK_LAUNCH();
    colsum_finalize_kernel<<<F_FC / 32, 1024, 0, st>>>(w.pa, G, F_FC, gr->fc_b[LL], 0);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

bool opts_ok(const cp_encoder_opts* o) {
    return o && o->bn_mode >= 0 && o->bn_mode <= 2 && o->dropout_p >= 0.f && o->dropout_p < 1.f &&
           (o->engine == CP_ENGINE_SIMT || o->engine == CP_ENGINE_TC || o->engine == CP_ENGINE_TC_FP16);
}

// BatchNorm of linear blocks 1..3 folded into the weights of layers 2..4 (tensor-core engine, CTA-pair GEMM): forward and
// backward of one call must agree, so both ask here.  CP_FOLD_BN=0 in the environment keeps the BN-apply passes (A/B runs)
const bool g_fold_bn = []() { const char* e = getenv("CP_FOLD_BN"); return !(e && e[0] == '0'); }();
bool fold_bn_active(const cp_encoder_opts* o, int64_t n) {
    return g_fold_bn && o->engine != CP_ENGINE_SIMT && tcg::bnbwd_supported(n);
}

// the A operand of a weight gradient is the pre-BatchNorm activation of a folded stage: see wgrad_reduce_kernel
struct FoldFix { const float *scale, *shift, *db; };

// weight-gradient through the tensor-core split-K kernel + the shared re-layout / reduce kernel
int tc_wgrad(const plane_t* Gh, const plane_t* Gl, int Mo, const plane_t* Ah, const plane_t* Al, int No, int64_t R,
             float* wpart, float* out, int mode, cudaStream_t st, const float* g_scale_inv, int fast = 0,
             bool alone = false, const float* a_scale_inv = nullptr, const FoldFix* fx = nullptr) {
    int S = 0;
    CP_TRY(tcg::launch_tn(Gh, Gl, Mo, Mo, Ah, Al, No, No, R, wpart, WPART_ELEMS, &S, st, fast, alone));
    wgrad_reduce_kernel<<<(unsigned)cp_cdiv((int64_t)Mo * No, 256), 256, 0, st>>>(wpart, S, Mo, No, out, mode, g_scale_inv,
                                                                                  a_scale_inv, fx ? fx->scale : nullptr,
                                                                                  fx ? fx->shift : nullptr,
                                                                                  fx ? fx->db : nullptr);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

}  // namespace

extern "C" size_t cp_encoder_workspace_bytes(int64_t n, const cp_encoder_opts* opts) {
    if (n <= 0 || !opts_ok(opts)) return 0;
    return carve(nullptr, n, opts).bytes;
}

extern "C" int cp_encoder_forward(const cp_encoder_tensors* p, const float* x, int64_t n, float* emb,
                                  void* workspace, size_t workspace_bytes, const cp_encoder_opts* o,
                                  void* stream) {
    if (!p || !x || !emb || !workspace || n <= 0 || !opts_ok(o)) return CP_ERR_ARG;
    if (((uintptr_t)workspace) % 256 != 0) return CP_ERR_ARG;
    const bool tcE = o->engine != CP_ENGINE_SIMT;
    const int fast = o->engine == CP_ENGINE_TC_FP16;      // the GEMMs read the hi planes only
    const Ws w = carve(workspace, n, o);
    if (workspace_bytes < w.bytes) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t R12 = n * 12;
    const size_t conv_elems = (size_t)n * 12 * F_CONV, fc_elems = (size_t)n * F_FC;

    CP_CUDA(cudaMemsetAsync(w.tickets, 0, WS_ZERO_WORDS * sizeof(unsigned int), st));
    const bool fold = fold_bn_active(o, n);
    {
        WBoundsArgs ba;
        ba.rowl1 = w.rowl1; ba.bmax = w.bmax; ba.wmax = w.wmax; ba.l1max = w.l1max;
        for (int l = 0; l < CP_N_FC; ++l) {
            ba.ra.W[l] = p->fc_w[l]; ba.ra.b[l] = p->fc_b[l]; ba.ra.K[l] = l == 0 ? K_FC1 : F_FC;
            ba.wa.W[l] = p->fc_w[l]; ba.wa.n[l] = F_FC * (l == 0 ? K_FC1 : F_FC);
        }
        ba.wa.W[7] = p->conv2_w; ba.wa.n[7] = 64 * 64 * 9;
        static_assert(F_FC / 8 == 64 && K_FC1 / 256 == 3, "weights_bounds_kernel's block layout");
        ba.n_row = fold ? 64 * CP_N_FC : 0;
        ba.n_abs = tcE ? 48 * 8 : 0;
        ba.n_col = tcE && o->save_for_backward ? 3 * CP_N_FC : 0;
        if (ba.n_row + ba.n_abs + ba.n_col > 0) {
            weights_bounds_kernel<<<ba.n_row + ba.n_abs + ba.n_col, 256, 0, st>>>(ba);
            CP_CHECK_LAUNCH();
        }
    }
    prep_weights_kernel<<<(F_FC * K_FC1 + 255) / 256, 256, 0, st>>>(p->conv2_w, p->fc_w[0], w.Wc2, w.Wc2d, w.W1p,
                                                                    w.Wc2_lo, w.Wc2d_lo, p->conv1_w, p->conv1_b, w.c1w, w.c1b,
                                                                    w.wmax, w.wscale_inv);
    CP_CHECK_LAUNCH();
    if (tcE) {
        PrepTcArgs a;
        a.wmax = w.wmax; a.wscale_inv = w.wscale_inv;
        for (int l = 0; l < CP_N_FC; ++l) {
            a.W[l] = p->fc_w[l];
            a.Wh[l] = w.Wh[l]; a.Wl[l] = w.Wl[l]; a.Wth[l] = w.Wth[l]; a.Wtl[l] = w.Wtl[l];
        }
        prep_weights_tc_kernel<<<dim3(K_FC1 / 32, F_FC / 32, CP_N_FC), 256, 0, st>>>(a);
        CP_CHECK_LAUNCH();
    }
    CP_CUDA(cudaMemcpyAsync(w.X0, x, sizeof(float) * R12, cudaMemcpyDeviceToDevice, st));

    // conv1 -> ReLU -> BN: a statistics pass and an apply pass, both recomputing the activation from x
    const int P1 = (int)cp_cdiv(n, C1_WIN);
    conv1_fwd_kernel<<<P1, 256, 0, st>>>(w.X0, n, p->conv1_w, p->conv1_b, nullptr, w.pa, w.pb);
    CP_CHECK_LAUNCH();
    CP_TRY(bn_finalize(w, 0, F_CONV, P1, R12, p, o, st));
    if (tcE)
        conv1_bn_apply_kernel<true><<<ew_grid(n * 16), 256, 0, st>>>(
            w.X0, n, p->conv1_w, p->conv1_b, w.scale[0], w.shift[0], w.A1,
            reinterpret_cast<float*>(reinterpret_cast<plane_t*>(w.A1) + conv_elems), w.abound + 0, w.ascale_inv + 0);
    else
        conv1_bn_apply_kernel<false><<<ew_grid(n * 16), 256, 0, st>>>(w.X0, n, p->conv1_w, p->conv1_b, w.scale[0],
                                                                       w.shift[0], w.A1, nullptr);
    CP_CHECK_LAUNCH();

    // conv2 as implicit GEMM [n*12, 192] x [64, 192]^T
    if (tcE) {
        CP_TRY(tcg::launch_conv_nt(hi_of(w.A1), lo_of(w.A1, conv_elems), n, hi_of(w.Wc2), hi_of(w.Wc2_lo), p->conv2_b,
                                   w.Y2, w.pa, w.pb, 1, st, w.ascale_inv + 0, fast, w.wscale_inv + 7));
        CP_TRY(bn_finalize(w, 1, F_CONV, (int)cp_cdiv(n, tcg::CONV_WIN), R12, p, o, st));
    } else {
        CP_TRY((launch_nt<128, 64, 0, true>(w.A1, R12, 192, 64, w.Wc2, 64, 192, p->conv2_b, w.Y2, 64, w.pa, w.pb, 1, st)));
        CP_TRY(bn_finalize(w, 1, F_CONV, (int)cp_cdiv(R12, 128), R12, p, o, st));
    }
    CP_TRY(bn_apply<F_CONV>(w.Y2, w.A2, tcE, R12, w, 1, nullptr, 1.f, st, 0.f, 0, 0, nullptr, fast));

    // 7 x Linear -> ReLU -> BN (-> Dropout)
    const float inv_keep = o->dropout_p > 0.f ? 1.f / (1.f - o->dropout_p) : 1.f;
    for (int l = 0; l < CP_N_FC; ++l) {
        const float* in = l == 0 ? w.A2 : w.A[l - 1];
        const int K = l == 0 ? K_FC1 : F_FC;
        const float* W = l == 0 ? w.W1p : p->fc_w[l];
        // BatchNorm folding (linear blocks 1..3 -> layers 2..4, no dropout in between): layer l's epilogue also writes the
        // fp16 planes of its OUTPUT y (into the slot the BN-apply kernel would have filled), the stage's (scale, shift)
        // go into the next layer's weights / bias, and that layer's GEMM reads the y planes: no BN-apply pass.
        const bool in_folded = fold && l >= 1 && l <= N_FOLD;
        const bool out_planes = fold && l < N_FOLD;
        if (tcE) {
            const plane_t* in_lo = lo_of(in, l == 0 ? conv_elems : fc_elems);
            tcg::YPlanes yp{reinterpret_cast<plane_t*>(w.A[l]), reinterpret_cast<plane_t*>(w.A[l]) + fc_elems,
                            w.abound + 1 + l, w.rowl1 + l, w.bmax + l, w.ascale_inv + 2 + l};
            CP_TRY(tcg::launch_nt(hi_of(in), in_lo, n, K, K, in_folded ? w.Wfh[l - 1] : w.Wh[l],
                                  in_folded ? w.Wfl[l - 1] : w.Wl[l], F_FC, K, in_folded ? w.bias_f[l - 1] : p->fc_b[l],
                                  w.Y[l], F_FC, w.pa, w.pb, 1, st, w.ascale_inv + 1 + l, fast, nullptr, nullptr, 1.f,
                                  in_folded ? w.wfscale_inv + (l - 1) : w.wscale_inv + l, nullptr,
                                  out_planes ? &yp : nullptr));
        } else {
            CP_TRY((launch_nt<128, 128, 0, false>(in, n, K, K, W, F_FC, K, p->fc_b[l], w.Y[l], F_FC, w.pa, w.pb, 1, st)));
        }
        CP_TRY(bn_finalize(w, 2 + l, F_FC, (int)cp_cdiv(n, 128), n, p, o, st));
        uint8_t* keep = nullptr;
        float gen_p = 0.f;
        if (l >= 3 && o->dropout_p > 0.f) {
            const int d = l - 3;
            if (o->ext_masks)
                CP_CUDA(cudaMemcpyAsync(w.keep[d], o->ext_masks + (size_t)d * n * F_FC, (size_t)n * F_FC,
                                        cudaMemcpyDeviceToDevice, st));
            else
                gen_p = o->dropout_p;              // mask drawn (and stored) inside the BN-apply kernel
            keep = w.keep[d];
        }
        if (out_planes) {
            fold_bn_weights_kernel<<<F_FC, 128, 0, st>>>(p->fc_w[l + 1], p->fc_b[l + 1], w.scale[2 + l], w.shift[2 + l],
                                                         w.wmax + l + 1, w.Wfh[l], w.Wfl[l], w.bias_f[l], w.wfscale_inv + l);
            CP_CHECK_LAUNCH();
            continue;
        }
        if (l + 1 < CP_N_FC) {
            CP_TRY(bn_apply<F_FC>(w.Y[l], w.A[l], tcE, n, w, 2 + l, keep, inv_keep, st, gen_p, o->dropout_seed,
                                  (uint64_t)(l - 3), (const unsigned long long*)o->dropout_step, fast));
            continue;
        }
        if (o->trunk_only) {
            // --prediction mode: the block output itself (fp32) is the result; the classifier head follows (cls.cuh)
            CP_TRY(bn_apply<F_FC>(w.Y[l], emb, false, n, w, 2 + l, keep, inv_keep, st, gen_p, o->dropout_seed,
                                  (uint64_t)(l - 3), (const unsigned long long*)o->dropout_step, fast));
            continue;
        }
        // last block: BN (+ dropout) fused with the 512 -> 16 projection
        const int G = pf::grid_for(n);
        const unsigned long long* step = (const unsigned long long*)o->dropout_step;
        CP_ONCE_PER_DEVICE({
            CP_TRY(pf::set_smem(pf::bn_apply_proj_kernel<0>));
            CP_TRY(pf::set_smem(pf::bn_apply_proj_kernel<1>));
            CP_TRY(pf::set_smem(pf::bn_apply_proj_kernel<2>));
        });
        if (!keep)
            pf::bn_apply_proj_kernel<0><<<G, pf::FWD_THREADS, pf::SMEM, st>>>(w.Y[l], n, w.scale[2 + l], w.shift[2 + l], nullptr, 1.f, 0.f, 0, 0,
                                                                  nullptr, p->proj_w, emb);
        else if (gen_p > 0.f)
            pf::bn_apply_proj_kernel<2><<<G, pf::FWD_THREADS, pf::SMEM, st>>>(w.Y[l], n, w.scale[2 + l], w.shift[2 + l], keep, inv_keep, gen_p,
                                                                  o->dropout_seed, (uint64_t)(l - 3), step, p->proj_w, emb);
        else
            pf::bn_apply_proj_kernel<1><<<G, pf::FWD_THREADS, pf::SMEM, st>>>(w.Y[l], n, w.scale[2 + l], w.shift[2 + l], keep, inv_keep, 0.f, 0,
                                                                  0, nullptr, p->proj_w, emb);
        CP_CHECK_LAUNCH();
    }
    return CP_OK;
}

extern "C" int cp_encoder_backward(const cp_encoder_tensors* p, const float* d_emb, int64_t n,
                                   const cp_encoder_tensors* gr, void* workspace, size_t workspace_bytes,
                                   const cp_encoder_opts* o, void* stream) {
    if (!p || !d_emb || !gr || !workspace || n <= 0 || !opts_ok(o)) return CP_ERR_ARG;
    if (!o->save_for_backward || o->bn_mode == CP_BN_RUNNING) return CP_ERR_UNSUPPORTED;
    const bool tcE = o->engine != CP_ENGINE_SIMT;
    const int fast = o->engine == CP_ENGINE_TC_FP16;      // the GEMMs read the hi planes only
    const Ws w = carve(workspace, n, o);
    if (workspace_bytes < w.bytes) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t R12 = n * 12;
    const size_t conv_elems = (size_t)n * 12 * F_CONV, fc_elems = (size_t)n * F_FC;
    const float inv_keep = o->dropout_p > 0.f ? 1.f / (1.f - o->dropout_p) : 1.f;
    CP_CUDA(cudaMemsetAsync(w.gmax, 0, 48 * sizeof(unsigned int), st));
    const bool fold = fold_bn_active(o, n);

    const int Pp = pf::grid_for(n);          // projection weight-gradient partial rows (last_block_backward)

    // linear blocks, last to first.  G0 = grad w.r.t. block output, G1 = grad w.r.t. pre-activation
    cudaEvent_t join_event = nullptr;
    SideStream* side = tcE ? side_of_current_device() : nullptr;
    if (tcE && !side) return CP_ERR_UNSUPPORTED;
    std::unique_lock<std::mutex> side_guard;
    if (side) side_guard = std::unique_lock<std::mutex>(side->lock);
    if (tcE) {
        SideStream& g_side = *side;
        CP_TRY(g_side.init());
        cudaStream_t ss = g_side.stream;
        int nb = 0;                                    // stages processed so far -> G1 buffer parity
        bool used[2] = {false, false};
        auto g1 = [&](int b) { return b ? w.G1b : w.G1; };
        auto g1lo = [&](int b, size_t elems) { return lo_of(g1(b), elems); };
        bool stats_ready = false;                      // BN sums of the stage about to be processed already known
        bool stage_done = false;                       // ... and its whole BN backward already done (fused epilogue)
        // The BN + ReLU backward of a stage that feeds its layer WITHOUT dropout (conv2 stage, linear blocks 1..3) is
        // fused into that layer's data-gradient GEMM: the gradient w.r.t. the stage output is never stored.
        const bool fuse_ok = o->allreduce == nullptr && tcg::bnbwd_supported(n) && g_fuse_bnbwd;
        cudaEvent_t last_side = nullptr;               // completion of the latest side-stream GEMM (owner of w.wpart)
        for (int l = CP_N_FC - 1; l >= 0; --l, ++nb) {
            const int b = nb & 1;
            const uint8_t* keep = (l >= 3 && o->dropout_p > 0.f) ? w.keep[l - 3] : nullptr;
            if (used[b]) CP_CUDA(cudaStreamWaitEvent(st, g_side.done[b], 0));     // WAR on the G1 buffer
            if (stage_done)
                ;                                      // g1(b) already holds this stage's pre-activation gradient
            else if (l == CP_N_FC - 1 && o->trunk_only)  // d_emb IS the gradient w.r.t. the last block's output
                CP_TRY(bn_backward<F_FC>(d_emb, w.Y[l], g1(b), true, n, w, 2 + l, keep, inv_keep, p->bn_w[2 + l],
                                         gr->bn_w[2 + l], gr->bn_b[2 + l], gr->fc_b[l], st, o, false));
            else if (l == CP_N_FC - 1)
                CP_TRY(last_block_backward(d_emb, g1(b), true, n, w, keep, inv_keep, p, gr, st, o));
            else
                CP_TRY(bn_backward<F_FC>(w.G0, w.Y[l], g1(b), true, n, w, 2 + l, keep, inv_keep, p->bn_w[2 + l],
                                         gr->bn_w[2 + l], gr->bn_b[2 + l], gr->fc_b[l], st, o, stats_ready));
            stage_done = false;
            const int K = l == 0 ? K_FC1 : F_FC;
            const float* ain = l == 0 ? w.A2 : w.A[l - 1];
            const plane_t* ah = hi_of(ain);
            const plane_t* al = lo_of(ain, l == 0 ? conv_elems : fc_elems);
            const float* gsi = w.gscale_inv + 2 + l;
            // layers 2..4 with their input BatchNorm folded: the "A planes" hold the stage's pre-BN activation
            const FoldFix fx{w.scale[1 + l], w.shift[1 + l], gr->fc_b[l]};
            const FoldFix* fxp = (fold && l >= 1 && l <= N_FOLD) ? &fx : nullptr;
            // The stage below (BN stage 1 + l: conv2 for l = 0) feeds this layer without dropout and statistics are
            // rank-local: its BN-backward sums follow from this layer's dW / db (bn_bwd_stats_from_wgrad_kernel), so
            // the weight gradient runs in line (its overlap with the BN backward bought nothing) and the
            // reduce pass over (dA, Y) of that stage is skipped
            // With dropout between the stage and this layer the column sums of g' = dA * keep/(1-p) come from the
            // data-gradient GEMM's epilogue (masked column sums), the second sum still from dW.
            const bool below_has_dropout = l - 1 >= 3 && o->dropout_p > 0.f;
            const bool algebraic = o->allreduce == nullptr;
            const uint8_t* keep_below = below_has_dropout ? w.keep[l - 1 - 3] : nullptr;
            const bool masked = algebraic && keep_below != nullptr;
            if (fuse_ok && !below_has_dropout) {
                const int s_below = 1 + l;             // BN stage of this layer's input
                const int Fb = l == 0 ? F_CONV : F_FC;
                const int64_t Rb = l == 0 ? R12 : n;
                const float* y_below = l == 0 ? w.Y2 : w.Y[l - 1];
                const unsigned int* flag = w.gmax + 16 + s_below;
                // (1) dW_l, (2) the stage's BN-backward sums from dW_l / db_l
                if (last_side) CP_CUDA(cudaStreamWaitEvent(st, last_side, 0));
                CP_TRY(tc_wgrad(hi_of(g1(b)), g1lo(b, fc_elems), F_FC, ah, al, K, n, w.wpart, gr->fc_w[l], l == 0 ? 1 : 0, st, gsi, fast,
                                true, w.ascale_inv + 1 + l, fxp));
                if (l == 0)
                    bn_bwd_stats_from_wgrad_kernel<12><<<K_FC1 / WgradStats<12>::COLS, 512, 0, st>>>(
                        p->fc_w[0], gr->fc_w[0], gr->fc_b[0], F_FC, K_FC1, n, p->bn_w[s_below], p->bn_b[s_below], w.m1, w.m2,
                        gr->bn_w[s_below], gr->bn_b[s_below], nullptr, w.gmax + 16 + s_below,
                        w.cancel + (size_t)s_below * WS_CANCEL_PARTS * 2, w.tickets + 32 + s_below);
                else
                    bn_bwd_stats_from_wgrad_kernel<1><<<F_FC / WgradStats<1>::COLS, 512, 0, st>>>(
                        p->fc_w[l], gr->fc_w[l], gr->fc_b[l], F_FC, F_FC, n, p->bn_w[s_below], p->bn_b[s_below], w.m1, w.m2,
                        gr->bn_w[s_below], gr->bn_b[s_below], nullptr, w.gmax + 16 + s_below,
                        w.cancel + (size_t)s_below * WS_CANCEL_PARTS * 2, w.tickets + 32 + s_below);
                CP_CHECK_LAUNCH();
                // (3) exact fallback, all three launches return at once unless the derivation above asked for it:
                //     the plain data gradient -> G0, then the sums the long way
                CP_TRY(tcg::launch_nt(hi_of(g1(b)), g1lo(b, fc_elems), n, F_FC, F_FC, w.Wth[l], w.Wtl[l], K, F_FC, nullptr, w.G0,
                                      K, nullptr, nullptr, 0, st, gsi, fast, nullptr, nullptr, 1.f, w.wscale_inv + l, flag));
                if (l == 0)
                    CP_TRY(bn_bwd_sums<F_CONV>(w.G0, y_below, false, Rb, w, s_below, nullptr, 1.f, gr->bn_w[s_below],
                                               gr->bn_b[s_below], st, o, true));
                else
                    CP_TRY(bn_bwd_sums<F_FC>(w.G0, y_below, false, Rb, w, s_below, nullptr, 1.f, gr->bn_w[s_below],
                                             gr->bn_b[s_below], st, o, true));
                // (4) epilogue coefficients + plane-scale bound, (5) the fused data gradient
                float *c1 = w.coef, *c2 = w.coef + F_FC, *c3 = w.coef + 2 * F_FC;
                if (l == 0)
                    bn_bwd_coef_kernel<F_CONV><<<1, F_CONV, 0, st>>>(p->bn_w[s_below], w.mean[s_below], w.istd[s_below], w.m1, w.m2,
                                                                     w.g1max + 2 + l, w.l1max + l, c1, c2, c3, w.gz_bound + s_below);
                else
                    bn_bwd_coef_kernel<F_FC><<<1, F_FC, 0, st>>>(p->bn_w[s_below], w.mean[s_below], w.istd[s_below], w.m1, w.m2,
                                                                 w.g1max + 2 + l, w.l1max + l, c1, c2, c3, w.gz_bound + s_below);
                CP_CHECK_LAUNCH();
                const int bo = b ^ 1;                  // the other G1 buffer receives the stage's pre-activation gradient
                if (used[bo]) CP_CUDA(cudaStreamWaitEvent(st, g_side.done[bo], 0));
                plane_t* go_hi = reinterpret_cast<plane_t*>(g1(bo));
                plane_t* go_lo = go_hi + (l == 0 ? conv_elems : fc_elems);
                CP_TRY(tcg::launch_nt_bnbwd(hi_of(g1(b)), g1lo(b, fc_elems), n, F_FC, w.Wth[l], w.Wtl[l], K, y_below, c1, c2, c3, Fb,
                                            w.gz_bound + s_below, go_hi, go_lo, w.pa, w.g1max + s_below, w.gscale_inv + s_below,
                                            gsi, w.wscale_inv + l, fast, st));
                if (l == 0)
                    colsum_fold12_kernel<<<F_CONV / 32, 1024, 0, st>>>(w.pa, (int)cp_cdiv(n, 128), gr->conv2_b);
                else
                    colsum_finalize_kernel<<<F_FC / 32, 1024, 0, st>>>(w.pa, (int)cp_cdiv(n, 128), F_FC, gr->fc_b[l - 1], 0);
                CP_CHECK_LAUNCH();
                (void)Fb; (void)flag;
                stage_done = true;
                stats_ready = false;
                continue;
            }
            // main stream: G0 = G1 . W_l
            CP_TRY(tcg::launch_nt(hi_of(g1(b)), g1lo(b, fc_elems), n, F_FC, F_FC, w.Wth[l], w.Wtl[l], K, F_FC, nullptr, w.G0,
                                  K, masked ? w.pa : nullptr, masked ? w.pb : nullptr, 0, st, gsi, fast,
                                  algebraic ? w.gmax + 1 + l : nullptr, masked ? keep_below : nullptr, inv_keep,
                                  w.wscale_inv + l));
            if (algebraic) {
                if (last_side) CP_CUDA(cudaStreamWaitEvent(st, last_side, 0));     // w.wpart is shared with the side stream
                CP_TRY(tc_wgrad(hi_of(g1(b)), g1lo(b, fc_elems), F_FC, ah, al, K, n, w.wpart, gr->fc_w[l], l == 0 ? 1 : 0, st, gsi, fast,
                                true, w.ascale_inv + 1 + l, fxp));
                const int s_below = 1 + l;             // BN stage of this layer's input
                if (masked) {
                    // sum over the GEMM's per-tile partial rows -> w.m1 (scratch until the statistics kernel overwrites it)
                    colsum_finalize_kernel<<<F_FC / 32, 1024, 0, st>>>(w.pa, (int)cp_cdiv(n, 128), F_FC, w.m1, 0);
                    CP_CHECK_LAUNCH();
                    bn_bwd_stats_from_wgrad_kernel<1><<<F_FC / WgradStats<1>::COLS, 512, 0, st>>>(
                        p->fc_w[l], gr->fc_w[l], gr->fc_b[l], F_FC, F_FC, n, p->bn_w[s_below], p->bn_b[s_below], w.m1, w.m2,
                        gr->bn_w[s_below], gr->bn_b[s_below], w.m1, w.gmax + 16 + s_below,
                        w.cancel + (size_t)s_below * WS_CANCEL_PARTS * 2, w.tickets + 32 + s_below);
                } else if (l == 0)
                    bn_bwd_stats_from_wgrad_kernel<12><<<K_FC1 / WgradStats<12>::COLS, 512, 0, st>>>(
                        p->fc_w[0], gr->fc_w[0], gr->fc_b[0], F_FC, K_FC1, n, p->bn_w[s_below], p->bn_b[s_below], w.m1, w.m2,
                        gr->bn_w[s_below], gr->bn_b[s_below], nullptr, w.gmax + 16 + s_below,
                        w.cancel + (size_t)s_below * WS_CANCEL_PARTS * 2, w.tickets + 32 + s_below);
                else
                    bn_bwd_stats_from_wgrad_kernel<1><<<F_FC / WgradStats<1>::COLS, 512, 0, st>>>(
                        p->fc_w[l], gr->fc_w[l], gr->fc_b[l], F_FC, F_FC, n, p->bn_w[s_below], p->bn_b[s_below], w.m1, w.m2,
                        gr->bn_w[s_below], gr->bn_b[s_below], nullptr, w.gmax + 16 + s_below,
                        w.cancel + (size_t)s_below * WS_CANCEL_PARTS * 2, w.tickets + 32 + s_below);
                CP_CHECK_LAUNCH();
                stats_ready = true;
                continue;
            }
            stats_ready = false;
            // side stream: dW_l = G1^T . A_{l-1}.  It starts when the data-gradient GEMM above has finished (two
            // persistent GEMMs cannot share an SM), i.e. alongside the HBM-bound BN-backward kernels of layer l-1
            CP_CUDA(cudaEventRecord(g_side.ready[b], st));
            CP_CUDA(cudaStreamWaitEvent(ss, g_side.ready[b], 0));
            CP_TRY(tc_wgrad(hi_of(g1(b)), g1lo(b, fc_elems), F_FC, ah, al, K, n, w.wpart, gr->fc_w[l], l == 0 ? 1 : 0, ss, gsi, fast,
                            false, w.ascale_inv + 1 + l, fxp));
            CP_CUDA(cudaEventRecord(g_side.done[b], ss));
            used[b] = true;
            last_side = g_side.done[b];
        }
        // conv2 block: G0 is [n*12, 64] (same memory order as the [n,768] position-major flatten)
        const int b = nb & 1;
        if (used[b]) CP_CUDA(cudaStreamWaitEvent(st, g_side.done[b], 0));
        if (!stage_done)
            CP_TRY(bn_backward<F_CONV>(w.G0, w.Y2, g1(b), true, R12, w, 1, nullptr, 1.f, p->bn_w[1], gr->bn_w[1],
                                       gr->bn_b[1], gr->conv2_b, st, o, stats_ready));
        CP_CUDA(cudaEventRecord(g_side.ready[b], st));
        CP_CUDA(cudaStreamWaitEvent(ss, g_side.ready[b], 0));
        CP_CUDA(cudaMemsetAsync(gr->conv2_w, 0, sizeof(float) * 64 * 64 * 9, ss));
        int S = 0;
        CP_TRY(tcg::launch_conv_tn(hi_of(w.A1), lo_of(w.A1, conv_elems), hi_of(g1(b)), g1lo(b, conv_elems), n, w.wpart,
                                   WPART_ELEMS, &S, ss, fast));
        wgrad_reduce_kernel<<<(192 * 64 + 255) / 256, 256, 0, ss>>>(w.wpart, S, 192, 64, gr->conv2_w, 3, w.gscale_inv + 1,
                                                                    w.ascale_inv + 0);
        CP_CHECK_LAUNCH();
        CP_CUDA(cudaEventRecord(g_side.done[b], ss));
        CP_TRY(tcg::launch_conv_nt(hi_of(g1(b)), g1lo(b, conv_elems), n, hi_of(w.Wc2d), hi_of(w.Wc2d_lo), nullptr, w.G0,
                                   nullptr, nullptr, 0, st, w.gscale_inv + 1, fast, w.wscale_inv + 7));
        join_event = g_side.done[b];
    } else {
        for (int l = CP_N_FC - 1; l >= 0; --l) {
            const uint8_t* keep = (l >= 3 && o->dropout_p > 0.f) ? w.keep[l - 3] : nullptr;
            if (l == CP_N_FC - 1 && o->trunk_only)
                CP_TRY(bn_backward<F_FC>(d_emb, w.Y[l], w.G1, false, n, w, 2 + l, keep, inv_keep, p->bn_w[2 + l],
                                         gr->bn_w[2 + l], gr->bn_b[2 + l], gr->fc_b[l], st, o));
            else if (l == CP_N_FC - 1)
                CP_TRY(last_block_backward(d_emb, w.G1, false, n, w, keep, inv_keep, p, gr, st, o));
            else
                CP_TRY(bn_backward<F_FC>(w.G0, w.Y[l], w.G1, false, n, w, 2 + l, keep, inv_keep, p->bn_w[2 + l],
                                         gr->bn_w[2 + l], gr->bn_b[2 + l], gr->fc_b[l], st, o));
            if (l > 0) {
                CP_TRY((launch_wgrad<128, 128, false>(w.G1, F_FC, F_FC, w.A[l - 1], F_FC, F_FC, n, w.wpart, gr->fc_w[l], 0, st)));
                CP_TRY((launch_nt<128, 128, 1, false>(w.G1, n, F_FC, F_FC, p->fc_w[l], F_FC, F_FC, nullptr, w.G0, F_FC,
                                                      nullptr, nullptr, 0, st)));
            } else {
                CP_TRY((launch_wgrad<128, 128, false>(w.G1, F_FC, F_FC, w.A2, K_FC1, K_FC1, n, w.wpart, gr->fc_w[0], 1, st)));
                CP_TRY((launch_nt<128, 128, 1, false>(w.G1, n, F_FC, F_FC, w.W1p, K_FC1, K_FC1, nullptr, w.G0, K_FC1,
                                                      nullptr, nullptr, 0, st)));
            }
        }
        // conv2 block: G0 is [n*12, 64] (same memory order as the [n,768] position-major flatten)
        CP_TRY(bn_backward<F_CONV>(w.G0, w.Y2, w.G1, false, R12, w, 1, nullptr, 1.f, p->bn_w[1], gr->bn_w[1],
                                   gr->bn_b[1], gr->conv2_b, st, o));
        CP_CUDA(cudaMemsetAsync(gr->conv2_w, 0, sizeof(float) * 64 * 64 * 9, st));
        CP_TRY((launch_wgrad<64, 64, true>(w.G1, 64, 64, w.A1, 64, 192, R12, w.wpart, gr->conv2_w, 2, st)));
        CP_TRY((launch_nt<128, 64, 0, true>(w.G1, R12, 192, 64, w.Wc2d, 64, 192, nullptr, w.G0, 64, nullptr, nullptr, 0, st)));
    }
    // conv1 block: BN backward + ReLU backward + weight / bias gradients in two passes over (G0, x); the
    // pre-activation gradient is never written (the first layer has no data gradient)
    const int P1 = (int)cp_cdiv(n, C1_WIN);
    float* c1part = w.ppart + (size_t)Pp * CP_EMB_DIM * 512;
    conv1_bn_bwd_reduce_kernel<<<P1, 256, 0, st>>>(w.G0, w.X0, n, p->conv1_w, p->conv1_b, w.mean[0], w.istd[0], w.pa, w.pb);
    CP_CHECK_LAUNCH();
    bn_bwd_finalize_kernel<<<dim3(F_CONV / 32, rp_slabs(P1)), 1024, 0, st>>>(w.pa, w.pb, P1, F_CONV, R12, w.m1, w.m2, gr->bn_w[0],
                                                                         gr->bn_b[0], w.rscratch, w.tickets,
                                                                         o->allreduce ? w.totals : nullptr);
    CP_CHECK_LAUNCH();
    if (o->allreduce) {
        CP_TRY(sync_totals(w, F_CONV, o, st));
        bn_bwd_means_totals_kernel<<<1, 512, 0, st>>>(w.totals, F_CONV, w.m1, w.m2);
        CP_CHECK_LAUNCH();
    }
    conv1_bn_bwd_apply_kernel<<<P1, 256, 0, st>>>(w.G0, w.X0, n, p->conv1_w, p->conv1_b, w.mean[0], w.istd[0], p->bn_w[0],
                                                  w.m1, w.m2, w.pa, c1part);
    CP_CHECK_LAUNCH();
    colsum_finalize_kernel<<<F_CONV / 32, 1024, 0, st>>>(w.pa, P1, F_CONV, gr->conv1_b, 0);
    CP_CHECK_LAUNCH();
    CP_CUDA(cudaMemsetAsync(gr->conv1_w, 0, sizeof(float) * 64 * 9, st));
    colsum_finalize_kernel<<<3 * 64 / 32, 1024, 0, st>>>(c1part, P1, 3 * 64, gr->conv1_w, 1);
    CP_CHECK_LAUNCH();
    // join: every weight gradient is complete before the caller's stream continues
    if (join_event) CP_CUDA(cudaStreamWaitEvent(st, join_event, 0));
    return CP_OK;
}

// Debug / parity tap: copy a saved activation out of the workspace.
extern "C" int cp_encoder_read_activation(const void* workspace, size_t workspace_bytes, int64_t n,
                                          const cp_encoder_opts* o, int stage, int which, float* dst,
                                          void* stream) {
    if (!workspace || !dst || n <= 0 || !opts_ok(o) || stage < 0 || stage >= CP_N_BN || which < 0 || which > 1)
        return CP_ERR_ARG;
    if (!o->save_for_backward) return CP_ERR_UNSUPPORTED;
    const Ws w = carve(const_cast<void*>(workspace), n, o);
    if (workspace_bytes < w.bytes) return CP_ERR_WORKSPACE;
    // tensor-core engine: the post-BN slots of all stages but the last hold fp16 planes, not fp32 values
    // ... and the last block's output is fused into the projection, never stored
    if (which == 1 && (o->engine != CP_ENGINE_SIMT || stage == CP_N_BN - 1)) return CP_ERR_UNSUPPORTED;
    if (stage == 0 && which == 0) {
        // the conv1 activation is not stored: recompute it from the saved input and parameter copy
        conv1_fwd_kernel<<<(unsigned)cp_cdiv(n, C1_WIN), 256, 0, (cudaStream_t)stream>>>(w.X0, n, w.c1w, w.c1b, dst,
                                                                                         nullptr, nullptr);
        CP_CHECK_LAUNCH();
        return CP_OK;
    }
    const float* src;
    size_t elems;
    if (stage < 2) {
        src = which == 0 ? (stage == 0 ? w.Y1 : w.Y2) : (stage == 0 ? w.A1 : w.A2);
        elems = (size_t)n * 12 * F_CONV;
    } else {
        src = which == 0 ? w.Y[stage - 2] : w.A[stage - 2];
        elems = (size_t)n * F_FC;
    }
    CP_CUDA(cudaMemcpyAsync(dst, src, elems * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CP_OK;
}

// ------------------------------------------------------------------- layer-level entry points
namespace {
struct LinWs {
    float *pa, *pb, *wpart;
    plane_t *Ah, *Al, *Gh, *Gl, *Wh, *Wl;
    unsigned int* gmax;
    float* gscale;         // [0] scale of the G planes, [1] its inverse
    size_t bytes;
};
LinWs carve_linear(void* base, int64_t M, int N, int K) {
    LinWs w;
    Carver c{reinterpret_cast<char*>(base)};
    const size_t part = (size_t)cp_cdiv(M, 128) * N;
    w.pa = c.take<float>(part);
    w.pb = c.take<float>(part);
    w.wpart = c.take<float>(WPART_ELEMS);
    w.Ah = c.take<plane_t>((size_t)M * K); w.Al = c.take<plane_t>((size_t)M * K);
    w.Gh = c.take<plane_t>((size_t)M * N); w.Gl = c.take<plane_t>((size_t)M * N);
    w.Wh = c.take<plane_t>((size_t)N * K); w.Wl = c.take<plane_t>((size_t)N * K);
    w.gmax = c.take<unsigned int>(4);
    w.gscale = c.take<float>(4);
    w.bytes = c.off;
    return w;
}
int split_planes(const float* x, plane_t* hi, plane_t* lo, size_t elems, cudaStream_t st, float scale = 1.f) {
    if (elems % 4 != 0) return CP_ERR_ARG;
    split_planes_kernel<<<ew_grid((int64_t)(elems / 4)), 256, 0, st>>>(x, hi, lo, (int64_t)(elems / 4), scale);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ W, int N, int K, plane_t* __restrict__ th, plane_t* __restrict__ tl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * K) return;
    plane_t h, l;
    split_f16(__ldg(W + i), h, l);
    const int n = i / K, k = i % K;
    th[(size_t)k * N + n] = h; tl[(size_t)k * N + n] = l;
}
// max |x| (bit pattern of a non-negative float) -> atomicMax into *out (zero-initialised)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, int64_t n, unsigned int* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(x + i)));
    m = warp_max(m);
    if (threadIdx.x % 32 == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
// gradient operand of the layer-level backward: planes of G * S, S = 2^(8 - ceil(log2 max|G|)); gscale = {S, 1/S}
__global__ void __launch_bounds__(256)
split_scaled_kernel(const float* __restrict__ x, plane_t* __restrict__ hi, plane_t* __restrict__ lo, int64_t n4,
                    const unsigned int* __restrict__ gmax, float* __restrict__ gscale) {
    const float mx = __uint_as_float(__ldg(gmax));
    const float S = (mx > 0.f && mx < 3.0e38f) ? exp2f(8.f - ceilf(log2f(mx))) : 1.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) { gscale[0] = S; gscale[1] = 1.f / S; }
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n4; v += (int64_t)gridDim.x * blockDim.x) {
        float4 a = __ldg(reinterpret_cast<const float4*>(x) + v);
        a.x *= S; a.y *= S; a.z *= S; a.w *= S;
        split_store4(a, hi, lo, v);
    }
}
}  // namespace

extern "C" size_t cp_linear_workspace_bytes(int64_t M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return carve_linear(nullptr, M, N, K).bytes;
}

extern "C" int cp_linear_forward(const float* A, const float* W, const float* bias, float* Y, int64_t M,
                                 int N, int K, int relu, float* col_sum, float* col_sqsum, void* workspace,
                                 size_t workspace_bytes, int engine, void* stream) {
    if (!A || !W || !Y || M <= 0 || N <= 0 || K <= 0) return CP_ERR_ARG;
    if (engine != CP_ENGINE_SIMT && engine != CP_ENGINE_TC) return CP_ERR_UNSUPPORTED;
    if ((col_sum || col_sqsum) && (!col_sum || !col_sqsum)) return CP_ERR_ARG;
    if ((col_sum || engine == CP_ENGINE_TC) && !workspace) return CP_ERR_ARG;
    const LinWs w = carve_linear(workspace, M, N, K);
    if (workspace && workspace_bytes < w.bytes) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float* pa = col_sum ? w.pa : nullptr;
    float* pb = col_sum ? w.pb : nullptr;
    if (engine == CP_ENGINE_TC) {
        CP_TRY(split_planes(A, w.Ah, w.Al, (size_t)M * K, st));
        CP_TRY(split_planes(W, w.Wh, w.Wl, (size_t)N * K, st));
        CP_TRY(tcg::launch_nt(w.Ah, w.Al, M, K, K, w.Wh, w.Wl, N, K, bias, Y, N, pa, pb, relu, st));
    } else if (N % 128 == 0) {
        CP_TRY((launch_nt<128, 128, 0, false>(A, M, K, K, W, N, K, bias, Y, N, pa, pb, relu, st)));
    } else {
        CP_TRY((launch_nt<128, 64, 0, false>(A, M, K, K, W, N, K, bias, Y, N, pa, pb, relu, st)));
    }
    if (col_sum) {
        const int P = (int)cp_cdiv(M, 128);
        colsum_finalize_kernel<<<(N + 31) / 32, 1024, 0, st>>>(pa, P, N, col_sum, 0);
        CP_CHECK_LAUNCH();
        colsum_finalize_kernel<<<(N + 31) / 32, 1024, 0, st>>>(pb, P, N, col_sqsum, 0);
        CP_CHECK_LAUNCH();
    }
    return CP_OK;
}

// fp16 (hi, lo) planes of a fp32 array (n multiple of 4): the operand format of the tensor-core engine
extern "C" int cp_split_planes(const float* x, uint16_t* hi, uint16_t* lo, int64_t n, void* stream) {
    if (!x || !hi || !lo || n < 0) return CP_ERR_ARG;
    if (n == 0) return CP_OK;
    return split_planes(x, reinterpret_cast<plane_t*>(hi), reinterpret_cast<plane_t*>(lo), (size_t)n, (cudaStream_t)stream);
}

// cp_linear_forward on operands that are already split: exactly the launch the encoder issues per
// linear layer (TMA-fed tcgen05 main loop + bias/ReLU/statistics epilogue)
extern "C" int cp_linear_forward_planes(const uint16_t* A_hi_, const uint16_t* A_lo_, const uint16_t* W_hi_,
                                        const uint16_t* W_lo_, const float* bias, float* Y, int64_t M, int N, int K, int relu,
                                        float* col_sum, float* col_sqsum, void* workspace, size_t workspace_bytes,
                                        void* stream) {
    const plane_t *A_hi = reinterpret_cast<const plane_t*>(A_hi_), *A_lo = reinterpret_cast<const plane_t*>(A_lo_);
    const plane_t *W_hi = reinterpret_cast<const plane_t*>(W_hi_), *W_lo = reinterpret_cast<const plane_t*>(W_lo_);
    // both lo planes NULL: the single-product engine (CP_ENGINE_TC_FP16) on the hi planes alone
    const int fast = !A_lo && !W_lo;
    if (!A_hi || !W_hi || (!fast && (!A_lo || !W_lo)) || !Y || !workspace || M <= 0 || N <= 0 || K <= 0) return CP_ERR_ARG;
    if ((col_sum == nullptr) != (col_sqsum == nullptr)) return CP_ERR_ARG;
    const LinWs w = carve_linear(workspace, M, N, K);
    if (workspace_bytes < w.bytes) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    CP_TRY(tcg::launch_nt(A_hi, fast ? A_hi : A_lo, M, K, K, W_hi, fast ? W_hi : W_lo, N, K, bias, Y, N, w.pa, w.pb, relu, st,
                          nullptr, fast));
    if (col_sum) {
        const int P = (int)cp_cdiv(M, 128);
        colsum_finalize_kernel<<<(N + 31) / 32, 1024, 0, st>>>(w.pa, P, N, col_sum, 0);
        CP_CHECK_LAUNCH();
        colsum_finalize_kernel<<<(N + 31) / 32, 1024, 0, st>>>(w.pb, P, N, col_sqsum, 0);
        CP_CHECK_LAUNCH();
    }
    return CP_OK;
}

extern "C" int cp_linear_backward(const float* G, const float* A, const float* W, float* dA, float* dW,
                                  float* db, int64_t M, int N, int K, void* workspace, size_t workspace_bytes,
                                  int engine, void* stream) {
    if (!G || !A || !W || !workspace || M <= 0 || N % 128 != 0 || K % 128 != 0) return CP_ERR_ARG;
    if ((size_t)N * K > WPART_ELEMS) return CP_ERR_ARG;
    if (engine != CP_ENGINE_SIMT && engine != CP_ENGINE_TC) return CP_ERR_UNSUPPORTED;
    const LinWs w = carve_linear(workspace, M, N, K);
    if (workspace_bytes < w.bytes) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (engine == CP_ENGINE_TC) {
        if (((size_t)M * N) % 4 != 0) return CP_ERR_ARG;
        CP_CUDA(cudaMemsetAsync(w.gmax, 0, sizeof(unsigned int), st));
        absmax_kernel<<<ew_grid((int64_t)M * N / 4), 256, 0, st>>>(G, (int64_t)M * N, w.gmax);
        CP_CHECK_LAUNCH();
        split_scaled_kernel<<<ew_grid((int64_t)M * N / 4), 256, 0, st>>>(G, w.Gh, w.Gl, (int64_t)M * N / 4, w.gmax, w.gscale);
        CP_CHECK_LAUNCH();
        if (dA) {
            transpose_split_kernel<<<(N * K + 255) / 256, 256, 0, st>>>(W, N, K, w.Wh, w.Wl);
            CP_CHECK_LAUNCH();
            CP_TRY(tcg::launch_nt(w.Gh, w.Gl, M, N, N, w.Wh, w.Wl, K, N, nullptr, dA, K, nullptr, nullptr, 0, st,
                                  w.gscale + 1));
        }
        if (dW) {
            CP_TRY(split_planes(A, w.Ah, w.Al, (size_t)M * K, st));
            CP_TRY(tc_wgrad(w.Gh, w.Gl, N, w.Ah, w.Al, K, M, w.wpart, dW, 0, st, w.gscale + 1));
        }
    } else {
        if (dA) CP_TRY((launch_nt<128, 128, 1, false>(G, M, N, N, W, K, K, nullptr, dA, K, nullptr, nullptr, 0, st)));
        if (dW) CP_TRY((launch_wgrad<128, 128, false>(G, N, N, A, K, K, M, w.wpart, dW, 0, st)));
    }
    if (db) {
        const int P = (int)cp_cdiv(M, 128);
        colsum_rows_kernel<<<dim3((N + 31) / 32, P), 256, 0, st>>>(G, M, N, w.pa);
        CP_CHECK_LAUNCH();
        colsum_finalize_kernel<<<(N + 31) / 32, 1024, 0, st>>>(w.pa, P, N, db, 0);
        CP_CHECK_LAUNCH();
    }
    return CP_OK;
}

#include "tower.cuh"
#include "cls.cuh"
