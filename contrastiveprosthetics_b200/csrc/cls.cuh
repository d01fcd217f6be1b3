// --prediction mode: the classifier head on the 512-wide trunk output (models.py:300-309) and its loss
// (Model.forward's prediction branch, models.py:113-119; prediction_loss, models.py:175-196), forward + backward.
// Included at the end of encoder.cu (shares its fp32 GEMM launchers and BatchNorm kernels).
//
//   a7 (n,512) -> Linear(512->128) -> ReLU -> BN(128) -> Linear(128->41, no bias) -> z (n,41)
//   features = z / ||z||_2                                   (models.py:118; what Model.forward returns)
//   loss = mean_r CE(features_r, label_r)                    (F.cross_entropy, models.py:187)
//   pred_r = argmax features_r (first maximum)               (models.py:189)
// The head is 3 % of the trunk's work per window: fp32 FFMA GEMMs; the 41-wide logits are computed on a
// zero-padded 64-row copy of the last weight so that the 64-wide tile kernels apply unchanged (exact).
#pragma once

namespace {

constexpr int CH = CP_CLS_HIDDEN;            // 128
constexpr int CPAD = 64;                     // padded number of classes
constexpr size_t CWPART_ELEMS = (size_t)160 * CH * F_FC;

struct CWs {
    float *Y1, *A8;                          // relu(linear1), bn(.)            (n,128)
    float *Z, *DZ;                           // padded logits / their gradient  (n,64)
    float *G0, *G1;                          // (n,128)
    float *mean, *istd, *scale, *shift, *pa, *pb, *m1, *m2;
    double* rscratch;
    double* lpart;                           // per-CTA loss partials
    unsigned int* tickets;
    int* ncor;
    float *W2p, *dW2p, *wpart;
    size_t bytes;
};

constexpr int CE_ROWS = 256;                 // rows per CTA of the loss kernel

CWs cls_carve(void* base, int64_t n) {
    CWs w;
    Carver c{reinterpret_cast<char*>(base)};
    const size_t he = (size_t)n * CH, ze = (size_t)n * CPAD;
    w.Y1 = c.take<float>(he); w.A8 = c.take<float>(he);
    w.Z = c.take<float>(ze); w.DZ = c.take<float>(ze);
    w.G0 = c.take<float>(he); w.G1 = c.take<float>(he);
    w.mean = c.take<float>(CH); w.istd = c.take<float>(CH); w.scale = c.take<float>(CH); w.shift = c.take<float>(CH);
    const size_t pr = (size_t)cp_cdiv(n, 128) + 8;
    w.pa = c.take<float>(pr * CH); w.pb = c.take<float>(pr * CH);
    w.m1 = c.take<float>(CH); w.m2 = c.take<float>(CH);
    w.rscratch = c.take<double>((size_t)RP_SLABS * 2 * CH);
    w.lpart = c.take<double>((size_t)cp_cdiv(n, CE_ROWS));
    w.tickets = c.take<unsigned int>(64);
    w.ncor = c.take<int>(4);
    w.W2p = c.take<float>((size_t)CPAD * CH);
    w.dW2p = c.take<float>((size_t)CPAD * CH);
    w.wpart = c.take<float>(CWPART_ELEMS);
    w.bytes = c.off;
    return w;
}

// dst[r, :] = r < rows_src ? src[r, :] : 0      (row padding of the (41,128) weight to (64,128), and back)
__global__ void __launch_bounds__(256)
copy_rows_kernel(const float* __restrict__ src, int rows_src, float* __restrict__ dst, int rows_dst, int width) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_dst * width) return;
    dst[i] = i / width < rows_src ? __ldg(src + i) : 0.f;
}

// One thread per row: f = z/||z|| (stored to `features`, (n,41)), CE(f, label), first-max argmax; with dz != null the
// gradient of the MEAN loss w.r.t. the padded logits:  df = (softmax(f) - onehot)/n,  dz = (df - f (f . df)) / ||z||.
__global__ void __launch_bounds__(CE_ROWS)
cls_loss_kernel(const float* __restrict__ z, const int64_t* __restrict__ labels, int64_t n, float* __restrict__ features,
                float* __restrict__ dz, int32_t* __restrict__ pred, double* __restrict__ lpart, int* __restrict__ ncor) {
    __shared__ double red[CE_ROWS / 32];
    __shared__ int cred[CE_ROWS / 32];
    const int64_t r = (int64_t)blockIdx.x * CE_ROWS + threadIdx.x;
    double loss = 0.0;
    int correct = 0;
    if (r < n) {
        float f[CP_TASKS];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CP_TASKS; ++j) { f[j] = __ldg(z + r * CPAD + j); s = fmaf(f[j], f[j], s); }
        const float nrm = sqrtf(s);
        float m = -INFINITY;
        int am = 0;
#pragma unroll
        for (int j = 0; j < CP_TASKS; ++j) {
            f[j] = f[j] / nrm;
            if (f[j] > m) { m = f[j]; am = j; }
        }
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < CP_TASKS; ++j) se += expf(f[j] - m);
        const float lse = m + logf(se);
        const int lab = (int)__ldg(labels + r);
        const bool lab_ok = lab >= 0 && lab < CP_TASKS;
        float fl = 0.f;
#pragma unroll
        for (int j = 0; j < CP_TASKS; ++j) fl = j == lab ? f[j] : fl;
        loss = lab_ok ? (double)(lse - fl) : 0.0;
        correct = lab_ok && am == lab;
        if (pred) pred[r] = am;
        if (features) {
#pragma unroll
            for (int j = 0; j < CP_TASKS; ++j) features[r * CP_TASKS + j] = f[j];
        }
        if (dz) {
            const float inv_n = 1.f / (float)n;
            float df[CP_TASKS], dot = 0.f;
#pragma unroll
            for (int j = 0; j < CP_TASKS; ++j) {
                df[j] = lab_ok ? (expf(f[j] - lse) - (j == lab ? 1.f : 0.f)) * inv_n : 0.f;
                dot = fmaf(f[j], df[j], dot);
            }
#pragma unroll
            for (int j = 0; j < CP_TASKS; ++j) dz[r * CPAD + j] = (df[j] - f[j] * dot) / nrm;
#pragma unroll
            for (int j = CP_TASKS; j < CPAD; ++j) dz[r * CPAD + j] = 0.f;
        }
    }
    loss = warp_sum(loss);
    correct = __reduce_add_sync(0xffffffffu, correct);
    if (threadIdx.x % 32 == 0) { red[threadIdx.x / 32] = loss; cred[threadIdx.x / 32] = correct; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        int c = 0;
        for (int i = 0; i < CE_ROWS / 32; ++i) { t += red[i]; c += cred[i]; }
        lpart[blockIdx.x] = t;
        if (c) atomicAdd(ncor, c);                    // integer: order-independent
    }
}

__global__ void cls_loss_finish_kernel(const double* __restrict__ lpart, int P, int64_t n, float* __restrict__ loss,
                                       const int* __restrict__ ncor, int32_t* __restrict__ n_correct) {
    double t = 0.0;
    for (int i = 0; i < P; ++i) t += lpart[i];
    if (loss) *loss = (float)(t / (double)n);
    if (n_correct) *n_correct = *ncor;
}

}  // namespace

extern "C" size_t cp_cls_workspace_bytes(int64_t n) {
    if (n <= 0) return 0;
    return cls_carve(nullptr, n).bytes;
}

extern "C" int cp_cls_forward_backward(const cp_cls_tensors* p, const float* a7, const int64_t* labels, int64_t n,
                                       int bn_mode, float bn_momentum, float bn_eps, float* features, float* loss,
                                       int32_t* pred, int32_t* n_correct, float* d_a7, const cp_cls_tensors* gr,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    if (!p || !a7 || !labels || !workspace || n <= 0 || bn_mode < 0 || bn_mode > 2 || !(bn_eps > 0.f)) return CP_ERR_ARG;
    if (!p->w1 || !p->b1 || !p->bn_w || !p->bn_b || !p->w2) return CP_ERR_ARG;
    if (bn_mode != CP_BN_BATCH && (!p->bn_rm || !p->bn_rv)) return CP_ERR_ARG;
    if ((d_a7 != nullptr) != (gr != nullptr)) return CP_ERR_ARG;
    if (gr && bn_mode == CP_BN_RUNNING) return CP_ERR_UNSUPPORTED;
    if (((uintptr_t)workspace) % 256 != 0) return CP_ERR_ARG;
    const CWs w = cls_carve(workspace, n);
    if (workspace_bytes < w.bytes) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int P = (int)cp_cdiv(n, 128);

    CP_CUDA(cudaMemsetAsync(w.tickets, 0, 64 * sizeof(unsigned int), st));
    CP_CUDA(cudaMemsetAsync(w.ncor, 0, 4 * sizeof(int), st));
    copy_rows_kernel<<<(CPAD * CH + 255) / 256, 256, 0, st>>>(p->w2, CP_TASKS, w.W2p, CPAD, CH);
    CP_CHECK_LAUNCH();
    // Linear(512 -> 128) + ReLU with the BatchNorm statistics in the epilogue, BN, Linear(128 -> 41)
    CP_TRY((launch_nt<128, 128, 0, false>(a7, n, F_FC, F_FC, p->w1, CH, F_FC, p->b1, w.Y1, CH, w.pa, w.pb, 1, st)));
    bn_finalize_kernel<<<dim3(CH / 32, bn_mode == CP_BN_RUNNING ? 1 : rp_slabs(P)), 1024, 0, st>>>(
        w.pa, w.pb, P, CH, n, p->bn_w, p->bn_b, p->bn_rm, p->bn_rv, bn_mode, bn_momentum, bn_eps, w.mean, w.istd, w.scale,
        w.shift, w.rscratch, w.tickets);
    CP_CHECK_LAUNCH();
    bn_apply_kernel<CH, false><<<ew_grid(n * (CH / 4)), 256, 0, st>>>(w.Y1, w.A8, nullptr, n, w.scale, w.shift, nullptr, 1.f,
                                                                      0.f, 0, 0);
    CP_CHECK_LAUNCH();
    CP_TRY((launch_nt<128, 64, 0, false>(w.A8, n, CH, CH, w.W2p, CPAD, CH, nullptr, w.Z, CPAD, nullptr, nullptr, 0, st)));
    const int Pl = (int)cp_cdiv(n, CE_ROWS);
    cls_loss_kernel<<<Pl, CE_ROWS, 0, st>>>(w.Z, labels, n, features, gr ? w.DZ : nullptr, pred, w.lpart, w.ncor);
    CP_CHECK_LAUNCH();
    cls_loss_finish_kernel<<<1, 1, 0, st>>>(w.lpart, Pl, n, loss, w.ncor, n_correct);
    CP_CHECK_LAUNCH();
    if (!gr) return CP_OK;

    // backward: dW2 = dZ^T A8, dA8 = dZ W2, BN + ReLU backward, dW1 = G1^T a7, d_a7 = G1 W1
    CP_TRY((launch_wgrad<64, 64, false>(w.DZ, CPAD, CPAD, w.A8, CH, CH, n, w.wpart, w.dW2p, 0, st, nullptr, nullptr,
                                        CWPART_ELEMS)));
    copy_rows_kernel<<<(CP_TASKS * CH + 255) / 256, 256, 0, st>>>(w.dW2p, CP_TASKS, gr->w2, CP_TASKS, CH);
    CP_CHECK_LAUNCH();
    CP_TRY((launch_nt<128, 128, 1, false>(w.DZ, n, CPAD, CPAD, w.W2p, CH, CH, nullptr, w.G0, CH, nullptr, nullptr, 0, st)));
    const int Pb = (int)cp_cdiv(n, ColMap<CH>::ROWS);
    bn_bwd_reduce_kernel<CH><<<Pb, 256, 0, st>>>(w.G0, w.Y1, n, nullptr, 1.f, w.mean, w.istd, w.pa, w.pb);
    CP_CHECK_LAUNCH();
    bn_bwd_finalize_kernel<<<dim3(CH / 32, rp_slabs(Pb)), 1024, 0, st>>>(w.pa, w.pb, Pb, CH, n, w.m1, w.m2, gr->bn_w, gr->bn_b,
                                                                     w.rscratch, w.tickets);
    CP_CHECK_LAUNCH();
    bn_bwd_apply_kernel<CH, false><<<Pb, 256, 0, st>>>(w.G0, w.Y1, n, nullptr, 1.f, w.mean, w.istd, p->bn_w, w.m1, w.m2,
                                                       w.G1, nullptr, w.pa);
    CP_CHECK_LAUNCH();
    colsum_finalize_kernel<<<CH / 32, 1024, 0, st>>>(w.pa, Pb, CH, gr->b1, 0);
    CP_CHECK_LAUNCH();
    CP_TRY((launch_wgrad<128, 128, false>(w.G1, CH, CH, a7, F_FC, F_FC, n, w.wpart, gr->w1, 0, st, nullptr, nullptr,
                                          CWPART_ELEMS)));
    CP_TRY((launch_nt<128, 128, 1, false>(w.G1, n, CH, CH, p->w1, F_FC, F_FC, nullptr, d_a7, F_FC, nullptr, nullptr, 0, st)));
    return CP_OK;
}
