// Glove-angle tower of the batch x batch variant (BASELINE.json config 5): forward / backward.
// Included at the end of encoder.cu (shares its GEMM launchers and BatchNorm kernels).
//
// Restates the tower the reference keeps commented out (models.py:384-429):
//   glove (n, glove_dim) -> Linear(glove_dim -> 256, no bias) -> BN -> ReLU
//                        -> 3 x [Linear(256 -> 256) -> ReLU -> BN -> Dropout] -> Linear(256 -> 16, no bias)
// BatchNorm uses batch statistics (AdaBN, models.py:17-25).  The input is zero-padded to 64 columns so the 128 x 128 /
// 64 x 64 tile kernels apply unchanged (exact: the padded products are zeros).
// Round 2: the three 256 x 256 blocks (forward, data gradient, weight gradient) run on the tensor-core engine of
// gemm_tc.cuh -- tcgen05 CTA-pair GEMMs on the 3-product fp16 split, fp32-level accuracy -- with the encoder's plane
// conventions (BN-apply / BN-backward write (hi, lo) planes with per-tensor power-of-two scales that the consumers'
// epilogues undo); block 0 (K = glove_dim, 3 % of the tower's FLOPs) and the 256 -> 16 projection stay fp32 FFMA.
// The FFMA GEMMs were 2.1 ms of the 18 ms config-5 step.  CP_GLOVE_TC=0 in the environment keeps the FFMA path.
#pragma once

namespace {

constexpr int GH = CP_GLOVE_HIDDEN;          // 256
constexpr int GPAD = 64;                     // padded input width
constexpr size_t GWPART_ELEMS = (size_t)160 * GH * GH;

struct GWs {
    float *X0;                               // (n, 64) padded input
    float *Z0, *A0;                          // block 0: pre-BN linear output, relu(bn(.))
    float *Y[CP_GLOVE_BLOCKS], *A[CP_GLOVE_BLOCKS];
    uint8_t* keep[CP_GLOVE_BLOCKS];
    float *G0, *G1;
    float *mean[CP_GLOVE_BLOCKS + 1], *istd[CP_GLOVE_BLOCKS + 1], *scale[CP_GLOVE_BLOCKS + 1], *shift[CP_GLOVE_BLOCKS + 1];
    float *pa, *pb, *m1, *m2;
    double* rscratch;
    unsigned int* tickets;
    float *wpart, *ppart, *W0p, *dW0p;
    // tensor-core path (blocks 1..3)
    plane_t* A0p;                            // (hi, lo) planes of A0 (the fp32 A0 stays: ReLU mask of block 0's backward)
    plane_t *Wh[CP_GLOVE_BLOCKS], *Wl[CP_GLOVE_BLOCKS], *Wth[CP_GLOVE_BLOCKS], *Wtl[CP_GLOVE_BLOCKS];
    unsigned int* tcu;                       // [0..3] abound per stage, [4..6] wmax per block, [8..11] gmax, [12..15] g1max (bit patterns)
    float* tcf;                              // [0..3] ascale_inv per stage, [4..6] wscale_inv per block, [8..11] gscale_inv per stage
    size_t bytes;
};
const bool g_glove_tc = []() { const char* e = getenv("CP_GLOVE_TC"); return !(e && e[0] == '0'); }();

GWs glove_carve(void* base, int64_t n, const cp_glove_opts* o) {
    GWs w;
    Carver c{reinterpret_cast<char*>(base)};
    const bool save = o->save_for_backward != 0;
    const size_t he = (size_t)n * GH;
    w.X0 = c.take<float>((size_t)n * GPAD);
    w.Z0 = c.take<float>(he);
    w.A0 = c.take<float>(he);
    if (save) {
        for (int b = 0; b < CP_GLOVE_BLOCKS; ++b) { w.Y[b] = c.take<float>(he); w.A[b] = c.take<float>(he); }
        w.G0 = c.take<float>(he);
        w.G1 = c.take<float>(he);
    } else {
        float* y = c.take<float>(he);
        float* a = c.take<float>(he);
        for (int b = 0; b < CP_GLOVE_BLOCKS; ++b) { w.Y[b] = y; w.A[b] = (b & 1) ? w.A0 : a; }
        w.G0 = w.G1 = nullptr;
    }
    for (int b = 0; b < CP_GLOVE_BLOCKS; ++b) w.keep[b] = o->dropout_p > 0.f ? c.take<uint8_t>(he) : nullptr;
    for (int l = 0; l < CP_GLOVE_BLOCKS + 1; ++l) {
        w.mean[l] = c.take<float>(GH); w.istd[l] = c.take<float>(GH);
        w.scale[l] = c.take<float>(GH); w.shift[l] = c.take<float>(GH);
    }
    const size_t pr = (size_t)cp_cdiv(n, 128) + 8;
    w.pa = c.take<float>(pr * GH);
    w.pb = c.take<float>(pr * GH);
    w.m1 = c.take<float>(GH);
    w.m2 = c.take<float>(GH);
    w.rscratch = c.take<double>((size_t)RP_SLABS * 2 * GH);
    w.tickets = c.take<unsigned int>(64);
    w.wpart = save ? c.take<float>(GWPART_ELEMS) : nullptr;
    w.ppart = save ? c.take<float>((size_t)cp_cdiv(n, PROJ_W_ROWS) * CP_EMB_DIM * GH) : nullptr;
    w.W0p = c.take<float>((size_t)GH * GPAD);
    w.dW0p = save ? c.take<float>((size_t)GH * GPAD) : nullptr;
    w.A0p = c.take<plane_t>(2 * he);
    for (int b = 0; b < CP_GLOVE_BLOCKS; ++b) {
        w.Wh[b] = c.take<plane_t>((size_t)GH * GH); w.Wl[b] = c.take<plane_t>((size_t)GH * GH);
        w.Wth[b] = c.take<plane_t>((size_t)GH * GH); w.Wtl[b] = c.take<plane_t>((size_t)GH * GH);
    }
    w.tcu = c.take<unsigned int>(16);
    w.tcf = c.take<float>(16);
    w.bytes = c.off;
    return w;
}

// dst[r, 0..wd) = c < ws ? src[r, c] : 0
__global__ void __launch_bounds__(256)
copy_cols_kernel(const float* __restrict__ src, int ws, float* __restrict__ dst, int wd, int64_t rows) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= rows * wd) return;
    const int64_t r = i / wd;
    const int c = (int)(i % wd);
    dst[i] = c < ws ? __ldg(src + r * ws + c) : 0.f;
}

// block 0 (Linear -> BN -> ReLU): a = relu(z*scale + shift)
__global__ void __launch_bounds__(256)
bn_relu_apply_kernel(const float* __restrict__ z, float* __restrict__ a, int64_t R, int F,
                     const float* __restrict__ scale, const float* __restrict__ shift) {
    const int64_t total = R * (F / 4);
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(v % (F / 4)) * 4;
        const float4 x = __ldg(reinterpret_cast<const float4*>(z) + v);
        const float4 s = __ldg(reinterpret_cast<const float4*>(scale + c));
        const float4 t = __ldg(reinterpret_cast<const float4*>(shift + c));
        reinterpret_cast<float4*>(a)[v] = make_float4(fmaxf(fmaf(x.x, s.x, t.x), 0.f), fmaxf(fmaf(x.y, s.y, t.y), 0.f),
                                                      fmaxf(fmaf(x.z, s.z, t.z), 0.f), fmaxf(fmaf(x.w, s.w, t.w), 0.f));
    }
}

// the same, also as fp16 (hi, lo) planes of a * S (S from the BatchNorm output bound, like bn_apply_kernel<.., true>)
__global__ void __launch_bounds__(256)
bn_relu_apply_planes_kernel(const float* __restrict__ z, float* __restrict__ a, plane_t* __restrict__ a_hi,
                            plane_t* __restrict__ a_lo, int64_t R, int F, const float* __restrict__ scale,
                            const float* __restrict__ shift, const unsigned int* __restrict__ abound,
                            float* __restrict__ ascale_inv) {
    const int64_t total = R * (F / 4);
    const float S = plane_scale(__uint_as_float(__ldg(abound)));
    if (blockIdx.x == 0 && threadIdx.x == 0) *ascale_inv = 1.f / S;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(v % (F / 4)) * 4;
        const float4 x = __ldg(reinterpret_cast<const float4*>(z) + v);
        const float4 s = __ldg(reinterpret_cast<const float4*>(scale + c));
        const float4 t = __ldg(reinterpret_cast<const float4*>(shift + c));
        const float4 o = make_float4(fmaxf(fmaf(x.x, s.x, t.x), 0.f), fmaxf(fmaf(x.y, s.y, t.y), 0.f),
                                     fmaxf(fmaf(x.z, s.z, t.z), 0.f), fmaxf(fmaf(x.w, s.w, t.w), 0.f));
        reinterpret_cast<float4*>(a)[v] = o;
        split_store4(make_float4(o.x * S, o.y * S, o.z * S, o.w * S), a_hi, a_lo, v);
    }
}

// weight planes of one [N][K] layer: straight (forward B operand) and transposed (data-gradient B operand), 32 x 32 tiles
__global__ void __launch_bounds__(256)
glove_weight_planes_kernel(const float* __restrict__ W, int N, int K, const unsigned int* __restrict__ wmax,
                           plane_t* __restrict__ Wh, plane_t* __restrict__ Wl, plane_t* __restrict__ Wth,
                           plane_t* __restrict__ Wtl, float* __restrict__ wscale_inv) {
    __shared__ uint32_t tile[32][33];
    const int k0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    const float S = weight_scale(wmax);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *wscale_inv = 1.f / S;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int o = o0 + r, k = k0 + tx;
        plane_t h, lo;
        split_f16(__ldg(W + (size_t)o * K + k) * S, h, lo);
        Wh[(size_t)o * K + k] = h;
        Wl[(size_t)o * K + k] = lo;
        tile[r][tx] = (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(lo) << 16);
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, o = o0 + tx;
        const uint32_t v = tile[tx][r];
        Wth[(size_t)k * N + o] = __ushort_as_half((unsigned short)(v & 0xffffu));
        Wtl[(size_t)k * N + o] = __ushort_as_half((unsigned short)(v >> 16));
    }
}

bool glove_opts_ok(const cp_glove_opts* o) {
    return o && o->glove_dim >= 1 && o->glove_dim <= GPAD && o->dropout_p >= 0.f && o->dropout_p < 1.f && o->bn_eps > 0.f;
}

int glove_bn_finalize(const GWs& w, int l, int P, int64_t R, const float* gamma, const float* beta,
                      const cp_glove_opts* o, cudaStream_t st) {
    bn_finalize_kernel<<<dim3(GH / 32, rp_slabs(P)), 1024, 0, st>>>(w.pa, w.pb, P, GH, R, gamma, beta, nullptr, nullptr,
                                                                CP_BN_BATCH, 0.f, o->bn_eps, w.mean[l], w.istd[l],
                                                                w.scale[l], w.shift[l], w.rscratch, w.tickets, nullptr,
                                                                g_glove_tc ? w.tcu + l : nullptr);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

// BN (+ReLU) backward of tower stage l.  post != null: Linear -> BN -> ReLU order (block 0)
int glove_bn_backward(const float* g, const float* y, const float* post, float* gz, int64_t R, const GWs& w, int l,
                      const uint8_t* keep, float inv_keep, const float* gamma, float* d_gamma, float* d_beta,
                      float* d_bias, cudaStream_t st, bool planes = false) {
    const int P = (int)cp_cdiv(R, ColMap<GH>::ROWS);
    bn_bwd_reduce_kernel<GH><<<P, 256, 0, st>>>(g, y, R, keep, inv_keep, w.mean[l], w.istd[l], w.pa, w.pb, post,
                                                planes ? w.tcu + 8 + l : nullptr);
    CP_CHECK_LAUNCH();
    bn_bwd_finalize_kernel<<<dim3(GH / 32, rp_slabs(P)), 1024, 0, st>>>(w.pa, w.pb, P, GH, R, w.m1, w.m2, d_gamma, d_beta,
                                                                     w.rscratch, w.tickets);
    CP_CHECK_LAUNCH();
    if (planes)         // gz as fp16 planes of gz * S inside the fp32 slot (hi first, lo behind it), S from max |g'|
        bn_bwd_apply_kernel<GH, true><<<P, 256, 0, st>>>(g, y, R, keep, inv_keep, w.mean[l], w.istd[l], gamma, w.m1, w.m2,
                                                         gz, reinterpret_cast<float*>(reinterpret_cast<plane_t*>(gz) + (size_t)R * GH),
                                                         w.pa, post, w.tcu + 8 + l, w.tcf + 8 + l, w.tcu + 12 + l);
    else
        bn_bwd_apply_kernel<GH, false><<<P, 256, 0, st>>>(g, y, R, keep, inv_keep, w.mean[l], w.istd[l], gamma, w.m1, w.m2,
                                                          gz, nullptr, w.pa, post);
    CP_CHECK_LAUNCH();
    if (d_bias) {
        colsum_finalize_kernel<<<GH / 32, 1024, 0, st>>>(w.pa, P, GH, d_bias, 0);
        CP_CHECK_LAUNCH();
    }
    return CP_OK;
}

}  // namespace

extern "C" size_t cp_glove_workspace_bytes(int64_t n, const cp_glove_opts* opts) {
    if (n <= 0 || !glove_opts_ok(opts)) return 0;
    return glove_carve(nullptr, n, opts).bytes;
}

extern "C" int cp_glove_forward(const cp_glove_tensors* p, const float* glove, int64_t n, float* emb, void* workspace,
                                size_t workspace_bytes, const cp_glove_opts* o, void* stream) {
    if (!p || !glove || !emb || !workspace || n <= 0 || !glove_opts_ok(o)) return CP_ERR_ARG;
    if (((uintptr_t)workspace) % 256 != 0) return CP_ERR_ARG;
    const GWs w = glove_carve(workspace, n, o);
    if (workspace_bytes < w.bytes) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int P = (int)cp_cdiv(n, 128);

    CP_CUDA(cudaMemsetAsync(w.tickets, 0, 64 * sizeof(unsigned int), st));
    copy_cols_kernel<<<(unsigned)cp_cdiv(n * GPAD, 256), 256, 0, st>>>(glove, o->glove_dim, w.X0, GPAD, n);
    CP_CHECK_LAUNCH();
    copy_cols_kernel<<<(unsigned)cp_cdiv((int64_t)GH * GPAD, 256), 256, 0, st>>>(p->w0, o->glove_dim, w.W0p, GPAD, GH);
    CP_CHECK_LAUNCH();

    const bool tc = g_glove_tc;
    if (tc) {
        CP_CUDA(cudaMemsetAsync(w.tcu, 0, 16 * sizeof(unsigned int), st));
        WmaxArgs wa;
        for (int q = 0; q < 8; ++q) { wa.W[q] = p->w[0]; wa.n[q] = 0; }
        for (int b = 0; b < CP_GLOVE_BLOCKS; ++b) { wa.W[b] = p->w[b]; wa.n[b] = GH * GH; }
        weights_absmax_kernel<<<dim3(48, CP_GLOVE_BLOCKS), 256, 0, st>>>(wa, w.tcu + 4);
        CP_CHECK_LAUNCH();
        for (int b = 0; b < CP_GLOVE_BLOCKS; ++b) {
            glove_weight_planes_kernel<<<dim3(GH / 32, GH / 32), 256, 0, st>>>(p->w[b], GH, GH, w.tcu + 4 + b, w.Wh[b], w.Wl[b],
                                                                               w.Wth[b], w.Wtl[b], w.tcf + 4 + b);
            CP_CHECK_LAUNCH();
        }
    }
    // block 0: Linear (no bias) -> BN -> ReLU
    CP_TRY((launch_nt<128, 128, 0, false>(w.X0, n, GPAD, GPAD, w.W0p, GH, GPAD, nullptr, w.Z0, GH, w.pa, w.pb, 0, st)));
    CP_TRY(glove_bn_finalize(w, 0, P, n, p->bn0_w, p->bn0_b, o, st));
    if (tc)
        bn_relu_apply_planes_kernel<<<ew_grid(n * (GH / 4)), 256, 0, st>>>(w.Z0, w.A0, w.A0p, w.A0p + (size_t)n * GH, n, GH,
                                                                           w.scale[0], w.shift[0], w.tcu + 0, w.tcf + 0);
    else
        bn_relu_apply_kernel<<<ew_grid(n * (GH / 4)), 256, 0, st>>>(w.Z0, w.A0, n, GH, w.scale[0], w.shift[0]);
    CP_CHECK_LAUNCH();

    // blocks 1..3: Linear -> ReLU -> BN -> Dropout
    const float inv_keep = o->dropout_p > 0.f ? 1.f / (1.f - o->dropout_p) : 1.f;
    const float* in = w.A0;
    const plane_t* in_hi = w.A0p;                      // tensor-core path: planes of the block's input
    for (int b = 0; b < CP_GLOVE_BLOCKS; ++b) {
        if (tc)
            CP_TRY(tcg::launch_nt(in_hi, in_hi + (size_t)n * GH, n, GH, GH, w.Wh[b], w.Wl[b], GH, GH, p->b[b], w.Y[b], GH, w.pa,
                                  w.pb, 1, st, w.tcf + b, 0, nullptr, nullptr, 1.f, w.tcf + 4 + b));
        else
            CP_TRY((launch_nt<128, 128, 0, false>(in, n, GH, GH, p->w[b], GH, GH, p->b[b], w.Y[b], GH, w.pa, w.pb, 1, st)));
        CP_TRY(glove_bn_finalize(w, 1 + b, P, n, p->bn_w[b], p->bn_b[b], o, st));
        uint8_t* keep = nullptr;
        float gen_p = 0.f;
        if (o->dropout_p > 0.f) {
            if (o->ext_masks)
                CP_CUDA(cudaMemcpyAsync(w.keep[b], o->ext_masks + (size_t)b * n * GH, (size_t)n * GH,
                                        cudaMemcpyDeviceToDevice, st));
            else
                gen_p = o->dropout_p;
            keep = w.keep[b];
        }
        // the last block feeds the fp32 projection kernel; the others feed the next tensor-core GEMM (planes in the slot)
        if (tc && b + 1 < CP_GLOVE_BLOCKS)
            bn_apply_kernel<GH, true><<<ew_grid(n * (GH / 4)), 256, 0, st>>>(
                w.Y[b], w.A[b], reinterpret_cast<float*>(reinterpret_cast<plane_t*>(w.A[b]) + (size_t)n * GH), n, w.scale[1 + b],
                w.shift[1 + b], keep, inv_keep, gen_p, o->dropout_seed, (uint64_t)(16 + b),
                (const unsigned long long*)o->dropout_step, w.tcu + 1 + b, w.tcf + 1 + b);
        else
            bn_apply_kernel<GH, false><<<ew_grid(n * (GH / 4)), 256, 0, st>>>(w.Y[b], w.A[b], nullptr, n, w.scale[1 + b],
                                                                              w.shift[1 + b], keep, inv_keep, gen_p,
                                                                              o->dropout_seed, (uint64_t)(16 + b),
                                                                              (const unsigned long long*)o->dropout_step);
        CP_CHECK_LAUNCH();
        in = w.A[b];
        in_hi = reinterpret_cast<const plane_t*>(w.A[b]);
    }
    proj_fwd_kernel<GH><<<(unsigned)std::min<int64_t>(cp_cdiv(n, 8 * PROJ_RPW), (int64_t)CP_NUM_SMS * 4), 256, 0, st>>>(
        in, p->proj_w, emb, n);
    CP_CHECK_LAUNCH();
    return CP_OK;
}

extern "C" int cp_glove_backward(const cp_glove_tensors* p, const float* d_emb, int64_t n, const cp_glove_tensors* gr,
                                 void* workspace, size_t workspace_bytes, const cp_glove_opts* o, void* stream) {
    if (!p || !d_emb || !gr || !workspace || n <= 0 || !glove_opts_ok(o)) return CP_ERR_ARG;
    if (!o->save_for_backward) return CP_ERR_UNSUPPORTED;
    const GWs w = glove_carve(workspace, n, o);
    if (workspace_bytes < w.bytes) return CP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const float inv_keep = o->dropout_p > 0.f ? 1.f / (1.f - o->dropout_p) : 1.f;

    // projection
    const int Pp = (int)cp_cdiv(n, PROJ_W_ROWS);
    const float* a_last = w.A[CP_GLOVE_BLOCKS - 1];
    proj_bwd_weight_kernel<GH><<<Pp, 256, 0, st>>>(d_emb, a_last, n, w.ppart);
    CP_CHECK_LAUNCH();
    colsum_finalize_kernel<<<CP_EMB_DIM * GH / 32, 1024, 0, st>>>(w.ppart, Pp, CP_EMB_DIM * GH, gr->proj_w, 0);
    CP_CHECK_LAUNCH();
    proj_bwd_data_kernel<GH><<<(unsigned)std::min<int64_t>(cp_cdiv(n, 16), (int64_t)CP_NUM_SMS * 8), 256, 0, st>>>(
        d_emb, p->proj_w, w.G0, n);
    CP_CHECK_LAUNCH();

    const bool tc = g_glove_tc;
    if (tc) CP_CUDA(cudaMemsetAsync(w.tcu + 8, 0, 8 * sizeof(unsigned int), st));       // gmax / g1max of this call
    for (int b = CP_GLOVE_BLOCKS - 1; b >= 0; --b) {
        const uint8_t* keep = o->dropout_p > 0.f ? w.keep[b] : nullptr;
        CP_TRY(glove_bn_backward(w.G0, w.Y[b], nullptr, w.G1, n, w, 1 + b, keep, inv_keep, p->bn_w[b], gr->bn_w[b],
                                 gr->bn_b[b], gr->b[b], st, tc));
        const float* a_in = b == 0 ? w.A0 : w.A[b - 1];
        if (tc) {
            // G1 and the block's input are (hi, lo) planes: weight gradient and data gradient on the tensor-core engine
            const plane_t* gh = reinterpret_cast<const plane_t*>(w.G1);
            const plane_t* gl = gh + (size_t)n * GH;
            const plane_t* ah = b == 0 ? w.A0p : reinterpret_cast<const plane_t*>(w.A[b - 1]);
            const plane_t* al = ah + (size_t)n * GH;
            CP_TRY(tc_wgrad(gh, gl, GH, ah, al, GH, n, w.wpart, gr->w[b], 0, st, w.tcf + 8 + 1 + b, 0, true, w.tcf + b));
            CP_TRY(tcg::launch_nt(gh, gl, n, GH, GH, w.Wth[b], w.Wtl[b], GH, GH, nullptr, w.G0, GH, nullptr, nullptr, 0, st,
                                  w.tcf + 8 + 1 + b, 0, nullptr, nullptr, 1.f, w.tcf + 4 + b));
            continue;
        }
        CP_TRY((launch_wgrad<128, 128, false>(w.G1, GH, GH, a_in, GH, GH, n, w.wpart, gr->w[b], 0, st, nullptr, nullptr,
                                              GWPART_ELEMS)));
        CP_TRY((launch_nt<128, 128, 1, false>(w.G1, n, GH, GH, p->w[b], GH, GH, nullptr, w.G0, GH, nullptr, nullptr, 0, st)));
    }
    // block 0: ReLU after the BN, no bias, no data gradient
    CP_TRY(glove_bn_backward(w.G0, w.Z0, w.A0, w.G1, n, w, 0, nullptr, 1.f, p->bn0_w, gr->bn0_w, gr->bn0_b, nullptr, st));
    CP_TRY((launch_wgrad<64, 64, false>(w.G1, GH, GH, w.X0, GPAD, GPAD, n, w.wpart, w.dW0p, 0, st, nullptr, nullptr,
                                        GWPART_ELEMS)));
    copy_cols_kernel<<<(unsigned)cp_cdiv((int64_t)GH * o->glove_dim, 256), 256, 0, st>>>(w.dW0p, GPAD, gr->w0, o->glove_dim, GH);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
