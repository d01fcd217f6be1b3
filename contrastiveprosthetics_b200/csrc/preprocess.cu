// Offline sEMG preprocessing of (subject, stimulus, repetition) segments -- SURVEY.md section 8(f) row 4:
//   load.py:85-101   x = raw[:2010] * 2^10 -> band-pass -> moving RMS -> emg_[time_mask]
//   utils.py:134-147 filter: Butterworth order-4 band-pass 20-450 Hz, scipy lfilter per channel, result rounded to
//                    the input dtype (float32 for the NinaPro .mat files)
//   utils.py:151-156 moving_rms: sqrt(uniform_filter1d(x^2, 11, mode='nearest'))[5:-5]
// One thread per (segment, channel), sequential in time like scipy's own two C loops, and BIT-EXACT with them: the
// IIR runs in double with explicitly un-fused multiplies / adds in lfilter's direct-form-II-transposed order
//   y = z0 + b0*x;  z[n] = (z[n+1] + x*b[n+1]) - y*a[n+1];  z[last] = x*b[last] - y*a[last]
// and the moving average is uniform_filter1d's running double sum over the edge-replicated line
//   tmp = sum_{k<size} e[k];  out[0] = tmp/size;  tmp += e[l+size-1] - e[l-1];  out[l] = tmp/size.
// Both filters are causal / local, so only the first max(time_idx) + 1 + 2*edge samples of a segment are touched
// (the reference's uint8 time_mask wraps at 256: 263 of the 2010 samples).  rms[0 .. n_rms) goes to a scratch row
// of the thread, which then gathers out[j] = rms[time_idx[j]] (indices may repeat / be unordered).
#include "common.cuh"

#define PP_MAX_COEF 17            // Butterworth band-pass of order <= 8
#define PP_MAX_WINDOW 33
struct PpArgs {
    double b[PP_MAX_COEF], a[PP_MAX_COEF];
    int n_coef;
    float gain;
    int window, edge;
    int64_t n_seg;
    int seg_len, n_ch, n_rms, n_out;
};

__global__ void __launch_bounds__(128)
emg_preprocess_kernel(const float* __restrict__ raw, const PpArgs g, const int32_t* __restrict__ time_idx,
                      float* __restrict__ scratch, float* __restrict__ out) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (tid >= g.n_seg * g.n_ch) return;
    const int64_t seg = tid / g.n_ch;
    const int c = (int)(tid % g.n_ch);
    const float* x = raw + seg * (int64_t)g.seg_len * g.n_ch + c;
    float* rms = scratch + tid * (int64_t)g.n_rms;
    const int nz = g.n_coef - 1, W = g.window, E = g.edge;
    double z[PP_MAX_COEF - 1];
#pragma unroll
    for (int n = 0; n < PP_MAX_COEF - 1; ++n) z[n] = 0.0;
    float ring[PP_MAX_WINDOW];                      // squares e[l-1 .. l+W-1) of the edge-replicated line
    const double size = (double)W;
    double tmp = 0.0;
    const int need = g.n_rms + 2 * E;               // filtered samples 0 .. need-1 enter rms[0 .. n_rms)
    // Extended line e[k] = sq[clamp(k - h, 0, L-1)], h = W/2.  Output position l (0-based over the full line) uses
    // e[l .. l+W-1]; rms[j] = sqrt(out[j + E]).  Samples are produced in order; e-index of sample t is t + h.
    const int h = W / 2;
    int produced_e = 0;                             // number of extended-line entries fed into the running sum so far
    int l = -1;                                     // last output position written
    for (int t = 0; t < need; ++t) {
        const double xv = (double)(__fmul_rn(__ldg(x + (int64_t)t * g.n_ch), g.gain));
        const double yv = __dadd_rn(z[0], __dmul_rn(g.b[0], xv));
#pragma unroll
        for (int n = 0; n < PP_MAX_COEF - 1; ++n) {               // static indices: z stays in registers
            if (n < nz - 1)
                z[n] = __dsub_rn(__dadd_rn(z[n + 1 < PP_MAX_COEF - 1 ? n + 1 : n], __dmul_rn(xv, g.b[n + 1])),
                                 __dmul_rn(yv, g.a[n + 1]));
            else if (n == nz - 1)
                z[n] = __dsub_rn(__dmul_rn(xv, g.b[n + 1]), __dmul_rn(yv, g.a[n + 1]));
        }
        const float f = (float)yv;                  // stored back into the float32 input array (utils.py:146)
        const float sq = __fmul_rn(f, f);
        // feed e-entries: sample 0 also supplies the h replicated left-edge entries
        const int reps = (t == 0) ? h + 1 : 1;
        for (int r = 0; r < reps; ++r) {
            const int k = produced_e++;             // extended index of this entry
            if (k < W) {                            // still filling the first window
                ring[k % W] = sq;
                tmp = __dadd_rn(tmp, (double)sq);
                if (k == W - 1) {
                    l = 0;
                    const float u = (float)__ddiv_rn(tmp, size);
                    if (l >= E && l - E < g.n_rms) rms[l - E] = __fsqrt_rn(u);
                }
            } else {
                // entry k enters, entry k - W leaves: output position l = k - W + 1
                const float leaving = ring[k % W];
                ring[k % W] = sq;
                tmp = __dadd_rn(tmp, __dsub_rn((double)sq, (double)leaving));
                l = k - W + 1;
                const float u = (float)__ddiv_rn(tmp, size);
                if (l >= E && l - E < g.n_rms) rms[l - E] = __fsqrt_rn(u);
            }
        }
    }
    for (int j = 0; j < g.n_out; ++j)
        out[(seg * g.n_out + j) * g.n_ch + c] = rms[__ldg(time_idx + j)];
}

extern "C" size_t cp_emg_preprocess_scratch_elems(int64_t n_seg, int n_ch, int n_rms) {
    if (n_seg < 0 || n_ch <= 0 || n_rms <= 0) return 0;
    return (size_t)n_seg * n_ch * n_rms;
}

extern "C" int cp_emg_preprocess(const float* raw, int64_t n_seg, int seg_len, int n_ch, const double* b,
                                 const double* a, int n_coef, float gain, int rms_window, int n_rms,
                                 const int32_t* time_idx, int n_out, float* out, float* scratch,
                                 size_t scratch_elems, void* stream) {
    if (n_seg == 0) return CP_OK;
    if (!raw || !b || !a || !time_idx || !out || !scratch || n_seg < 0 || n_ch <= 0 || n_out <= 0 || n_rms <= 0)
        return CP_ERR_ARG;
    if (n_coef < 2 || n_coef > PP_MAX_COEF || rms_window < 1 || rms_window > PP_MAX_WINDOW || rms_window % 2 == 0)
        return CP_ERR_ARG;
    if (a[0] != 1.0) return CP_ERR_ARG;                       // scipy normalises by a[0]; butter() returns a[0] == 1
    const int edge = rms_window / 2;
    if (n_rms + 2 * edge > seg_len) return CP_ERR_ARG;        // rms positions beyond the trimmed line
    if (scratch_elems < cp_emg_preprocess_scratch_elems(n_seg, n_ch, n_rms)) return CP_ERR_WORKSPACE;
    PpArgs g;
    for (int i = 0; i < PP_MAX_COEF; ++i) {
        g.b[i] = i < n_coef ? b[i] : 0.0;
        g.a[i] = i < n_coef ? a[i] : 0.0;
    }
    g.n_coef = n_coef; g.gain = gain; g.window = rms_window; g.edge = edge;
    g.n_seg = n_seg; g.seg_len = seg_len; g.n_ch = n_ch; g.n_rms = n_rms; g.n_out = n_out;
    const int64_t threads = n_seg * n_ch;
    emg_preprocess_kernel<<<(unsigned)cp_cdiv(threads, 128), 128, 0, (cudaStream_t)stream>>>(raw, g, time_idx, scratch, out);
    CP_CHECK_LAUNCH();
    return CP_OK;
}
