/* libcpros -- B200 (sm_100a) kernels for the ContrastiveProsthetics hot path.  C ABI.
 *
 * Every entry point takes raw DEVICE pointers, explicit sizes and a CUDA stream (void* =
 * cudaStream_t), returns an int status (0 ok, <0 argument error, >0 cudaError_t), never
 * allocates or frees caller memory, never throws, never synchronises the device, and is
 * re-entrant per stream: host threads may call concurrently on different streams / devices.  Per-device state of
 * the library (function attributes, the side stream + events cp_encoder_backward forks its weight-gradient GEMMs
 * onto) is created on first use on that device; concurrent cp_encoder_backward calls on ONE device are serialised
 * on the host for the duration of their enqueue only.  There is no CPU implementation behind any symbol.
 *
 * The reference (FibonacciDude/ContrastiveProsthetics) is pure Python and has no FFI; each
 * symbol below cites the reference Python code it replaces (paths under /root/reference/code).
 * The reference-side binding (a ctypes stub + torch.autograd.Function) is in INTEGRATION.md and
 * shipped in contrastiveprosthetics_b200/_lib.py.
 */
#ifndef CPROS_H
#define CPROS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CP_OK 0
#define CP_ERR_ARG (-1)        /* null pointer / bad size */
#define CP_ERR_WORKSPACE (-2)  /* workspace too small */
#define CP_ERR_UNSUPPORTED (-3)
#define CP_ERR_COLLECTIVE (-4) /* the caller's all-reduce callback failed */

#define CP_TASKS 41            /* constants.py:46  MAX_TASKS  */
#define CP_EMG_DIM 12          /* constants.py:97  EMG_DIM    */
#define CP_EMB_DIM 16          /* d_e (train.py:182 des=[16]) */
#define CP_N_BN 9              /* 2 BatchNorm2d + 7 BatchNorm1d (models.py:248-298) */
#define CP_N_FC 7

int cp_version(void);
/* kernels launched by this library in this process so far (host-side counter) */
unsigned long long cp_launch_count(void);
/* static string for a status returned by any entry point */
const char *cp_status_string(int status);

/* ---------------------------------------------------------------- K1: gather (+ normalise)
 * Replaces TaskWrapper.__getitem__ -> DB23.__getitem__/slice_batch (utils.py:51-64,
 * load.py:256-273) + default_collate, and RunningStats.normalize (utils.py:129-130).
 *   dst[r, :] = (src[idx[r], :] - mean[c]) / std[c]          (true divide)
 * src: (src_rows, row_len) fp32 (EMG_use: row_len 12; tensor: row_len 25*12), idx: n_rows int64,
 * channel c = column % n_ch; mean/std: stat_len in {0 (no normalisation), 1 (scalar), n_ch}.
 * Out-of-range indices are reported through *err_flag (device int, may be NULL) and read row 0. */
int cp_gather_norm(const float *src, int64_t src_rows, int row_len, const int64_t *idx,
                   int64_t n_rows, float *dst, const float *mean, const float *std, int stat_len,
                   int n_ch, int *err_flag, void *stream);

/* ---------------------------------------------------------------- K2: EMG encoder
 * Replaces EMGNet.forward up to the projection (models.py:319-323; layers 248-315; BN types
 * 17-35, 238-243) and its autograd backward.  All tensors fp32, PyTorch state-dict layouts. */
typedef struct cp_encoder_tensors {
    float *conv1_w;            /* (64,1,3,3)  emg_net.conv_emg.0.weight */
    float *conv1_b;            /* (64)                                   */
    float *conv2_w;            /* (64,64,3,3) emg_net.conv_emg.3.weight */
    float *conv2_b;            /* (64)                                   */
    float *fc_w[CP_N_FC];      /* (512,768), 6 x (512,512)  emg_net.linear.{0,3,6,9,13,17,21} */
    float *fc_b[CP_N_FC];      /* (512)                                  */
    float *proj_w;             /* (16,512)    emg_net.last.0.weight      */
    float *bn_w[CP_N_BN];      /* gamma: 64,64,512 x 7                   */
    float *bn_b[CP_N_BN];      /* beta                                   */
    float *bn_rm[CP_N_BN];     /* running_mean (stock BN only, else NULL) */
    float *bn_rv[CP_N_BN];     /* running_var                             */
} cp_encoder_tensors;

#define CP_BN_BATCH 0          /* batch statistics (AdaBN always; models.py:22,32) */
#define CP_BN_BATCH_UPDATE 1   /* batch statistics + running-stat update (nn.BatchNorm train) */
#define CP_BN_RUNNING 2        /* running statistics (nn.BatchNorm eval) */

#define CP_ENGINE_SIMT 0       /* fp32 FFMA GEMMs */
#define CP_ENGINE_TC 1         /* tcgen05 GEMMs on a 3-product fp16 split (fp32-level accuracy) */
#define CP_ENGINE_TC_FP16 2    /* tcgen05 GEMMs, ONE fp16 product (11-bit operands like TF32, fp32 accumulate):
                                  the reduced-precision path BASELINE.json allows at 1e-2; not the parity path */

/* SyncBN hook (optional).  Sums `count` doubles at device pointer `buf` over all ranks, in place, ordered on
 * `stream`; returns 0 on success.  A C/C++ host implements it with
 *     ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, comm, (cudaStream_t)stream)
 * (user = the ncclComm_t); the Python host routes it to torch.distributed (NCCL).  The library calls it once
 * per BatchNorm layer in forward (column sums, sums of squares, row count) and once in backward. */
typedef int (*cp_allreduce_fn)(void *user, void *buf, size_t count, void *stream);

typedef struct cp_encoder_opts {
    int32_t bn_mode;
    int32_t engine;
    float bn_momentum;         /* 0.1 for nn.BatchNorm */
    float bn_eps;              /* 1e-5 */
    float dropout_p;           /* dropout after linear blocks 4..7 (models.py:282-297); 0 = off */
    int32_t save_for_backward; /* keep activations in the workspace for cp_encoder_backward */
    uint64_t dropout_seed;     /* Philox key; element stream = (layer, flat index) */
    const uint8_t *ext_masks;  /* optional 4 x (n,512) {0,1} keep masks (parity tests), else NULL */
    const uint64_t *dropout_step; /* optional DEVICE counter mixed into the Philox key at run time: a CUDA-graph replay
                                   * of the same launch then draws a fresh mask when the caller bumps the counter */
    cp_allreduce_fn allreduce; /* NULL: BatchNorm statistics over this rank's rows (local BN).  Non-NULL: SyncBN, */
    void *allreduce_user;      /* statistics over the rows of every rank (global-batch parity)      */
    int32_t trunk_only;        /* --prediction mode (models.py:300-309): stop after the 7th linear block.  `emb` of
                                * cp_encoder_forward is then the (n,512) block output, `d_emb` of cp_encoder_backward
                                * its gradient; proj_w is not used (may be NULL) and its gradient is not written */
    int32_t reserved;
} cp_encoder_opts;

size_t cp_encoder_workspace_bytes(int64_t n_windows, const cp_encoder_opts *opts);

/* x: (n,12) windows -> emb: (n,16).  params: weights (read-only except bn_rm/bn_rv in mode 1). */
int cp_encoder_forward(const cp_encoder_tensors *params, const float *x, int64_t n, float *emb,
                       void *workspace, size_t workspace_bytes, const cp_encoder_opts *opts,
                       void *stream);

/* d_emb: (n,16) -> grads (same layouts as params; bn_rm/bn_rv ignored; every grad tensor is
 * OVERWRITTEN).  Must follow a cp_encoder_forward with save_for_backward on the same workspace. */
int cp_encoder_backward(const cp_encoder_tensors *params, const float *d_emb, int64_t n,
                        const cp_encoder_tensors *grads, void *workspace, size_t workspace_bytes,
                        const cp_encoder_opts *opts, void *stream);

/* Dropout mask generator taps (tests).  cp_philox4x32_10: out[i] = Philox4x32-10(ctr[i] (4 words), key[i] (2 words))
 * -- the raw block function, checked against the published known-answer vectors.  cp_dropout_mask: the {0,1} keep
 * mask of dropout layer `layer` (0..3) exactly as cp_encoder_forward draws it for opts {dropout_seed = seed,
 * dropout_step = step}: element e = row*512 + col keeps iff word (e % 4) of Philox(ctr = {e/4 lo, e/4 hi, layer,
 * 0x43505253}, key = seed [+ *step * 0x9E3779B97F4A7C15]) >= p * 2^32.  n % 4 == 0. */
int cp_philox4x32_10(const uint32_t *ctr, const uint32_t *key, int64_t n, uint32_t *out, void *stream);
int cp_dropout_mask(uint8_t *keep, int64_t n, float p, uint64_t seed, int layer, const uint64_t *step,
                    void *stream);

/* Parity tap: copy the saved activation of BN stage `stage` (0,1: conv stages, layout (n*12,64)
 * position-major/channel-contiguous; 2..8: linear stages, (n,512)) out of a workspace written by
 * cp_encoder_forward(save_for_backward=1).  which = 0: post-ReLU pre-BN, 1: post-BN(/dropout). */
int cp_encoder_read_activation(const void *workspace, size_t workspace_bytes, int64_t n,
                               const cp_encoder_opts *opts, int stage, int which, float *dst,
                               void *stream);

/* Layer-level entry points (unit parity tests; the encoder calls the same kernels).
 * Y = relu?(A[M,K] @ W[N,K]^T + bias) with per-column sum / sum-of-squares (BN statistics). */
int cp_linear_forward(const float *A, const float *W, const float *bias, float *Y, int64_t M, int N,
                      int K, int relu, float *col_sum, float *col_sqsum, void *workspace,
                      size_t workspace_bytes, int engine, void *stream);
/* Tensor-core engine operand format: two fp16 planes, x = hi + lo/2048 with hi = fp16(x),
 * lo = fp16((x - hi) * 2048) (22 significand bits; n % 4 == 0). */
int cp_split_planes(const float *x, uint16_t *hi, uint16_t *lo, int64_t n, void *stream);
/* cp_linear_forward (CP_ENGINE_TC) on pre-split operands: exactly the per-layer launch of the encoder
 * (N % 128 == 0, K % 64 == 0).  col_sum / col_sqsum may both be NULL.  A_lo == W_lo == NULL selects the
 * single-product launch of CP_ENGINE_TC_FP16 (hi planes only). */
int cp_linear_forward_planes(const uint16_t *A_hi, const uint16_t *A_lo, const uint16_t *W_hi, const uint16_t *W_lo,
                             const float *bias, float *Y, int64_t M, int N, int K, int relu,
                             float *col_sum, float *col_sqsum, void *workspace, size_t workspace_bytes,
                             void *stream);
/* dA[M,K] = G[M,N] @ W[N,K];  dW[N,K] = G^T @ A;  db[N] = colsum(G) */
int cp_linear_backward(const float *G, const float *A, const float *W, float *dA, float *dW,
                       float *db, int64_t M, int N, int K, void *workspace, size_t workspace_bytes,
                       int engine, void *stream);
size_t cp_linear_workspace_bytes(int64_t M, int N, int K);

/* ---------------------------------------------------------------- K3: fused contrastive head
 * Replaces Model.forward's contrastive branch (models.py:121-130), GLOVENet.forward's default
 * branch (models.py:457-465) and Model.loss -> contrastive_loopy_loss x2 (models.py:198-208,
 * 132-173, float part).  Group g = (b, w): rows emb[((b*41 + i)*W + w)*16 ...], i = class.
 *   loss = 1/2 (mean row-CE + mean column-CE) of S_g = normalise(emb_g) normalise(table)^T
 * table[j,:] = table_w[:,j] + table_b  (glove_net.easy.0: weight (16,41), bias (16)).
 * Outputs (each may be NULL): loss (1 float), d_emb (same layout as emb), d_table_w (16,41),
 * d_table_b (16) -- gradients of `loss` w.r.t. the un-normalised inputs; pred (B*W,41) int32 =
 * first-max argmax per row; n_correct (B*W) int32 = #(pred == row); logits (B*W,41,41). */
size_t cp_head_workspace_bytes(int64_t n_groups);
int cp_head_forward_backward(const float *emb, int64_t B, int W, const float *table_w,
                             const float *table_b, float *loss, float *d_emb, float *d_table_w,
                             float *d_table_b, int32_t *pred, int32_t *n_correct, float *logits,
                             void *workspace, size_t workspace_bytes, void *stream);

/* Same loss / gradient / argmax from MATERIALISED logits (G,41,41) (a caller that kept only the
 * logits, e.g. results.py:40): d_logits (G,41,41) may be NULL. */
int cp_logits_loss(const float *logits, int64_t G, float *loss, float *d_logits, int32_t *pred,
                   int32_t *n_correct, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- K3': batch x batch (CLIP) head
 * BASELINE.json config 5.  Generalises the contrastive branch of Model.forward (models.py:112-130)
 * from the per-group 41 x 41 bmm to one B x B similarity matrix with the CLIP loss the reference is
 * modelled after (models.py:65): S = scale * Ehat Ghat^T, scale = exp(logit_scale) (models.py:81,129),
 *   loss = 1/(2B) sum_i [LSE_j S_ij - S_ii] + 1/(2B) sum_j [LSE_i S_ij - S_jj].
 * The B x B matrix is never materialised.  cp_clip_sums / cp_clip_grad run both contractions on the warp-level
 * tensor cores (mma.sync m16n8k16, fp16 operands with a 3-product hi/lo split, fp32 accumulate) when the loop
 * operand has more than 128 rows, and as fp32 FMAs below that.  The entry points are the pieces between which a
 * multi-GPU caller places its collectives (all-gather of Ghat, all-reduce of the column sums,
 * reduce-scatter of d Ghat); on one GPU they are simply called in sequence:
 *   cp_clip_normalize   xhat = x/||x|| (no epsilon, models.py:123,125), inv_norm = 1/||x||
 *   cp_clip_transpose   (n,16) -> (16,ld) k-major copy of the "loop" operand (ld % 4 == 0, ld >= n)
 *   cp_clip_sums        own_sum[i] = sum_j exp(scale*(own_i . loop_j - 1)); own_argmax[i] = first-max j
 *                       row pass: own = Ehat (local rows), loop = Ghat (all); column pass: swapped
 *   cp_clip_loss        sum over the n local samples of [log rowsum + log colsum + 2 scale
 *                       - 2 scale ehat.ghat] / (2B); n_correct = #(row_argmax[i] == row0 + i)
 *   cp_clip_grad        d_own[i,:] = coef * sum_j exp(scale*(own_i.loop_j - 1)) *
 *                                    (1/own_sum[i] + 1/loop_sum[j]) * loop_j     (coef = scale/(2B))
 *   cp_clip_embed_backward  dx = (I - xhat xhat^T)(d_hat - diag_coef * other_hat) * inv_norm
 *                       (diag_coef = scale/B: the -S_ii terms; then the normalisation backward) */
int cp_clip_normalize(const float *x, int64_t n, float *xhat, float *inv_norm, void *stream);
int cp_clip_transpose(const float *xhat, int64_t n, int64_t ld, float *xhat_t, void *stream);
int cp_clip_sums(const float *own, int64_t n_own, const float *loop_t, int64_t n_loop, int64_t ld_loop,
                 float scale, float *own_sum, int32_t *own_argmax, void *stream);
int cp_clip_loss(const float *ehat, const float *ghat, const float *rowsum, const float *colsum,
                 int64_t n, int64_t B, float scale, const int32_t *row_argmax, int64_t row0,
                 float *loss, int32_t *n_correct, void *stream);
int cp_clip_grad(const float *own, int64_t n_own, const float *loop_t, int64_t n_loop, int64_t ld_loop,
                 float scale, const float *own_sum, const float *loop_sum, float coef, float *d_own,
                 void *stream);
int cp_clip_embed_backward(const float *d_hat, const float *xhat, const float *other_hat,
                           const float *inv_norm, int64_t n, float diag_coef, float *dx, void *stream);

/* ---------------------------------------------------------------- K2': glove-angle tower (config 5)
 * The tower the reference keeps commented out (models.py:384-429):
 *   glove (n,glove_dim) -> Linear(glove_dim->256, no bias) -> BN -> ReLU
 *                       -> 3 x [Linear(256->256) -> ReLU -> BN -> Dropout] -> Linear(256->16, no bias)
 * Batch statistics (AdaBN, models.py:17-25) only.  Same conventions as the EMG encoder entry points. */
#define CP_GLOVE_HIDDEN 256
#define CP_GLOVE_BLOCKS 3
typedef struct cp_glove_tensors {
    float *w0;                          /* (256, glove_dim)  glove_net.linear.1.weight */
    float *bn0_w, *bn0_b;               /* (256)             glove_net.linear.2.bn.*   */
    float *w[CP_GLOVE_BLOCKS];          /* (256,256)         glove_net.linear.{4,8,12}.weight */
    float *b[CP_GLOVE_BLOCKS];          /* (256)                                        .bias */
    float *bn_w[CP_GLOVE_BLOCKS];       /* (256)             glove_net.linear.{6,10,14}.bn.*  */
    float *bn_b[CP_GLOVE_BLOCKS];
    float *proj_w;                      /* (16,256)          glove_net.last.0.weight   */
} cp_glove_tensors;

typedef struct cp_glove_opts {
    int32_t glove_dim;                  /* 20 (constants.py:96) or 22 (all sensors); <= 64 */
    int32_t save_for_backward;
    float bn_eps;                       /* 1e-5 */
    float dropout_p;                    /* after each of the 3 blocks; 0 = off */
    uint64_t dropout_seed;
    const uint8_t *ext_masks;           /* optional 3 x (n,256) {0,1} keep masks (parity tests) */
    const uint64_t *dropout_step;       /* optional device counter mixed into the Philox key (CUDA graphs) */
} cp_glove_opts;

size_t cp_glove_workspace_bytes(int64_t n, const cp_glove_opts *opts);
int cp_glove_forward(const cp_glove_tensors *params, const float *glove, int64_t n, float *emb,
                     void *workspace, size_t workspace_bytes, const cp_glove_opts *opts, void *stream);
/* every grad tensor is OVERWRITTEN; must follow a cp_glove_forward(save_for_backward=1) on the workspace */
int cp_glove_backward(const cp_glove_tensors *params, const float *d_emb, int64_t n,
                      const cp_glove_tensors *grads, void *workspace, size_t workspace_bytes,
                      const cp_glove_opts *opts, void *stream);

/* ---------------------------------------------------------------- --prediction mode: classifier head
 * EMGNet.last with prediction=True (models.py:300-309) on the trunk output (cp_encoder_forward, trunk_only), followed
 * by Model.forward's row normalisation (models.py:118) and prediction_loss (models.py:175-196), in one call:
 *   a7 (n,512) -> Linear(512->128) -> ReLU -> BN(128) -> Linear(128->41, no bias) -> z;  features = z / ||z||
 *   loss = mean CE(features, labels);  pred = first-max argmax;  n_correct = #(pred == label)
 * Outputs (each may be NULL): features (n,41), loss, pred (n) int32, n_correct (1) int32.  d_a7 (n,512) and grads are
 * both NULL (forward only) or both given (every grad tensor is OVERWRITTEN; bn_rm / bn_rv of grads ignored).
 * bn_mode as in cp_encoder_opts (stock BN: running statistics of nn.BatchNorm1d(128)).
 * Parity tap: on return the workspace begins with relu(Linear1(a7)), (n,128) fp32. */
#define CP_CLS_HIDDEN 128
typedef struct cp_cls_tensors {
    float *w1, *b1;            /* (128,512), (128)   emg_net.last.0.{weight,bias}      */
    float *bn_w, *bn_b;        /* (128)              emg_net.last.2[.bn].{weight,bias} */
    float *bn_rm, *bn_rv;      /* (128)              running statistics (stock BN) or NULL */
    float *w2;                 /* (41,128)           emg_net.last.3.weight             */
} cp_cls_tensors;
size_t cp_cls_workspace_bytes(int64_t n);
int cp_cls_forward_backward(const cp_cls_tensors *params, const float *a7, const int64_t *labels, int64_t n,
                            int bn_mode, float bn_momentum, float bn_eps, float *features, float *loss,
                            int32_t *pred, int32_t *n_correct, float *d_a7, const cp_cls_tensors *grads,
                            void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- K4: windowed majority vote
 * Replaces the vote loop of contrastive_loopy_loss (models.py:149-163, constants.py:74-78).
 * pred: (B,W,41) int32, W <= 256.  votes: (B,n_votes) int32 = #rows whose prefix-mode over the first
 * min(v+1,W) samples equals the row label (mode ties -> smallest label).  y_pred: (B,41) int64 =
 * full-window mode.  Entries of pred outside [0,41) cast no vote. */
int cp_vote_eval(const int32_t *pred, int64_t B, int W, int n_votes, int32_t *votes,
                 int64_t *y_pred, void *stream);

/* ---------------------------------------------------------------- K4': class-subset evaluator
 * Implements the README-only test-time evaluator (README.md:11,15; inputs = the logits dumped by
 * results.py:42-61).  Two steps:
 *  cp_rank_rows:   order[r, k] = label with the k-th largest logit of row r (ties -> smaller
 *                  label first), uint8, for r over (B*W*41) rows of 41 logits.
 *  cp_subset_eval: for trial t with class mask masks[t, 0..40], every group b and every row
 *                  i in the subset: pred_w = first label in order[(b,w,i), :] that is in the
 *                  subset; decision = mode over the W window (ties -> smallest label);
 *                  correct[t] += decision == i; total[t] = B * |subset|.  int64 outputs are
 *                  OVERWRITTEN. */
int cp_rank_rows(const float *logits, int64_t n_rows, uint8_t *order, void *stream);
int cp_subset_eval(const uint8_t *order, int64_t B, int W, const uint8_t *masks, int64_t n_trials,
                   int64_t *correct, int64_t *total, void *stream);

/* ---------------------------------------------------------------- results.py:58 confusion matrix
 * counts[t*C + p] = number of decisions with y_true == t and y_pred == p (sklearn.metrics.confusion_matrix
 * on labels 0..C-1, C <= 64).  counts (C*C int64) is OVERWRITTEN; labels outside [0, C) set *err_flag (may be
 * NULL) and are skipped. */
int cp_confusion_matrix(const int64_t *y_true, const int64_t *y_pred, int64_t n, int n_classes,
                        int64_t *counts, int *err_flag, void *stream);

/* ---------------------------------------------------------------- K5: l2 regulariser
 * Model.l2 / EMGNet.l2 / GLOVENet.l2 (models.py:225-228, 344-349, 467-472): sum over a parameter list of the
 * un-squared Frobenius norms.  `tensors` / `sizes` / `grads` are HOST arrays of n_tensors (<= 32) device pointers /
 * element counts.  forward: norms[t] = ||W_t||_2, *total = sum_t norms[t] (both device, OVERWRITTEN).
 * backward: grads[t] = coef * (*g_total) * W_t / norms[t] (0 where norms[t] == 0), OVERWRITTEN. */
size_t cp_l2_workspace_bytes(int n_tensors);
int cp_l2_forward(const float *const *tensors, const int64_t *sizes, int n_tensors, float *norms, float *total,
                  void *workspace, size_t workspace_bytes, void *stream);
int cp_l2_backward(const float *const *tensors, const int64_t *sizes, int n_tensors, const float *norms,
                   const float *g_total, float coef, float *const *grads, void *stream);

/* ---------------------------------------------------------------- K6: step prologue + optimiser (train.py:72-73, 95-108)
 * The two ends of the reference's loop body that it leaves to autograd and torch.optim: `+ model.l2()` /
 * `loss.backward()`'s regulariser gradient, and `optimizer_emg.step(); optimizer_glove.step()`.
 * cp_step_prologue: ONE launch.  norms[t] = ||W_t||_2 of the n_tensors (<= CP_STEP_MAX_TENSORS) regularised tensors
 *   (HOST arrays of device pointers / element counts; same arithmetic as cp_l2_forward) and counters[0..n_counters)
 *   (device int64: the dropout step of cp_encoder_opts.dropout_step, Adam's t) += 1.  The first 256 bytes of
 *   `workspace` must be ZERO before the first call; the kernel leaves them zero (graph replays included).
 * cp_adam_step: ONE launch over every parameter tensor of both optimisers.  params / sizes / offsets / lr_index / reg /
 *   norm_index are HOST arrays of n_tensors entries; grads, exp_avg, exp_avg_sq are FLAT device buffers with tensor t
 *   at element offset offsets[t]; lr is a DEVICE array of doubles (lr[lr_index[t]]: a scheduler rewrites it between
 *   graph replays); *step (device) is Adam's t >= 1.  g = grads + (norm_index[t] >= 0 ? reg[t] * W / norms[norm_index[t]]
 *   : 0) (the gradient of reg * ||W||_2, 0 where the norm is 0: models.py:225-228, 344-349), then torch.optim.Adam's
 *   default update (no weight decay / amsgrad / maximize): m = b1 m + (1-b1) g, v = b2 v + (1-b2) g^2,
 *   W -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).  params, exp_avg, exp_avg_sq are updated IN PLACE. */
#define CP_STEP_MAX_TENSORS 48
size_t cp_step_workspace_bytes(int n_tensors);
int cp_step_prologue(const float *const *tensors, const int64_t *sizes, int n_tensors, float *norms,
                     int64_t *counters, int n_counters, void *workspace, size_t workspace_bytes, void *stream);
int cp_adam_step(float *const *params, const int64_t *sizes, const int64_t *offsets, int n_tensors,
                 const float *grads, float *exp_avg, float *exp_avg_sq, const double *lr, const int32_t *lr_index,
                 const float *reg, const int32_t *norm_index, const float *norms, const int64_t *step,
                 double beta1, double beta2, double eps, void *stream);

/* ---------------------------------------------------------------- offline preprocessing
 * load.py:85-101 / utils.py:134-156: per (subject, stimulus, repetition) segment raw[seg_len, n_ch] (float32, time
 * major): x = raw * gain -> IIR filter (b, a: HOST arrays of n_coef <= 17 doubles, a[0] == 1; scipy lfilter
 * semantics, result rounded to float32) -> moving RMS (odd rms_window <= 33, uniform_filter1d mode 'nearest',
 * trimmed by rms_window/2 on both sides) -> out[seg, j, c] = rms[time_idx[j]] for the n_out DEVICE indices
 * time_idx[j] in [0, n_rms).  Bit-exact with scipy's double-precision loops.  scratch: n_seg*n_ch*n_rms floats. */
size_t cp_emg_preprocess_scratch_elems(int64_t n_seg, int n_ch, int n_rms);
int cp_emg_preprocess(const float *raw, int64_t n_seg, int seg_len, int n_ch, const double *b, const double *a,
                      int n_coef, float gain, int rms_window, int n_rms, const int32_t *time_idx, int n_out,
                      float *out, float *scratch, size_t scratch_elems, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CPROS_H */
