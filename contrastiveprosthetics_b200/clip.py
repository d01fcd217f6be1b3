"""Batch x batch (CLIP) contrastive head -- BASELINE.json config 5 (a8/a9
generalised, 8e "one exchange step").

    loss = clip_head(E, G, logit_scale)          E, G: this rank's (n,16) un-normalised embeddings

    S = exp(logit_scale) * Ehat Ghat^T over the GLOBAL batch B = n * world_size     (models.py:112-130)
    loss = 1/2 [ mean_i CE(S[i,:], i) + mean_j CE(S[:,j], j) ]                      (models.py:65, CLIP)

All arithmetic runs in libcpros.so (cp_clip_*; csrc/clip.cu): the B x B matrix is never materialised.
With world_size > 1 (one process per GPU, torch.distributed / NCCL over NVLink) every rank evaluates
its row strip E_local x G_all:
    all-gather      Ghat                       (B,16)   4 MB at B = 65,536
    all-reduce      column sums                (B,)     256 KB
    reduce-scatter  d Ghat partials            (B,16) -> (n,16)
    all-reduce      loss, correct count        scalars
The returned gradients are those of the GLOBAL loss w.r.t. the rank's own embeddings, so parameter
gradients must be SUMMED over ranks (dist.FlatGradAllReduce(average=False)).
"""
import math

import torch
import torch.distributed as dist

from . import _lib
from . import dist as cpdist

D_E = 16
MAX_SCALE = 43.0          # largest exp(logit_scale) for which no row / column sum can underflow (see clip_head)


class _CudaOps:
    """The six libcpros entry points of the head, on tensors."""

    @staticmethod
    def normalize(x):
        xhat = torch.empty_like(x)
        inv = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().cp_clip_normalize(_lib.ptr(x, torch.float32), x.shape[0], _lib.ptr(xhat), _lib.ptr(inv),
                                                _lib.stream()), "cp_clip_normalize")
        return xhat, inv

    @staticmethod
    def transpose(xhat):
        n = xhat.shape[0]
        ld = (n + 3) // 4 * 4
        xt = torch.empty((D_E, ld), dtype=torch.float32, device=xhat.device)
        _lib.check(_lib.lib().cp_clip_transpose(_lib.ptr(xhat, torch.float32), n, ld, _lib.ptr(xt), _lib.stream()),
                   "cp_clip_transpose")
        return xt

    @staticmethod
    def sums(own, loop_t, n_loop, scale, want_argmax):
        n_own = own.shape[0]
        s = torch.empty(n_own, dtype=torch.float32, device=own.device)
        arg = torch.empty(n_own, dtype=torch.int32, device=own.device) if want_argmax else None
        _lib.check(_lib.lib().cp_clip_sums(_lib.ptr(own, torch.float32), n_own, _lib.ptr(loop_t), n_loop,
                                           loop_t.shape[1], scale, _lib.ptr(s), _lib.ptr(arg), _lib.stream()),
                   "cp_clip_sums")
        return s, arg

    @staticmethod
    def loss(ehat, ghat, rowsum, colsum, B, scale, row_arg, row0):
        n = ehat.shape[0]
        loss = torch.empty((), dtype=torch.float32, device=ehat.device)
        ncor = torch.empty((), dtype=torch.int32, device=ehat.device)
        _lib.check(_lib.lib().cp_clip_loss(_lib.ptr(ehat), _lib.ptr(ghat), _lib.ptr(rowsum), _lib.ptr(colsum), n, B,
                                           scale, _lib.ptr(row_arg), row0, _lib.ptr(loss), _lib.ptr(ncor),
                                           _lib.stream()), "cp_clip_loss")
        return loss, ncor

    @staticmethod
    def grad(own, loop_t, n_loop, scale, own_sum, loop_sum, coef):
        d = torch.empty_like(own)
        _lib.check(_lib.lib().cp_clip_grad(_lib.ptr(own, torch.float32), own.shape[0], _lib.ptr(loop_t), n_loop,
                                           loop_t.shape[1], scale, _lib.ptr(own_sum), _lib.ptr(loop_sum), coef,
                                           _lib.ptr(d), _lib.stream()), "cp_clip_grad")
        return d

    @staticmethod
    def embed_backward(d_hat, xhat, other_hat, inv_norm, diag_coef):
        dx = torch.empty_like(xhat)
        _lib.check(_lib.lib().cp_clip_embed_backward(_lib.ptr(d_hat), _lib.ptr(xhat), _lib.ptr(other_hat),
                                                     _lib.ptr(inv_norm), xhat.shape[0], diag_coef, _lib.ptr(dx),
                                                     _lib.stream()), "cp_clip_embed_backward")
        return dx


def _world(group):
    if group == "local":                       # force the single-process path inside a distributed job
        return 1, 0
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


class _ClipHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, E, G, scale, want_grad, group):
        ops = _CudaOps
        world, rank = _world(group)
        E, G = E.contiguous(), G.contiguous()
        n = E.shape[0]
        if G.shape != E.shape or E.shape[1] != D_E:
            raise RuntimeError(f"clip_head expects two (n,{D_E}) tensors, got {tuple(E.shape)} and {tuple(G.shape)}")
        B, row0 = n * world, rank * n
        ehat, inv_e = ops.normalize(E)
        ghat, inv_g = ops.normalize(G)
        if world > 1:
            ghat_all = cpdist.all_gather_rows(ghat, group)
        else:
            ghat_all = ghat
        ght = ops.transpose(ghat_all)                       # (16, B)  loop operand of the row pass
        eht = ops.transpose(ehat)                           # (16, n)  loop operand of the column pass
        rowsum, row_arg = ops.sums(ehat, ght, B, scale, True)
        colsum, _ = ops.sums(ghat_all, eht, n, scale, False)
        if world > 1:
            dist.all_reduce(colsum, group=group)
        col_local = colsum[row0:row0 + n].contiguous()
        loss, ncor = ops.loss(ehat, ghat, rowsum, col_local, B, scale, row_arg, row0)
        if world > 1:
            dist.all_reduce(loss, group=group)
            dist.all_reduce(ncor, group=group)
        if want_grad:
            coef = scale / (2.0 * B)
            d_ehat = ops.grad(ehat, ght, B, scale, rowsum, colsum, coef)
            d_ghat = ops.grad(ghat_all, eht, n, scale, colsum, rowsum, coef)
            if world > 1:
                d_ghat = cpdist.reduce_scatter_rows(d_ghat, group)
            ctx.saved = (ops.embed_backward(d_ehat, ehat, ghat, inv_e, scale / B),
                         ops.embed_backward(d_ghat, ghat, ehat, inv_g, scale / B))
        else:
            ctx.saved = None
        ctx.mark_non_differentiable(ncor, row_arg)
        return loss, ncor, row_arg

    @staticmethod
    def backward(ctx, g_loss, _gn, _ga):
        if ctx.saved is None:
            raise RuntimeError("clip_head was run without gradients")
        dE, dG = ctx.saved
        return dE * g_loss, dG * g_loss, None, None, None


def clip_head(E, G, logit_scale=0.0, group=None):
    """Returns (loss, n_correct, row_argmax): the global symmetric CLIP loss (differentiable w.r.t. E and
    G), the global number of rows whose arg-max column is their own sample, and this rank's arg-max
    columns (global indices, first maximum)."""
    scale = float(math.exp(float(logit_scale)))
    if not scale <= MAX_SCALE:
        # the sweeps shift every exponent by the largest possible logit (`scale`, cos = 1) and use ex2.approx.ftz:
        # a row whose best cosine is c keeps a non-zero sum only while scale * (1 - c) * log2(e) < 126, which holds
        # for ANY data when scale <= 43 (c >= -1).  Beyond that the sums could flush to 0 (inf loss): refuse.
        raise ValueError(f"clip_head: exp(logit_scale) = {scale:.3g} exceeds the supported maximum {MAX_SCALE}")
    want_grad = torch.is_grad_enabled() and (E.requires_grad or G.requires_grad)
    return _ClipHeadFn.apply(E, G, scale, want_grad, group)


# =================================================================================== towers + model
import ctypes  # noqa: E402

import numpy as np  # noqa: E402
import torch.nn as nn  # noqa: E402

from .constants import GLOVE_DIM  # noqa: E402

GLOVE_HIDDEN, GLOVE_BLOCKS = 256, 3


class GloveTower(nn.Module):
    """The glove-angle tower the reference keeps commented out (models.py:384-429), with the module
    indices (and therefore state-dict keys: linear.1, linear.2.bn, linear.{4,8,12}, linear.{6,10,14}.bn,
    last.0) the block would have had live:
        Flatten, Linear(glove_dim->256, no bias), BN, ReLU, 3 x [Linear(256->256), ReLU, BN, Dropout];
        last = Linear(256->d_e, no bias).
    Parameter container only: forward / backward run in libcpros (cp_glove_forward / cp_glove_backward)."""

    def __init__(self, glove_dim=GLOVE_DIM, d_e=D_E, dp=.5, device="cuda"):
        super().__init__()
        from .models import AdaBatchNorm1d
        if d_e != D_E:
            raise NotImplementedError("libcpros is built for d_e = 16")
        self.device = torch.device(device)
        self.glove_dim, self.d_e, self.dp = glove_dim, d_e, dp
        blocks = [nn.Flatten(), nn.Linear(glove_dim, GLOVE_HIDDEN, bias=False),
                  AdaBatchNorm1d(GLOVE_HIDDEN, device=device), nn.ReLU()]
        for _ in range(GLOVE_BLOCKS):
            blocks += [nn.Linear(GLOVE_HIDDEN, GLOVE_HIDDEN), nn.ReLU(), AdaBatchNorm1d(GLOVE_HIDDEN, device=device),
                       nn.Dropout(dp)]
        self.linear = nn.Sequential(*blocks)
        self.last = nn.Sequential(nn.Linear(GLOVE_HIDDEN, d_e, bias=False))
        self.to(self.device)
        self.dropout_seed = 0x61073
        self._step = 0
        self.dropout_step = None               # device int64 counter mixed into the Philox key (CUDA graphs)
        self.ext_dropout_masks = None          # (3, n, 256) uint8 keep masks injected by parity tests

    def kernel_params(self):
        lin = [m for m in self.linear if isinstance(m, nn.Linear)]
        bns = [m.bn for m in self.linear if hasattr(m, "bn")]
        return ([lin[0].weight, bns[0].weight, bns[0].bias] + [m.weight for m in lin[1:]] + [m.bias for m in lin[1:]] +
                [m.weight for m in bns[1:]] + [m.bias for m in bns[1:]] + [self.last[0].weight])

    def forward(self, GLOVE):
        """(..., glove_dim) -> (n, d_e)."""
        x = GLOVE.reshape(-1, self.glove_dim)
        dp = float(self.dp) if self.training else 0.0
        self._step += 1
        cfg = {"glove_dim": self.glove_dim, "dropout_p": dp,
               "seed": (self.dropout_seed * 1000003 + self._step) & 0xFFFFFFFFFFFFFFFF,
               "ext_masks": self.ext_dropout_masks if dp > 0 else None, "dropout_step": self.dropout_step,
               "need_bwd": torch.is_grad_enabled() and self.training}
        return _GloveFn.apply(x, cfg, *self.kernel_params())

    def l2(self):
        from .models import _l2_of
        return _l2_of(self)


def _fill_glove(struct, t):
    P = _lib.ptr
    struct.w0, struct.bn0_w, struct.bn0_b = P(t[0]), P(t[1]), P(t[2])
    for b in range(GLOVE_BLOCKS):
        struct.w[b], struct.b[b] = P(t[3 + b]), P(t[6 + b])
        struct.bn_w[b], struct.bn_b[b] = P(t[9 + b]), P(t[12 + b])
    struct.proj_w = P(t[15])
    return struct


class _GloveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cfg, *params):
        L = _lib.lib()
        if x.dtype != torch.float32:
            raise RuntimeError("glove input must be float32")
        x = x.contiguous()
        n = x.shape[0]
        opts = _lib.GloveOpts(glove_dim=cfg["glove_dim"], save_for_backward=int(cfg["need_bwd"]), bn_eps=1e-5,
                              dropout_p=float(cfg["dropout_p"]), dropout_seed=int(cfg["seed"]),
                              ext_masks=_lib.ptr(cfg["ext_masks"], torch.uint8),
                              dropout_step=_lib.ptr(cfg.get("dropout_step"), torch.int64))
        nbytes = L.cp_glove_workspace_bytes(n, ctypes.byref(opts))
        if nbytes == 0:
            raise RuntimeError("cp_glove_workspace_bytes rejected the configuration")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        emb = torch.empty((n, D_E), dtype=torch.float32, device=x.device)
        tens = _fill_glove(_lib.GloveTensors(), params)
        _lib.check(L.cp_glove_forward(ctypes.byref(tens), _lib.ptr(x), n, _lib.ptr(emb), _lib.ptr(ws), nbytes,
                                      ctypes.byref(opts), _lib.stream()), "cp_glove_forward")
        if cfg["need_bwd"]:
            ctx.ws, ctx.opts, ctx.tens, ctx.n, ctx.params = ws, opts, tens, n, params
        return emb

    @staticmethod
    def backward(ctx, d_emb):
        L = _lib.lib()
        grads = [torch.empty_like(p) for p in ctx.params]
        gt = _fill_glove(_lib.GloveTensors(), grads)
        _lib.check(L.cp_glove_backward(ctypes.byref(ctx.tens), _lib.ptr(d_emb.contiguous()), ctx.n, ctypes.byref(gt),
                                       _lib.ptr(ctx.ws), ctx.ws.numel(), ctypes.byref(ctx.opts), _lib.stream()),
                   "cp_glove_backward")
        ctx.ws = None
        return (None, None) + tuple(grads)


class ClipModel(nn.Module):
    """Config 5: EMG tower (models.EMGNet, one window per sample) + glove-angle tower, trained with the
    batch x batch CLIP loss.  `forward(EMG, GLOVE)` -> (emg_emb, glove_emb); `loss(...)` -> global loss."""

    def __init__(self, params, glove_dim=GLOVE_DIM, device="cuda"):
        super().__init__()
        from .models import EMGNet
        self.params = params
        self.device = torch.device(device)
        self.emg_net = EMGNet(d_e=params['d_e'], dp=params['dp_emg'], adabn=True, device=device)
        self.glove_net = GloveTower(glove_dim=glove_dim, d_e=params['d_e'], dp=params['dp_glove'], device=device)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1) / 0.07, requires_grad=False)   # models.py:81
        self.to(self.device)
        self.n_correct = []

    def forward(self, EMG, GLOVE):
        return self.emg_net.encode_flat(EMG), self.glove_net(GLOVE)

    def loss(self, emg_emb, glove_emb, group=None):
        loss, ncor, _ = clip_head(emg_emb, glove_emb, float(self.logit_scale), group)
        self.n_correct.append(ncor)
        return loss

    def l2(self):
        return self.glove_net.l2() * self.params['reg_glove'] + self.emg_net.l2() * self.params['reg_emg']
