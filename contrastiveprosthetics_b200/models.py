"""Drop-in `Model` / `EMGNet` / `GLOVENet` / `AdaBatchNorm{1,2}d` (reference: code/models.py).

Same constructor signatures, attributes, method names, argument meaning and state-dict keys as
the reference, so train.py-style callers and checkpoints keep working -- but no
layer is ever *called*: the nn.Modules below are parameter containers, and every forward /
backward goes through libcpros.so (hand-written sm_100a CUDA, see include/cpros.h):

    EMGNet.forward            -> cp_encoder_forward / cp_encoder_backward      (models.py:319-342)
    Model.forward + .loss     -> cp_head_forward_backward (one fused launch)    (models.py:112-208)
    vote loop in .loss (eval) -> cp_vote_eval                                   (models.py:149-163)

    Model.forward + .loss, --prediction -> cp_encoder_forward(trunk_only) + cp_cls_forward_backward
                                  (classifier head + normalised-logit CE)           (models.py:113-119, 175-196, 300-309)

There is no PyTorch/CPU fallback: CPU tensors raise.  `--glove` (with `--prediction`) is broken in the reference
(models.py:417 vs 451: a 20-dim input into Linear(256,128)) and raises NotImplementedError; the vote evaluation of
`--prediction` fails the reference's own shape assertion (models.py:178) and raises the same AssertionError here.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .constants import (EMG_DIM, GLOVE_DIM, MAX_TASKS_TRAIN, PREDICTION_WINDOW,
                        PREDICTION_WINDOW_SIZE, VOTE)

N_VOTES = PREDICTION_WINDOW - 1      # `for win in range(1, PREDICTION_WINDOW)`  (models.py:153)


class AdaBatchNorm1d(nn.Module):
    """models.py:17-25 -- BatchNorm with batch statistics in train AND eval (momentum 0, no running
    stats).  Parameter container only; statistics are computed by the encoder kernels."""

    def __init__(self, num_features, device="cuda"):
        super().__init__()
        self.device = torch.device(device)
        self.bn = nn.BatchNorm1d(num_features=num_features, momentum=0, track_running_stats=False)
        self.to(self.device)

    def forward(self, X):
        raise RuntimeError("AdaBatchNorm1d is fused into cp_encoder_forward; call EMGNet.forward")


class AdaBatchNorm2d(nn.Module):
    """models.py:27-35."""

    def __init__(self, num_features, device="cuda"):
        super().__init__()
        self.device = torch.device(device)
        self.bn = nn.BatchNorm2d(num_features=num_features, momentum=0, track_running_stats=False)
        self.to(self.device)

    def forward(self, X):
        raise RuntimeError("AdaBatchNorm2d is fused into cp_encoder_forward; call EMGNet.forward")


def _bn_leaf(m):
    return m.bn if isinstance(m, (AdaBatchNorm1d, AdaBatchNorm2d)) else m


# ------------------------------------------------------------------------------ autograd glue
def _fill_tensors(struct, conv1_w, conv1_b, conv2_w, conv2_b, fc_w, fc_b, proj_w, bn_w, bn_b,
                  bn_rm=None, bn_rv=None):
    P = _lib.ptr
    struct.conv1_w, struct.conv1_b = P(conv1_w), P(conv1_b)
    struct.conv2_w, struct.conv2_b = P(conv2_w), P(conv2_b)
    for i in range(_lib.N_FC):
        struct.fc_w[i], struct.fc_b[i] = P(fc_w[i]), P(fc_b[i])
    struct.proj_w = P(proj_w)
    for i in range(_lib.N_BN):
        struct.bn_w[i], struct.bn_b[i] = P(bn_w[i]), P(bn_b[i])
        struct.bn_rm[i] = P(bn_rm[i]) if bn_rm is not None else None
        struct.bn_rv[i] = P(bn_rv[i]) if bn_rv is not None else None
    return struct


def _split(params):
    """flat list of 37 tensors -> named groups (order = EMGNet.kernel_params())."""
    conv1_w, conv1_b, conv2_w, conv2_b = params[0:4]
    fc_w, fc_b = params[4:11], params[11:18]
    proj_w = params[18]                   # None in trunk-only (--prediction) mode
    bn_w, bn_b = params[19:28], params[28:37]
    return conv1_w, conv1_b, conv2_w, conv2_b, fc_w, fc_b, proj_w, bn_w, bn_b


def _make_sync_hook(ws, group):
    """SyncBN hook (cp_allreduce_fn): the library hands back a device pointer inside the workspace tensor
    `ws`; sum that slice over the ranks with torch.distributed (NCCL) on the current stream."""
    import torch.distributed as dist
    base = ws.data_ptr()

    def hook(_user, buf, count, _stream):
        try:
            off = buf - base
            dist.all_reduce(ws[off:off + 8 * count].view(torch.float64), group=group)
            return 0
        except Exception as e:          # never let an exception cross the C boundary
            print(f"contrastiveprosthetics_b200: SyncBN all-reduce failed: {e!r}")
            return 1
    return _lib.ALLREDUCE_FN(hook)


class _EncoderFn(torch.autograd.Function):
    """x (N,12) -> emb (N,16) through cp_encoder_forward; backward through cp_encoder_backward."""

    @staticmethod
    def forward(ctx, x, cfg, *params):
        L = _lib.lib()
        if x.dtype != torch.float32:
            raise RuntimeError("encoder input must be float32")
        x = x.contiguous()
        n = x.shape[0]
        need_bwd = cfg["need_bwd"]
        trunk = bool(cfg.get("trunk_only"))
        opts = _lib.EncoderOpts(bn_mode=cfg["bn_mode"], engine=cfg["engine"], bn_momentum=0.1, bn_eps=1e-5,
                                dropout_p=float(cfg["dropout_p"]), save_for_backward=int(need_bwd),
                                dropout_seed=int(cfg["seed"]),
                                ext_masks=_lib.ptr(cfg["ext_masks"], torch.uint8),
                                dropout_step=_lib.ptr(cfg.get("dropout_step"), torch.int64),
                                trunk_only=int(trunk))
        nbytes = L.cp_encoder_workspace_bytes(n, ctypes.byref(opts))
        if nbytes == 0:
            raise RuntimeError("cp_encoder_workspace_bytes rejected the configuration")
        alloc = cfg.get("ws_alloc")          # tests: a caller-owned (guarded) workspace; default: a fresh tensor
        ws = alloc(nbytes) if alloc is not None else torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        emb = torch.empty((n, 512 if trunk else 16), dtype=torch.float32, device=x.device)
        hook = None
        if cfg.get("sync_bn"):
            hook = _make_sync_hook(ws, cfg.get("group"))
            opts.allreduce = ctypes.cast(hook, ctypes.c_void_p)
        tens = _fill_tensors(_lib.EncoderTensors(), *_split(params), bn_rm=cfg["bn_rm"], bn_rv=cfg["bn_rv"])
        _lib.check(L.cp_encoder_forward(ctypes.byref(tens), _lib.ptr(x), n, _lib.ptr(emb), _lib.ptr(ws),
                                        nbytes, ctypes.byref(opts), _lib.stream()), "cp_encoder_forward")
        if need_bwd:
            ctx.ws, ctx.opts, ctx.tens, ctx.n = ws, opts, tens, n
            ctx.hook = hook                # keeps the ctypes callback alive until backward has used it
            ctx.params = params            # keeps the storages (and pointers in `tens`) alive
            ctx.cfg = cfg
            if cfg.get("tap") is not None:
                cfg["tap"].update(ws=ws, opts=opts, n=n)
        return emb

    @staticmethod
    def backward(ctx, d_emb):
        grads = [torch.empty_like(p) if p is not None else None for p in ctx.params]
        _encoder_backward_into(ctx, d_emb, grads)
        return (None, None) + tuple(grads)


def _encoder_backward_into(ctx, d_emb, grads):
    """cp_encoder_backward of the forward recorded in `ctx`; the 37 parameter gradients (order of
    EMGNet.kernel_params()) are OVERWRITTEN in the caller's tensors `grads` (autograd: fresh ones; step.LeanTrainStep:
    views of its flat gradient bucket)."""
    d_emb = d_emb.contiguous()
    gt = _fill_tensors(_lib.EncoderTensors(), *_split(grads))
    _lib.check(_lib.lib().cp_encoder_backward(ctypes.byref(ctx.tens), _lib.ptr(d_emb), ctx.n, ctypes.byref(gt),
                                              _lib.ptr(ctx.ws), ctx.ws.numel(), ctypes.byref(ctx.opts),
                                              _lib.stream()), "cp_encoder_backward")
    ctx.ws = None


class PlainCtx:
    """Stands in for autograd's ctx when a fused op is driven without autograd (step.LeanTrainStep)."""

    def mark_non_differentiable(self, *_):
        pass


def _head_launch(emb, table_w, table_b, B, W, want_grad, want_logits, d_w=None, d_b=None):
    """cp_head_forward_backward.  Returns (loss, pred, ncor, logits | None, d_emb, d_w, d_b); d_w / d_b may be the
    caller's tensors (step.LeanTrainStep: views of its gradient bucket), otherwise fresh ones."""
    L = _lib.lib()
    dev = emb.device
    G = B * W
    emb = emb.contiguous()
    loss = torch.empty((), dtype=torch.float32, device=dev)
    pred = torch.empty((G, MAX_TASKS_TRAIN), dtype=torch.int32, device=dev)
    ncor = torch.empty((G,), dtype=torch.int32, device=dev)
    logits = torch.empty((G, MAX_TASKS_TRAIN, MAX_TASKS_TRAIN), dtype=torch.float32, device=dev) \
        if want_logits else None
    d_emb = torch.empty_like(emb) if want_grad else None
    if want_grad and d_w is None:
        d_w, d_b = torch.empty_like(table_w), torch.empty_like(table_b)
    nbytes = L.cp_head_workspace_bytes(G)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    P = _lib.ptr
    _lib.check(L.cp_head_forward_backward(P(emb), B, W, P(table_w.contiguous()), P(table_b.contiguous()),
                                          P(loss), P(d_emb), P(d_w), P(d_b), P(pred), P(ncor), P(logits),
                                          P(ws), nbytes, _lib.stream()), "cp_head_forward_backward")
    return loss, pred, ncor, logits, d_emb, d_w, d_b


class _HeadFn(torch.autograd.Function):
    """Fused head: emb (N,16) [+ class table] -> loss; gradients are produced by the SAME launch
    and only scaled by grad_output in backward."""

    @staticmethod
    def forward(ctx, emb, table_w, table_b, B, W, want_grad, want_logits):
        loss, pred, ncor, logits, d_emb, d_w, d_b = _head_launch(emb, table_w, table_b, B, W, want_grad, want_logits)
        ctx.saved = (d_emb, d_w, d_b)
        ctx.mark_non_differentiable(pred, ncor)
        if logits is None:
            logits = torch.empty(0, device=emb.device)
        ctx.mark_non_differentiable(logits)
        return loss, pred, ncor, logits

    @staticmethod
    def backward(ctx, g_loss, _gp, _gn, _gl):
        d_emb, d_w, d_b = ctx.saved
        if d_emb is None:
            raise RuntimeError("head was run without gradients")
        return d_emb * g_loss, d_w * g_loss, d_b * g_loss, None, None, None, None


class _LogitsLossFn(torch.autograd.Function):
    """Loss / argmax from materialised logits (callers that kept only the logits tensor)."""

    @staticmethod
    def forward(ctx, logits, want_grad):
        L = _lib.lib()
        logits = logits.contiguous()
        G = logits.shape[0]
        dev = logits.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        pred = torch.empty((G, MAX_TASKS_TRAIN), dtype=torch.int32, device=dev)
        ncor = torch.empty((G,), dtype=torch.int32, device=dev)
        d_logits = torch.empty_like(logits) if want_grad else None
        nbytes = L.cp_head_workspace_bytes(G)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        P = _lib.ptr
        _lib.check(L.cp_logits_loss(P(logits), G, P(loss), P(d_logits), P(pred), P(ncor), P(ws), nbytes,
                                    _lib.stream()), "cp_logits_loss")
        ctx.saved = d_logits
        ctx.mark_non_differentiable(pred, ncor)
        return loss, pred, ncor

    @staticmethod
    def backward(ctx, g_loss, _gp, _gn):
        return ctx.saved * g_loss, None


class _ClsHeadFn(torch.autograd.Function):
    """--prediction mode: trunk output a7 (N,512) + labels -> (features (N,41), loss, pred, n_correct) through
    cp_cls_forward_backward; the gradients come out of the SAME call and are scaled by grad_output in backward."""

    @staticmethod
    def forward(ctx, a7, labels, cfg, w1, b1, bn_w, bn_b, w2):
        L = _lib.lib()
        dev = a7.device
        a7 = a7.contiguous()
        n = a7.shape[0]
        labels = labels.reshape(-1).to(torch.int64).contiguous()
        if labels.numel() != n:
            raise RuntimeError(f"prediction head: {n} rows but {labels.numel()} labels")
        feats = torch.empty((n, MAX_TASKS_TRAIN), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        pred = torch.empty((n,), dtype=torch.int32, device=dev)
        ncor = torch.empty((), dtype=torch.int32, device=dev)
        want_grad = cfg["want_grad"]
        P = _lib.ptr
        params = _lib.ClsTensors(w1=P(w1), b1=P(b1), bn_w=P(bn_w), bn_b=P(bn_b), bn_rm=P(cfg["bn_rm"]),
                                 bn_rv=P(cfg["bn_rv"]), w2=P(w2))
        d_a7 = torch.empty_like(a7) if want_grad else None
        grads = [torch.empty_like(t) for t in (w1, b1, bn_w, bn_b, w2)] if want_grad else None
        gstruct = _lib.ClsTensors(w1=P(grads[0]), b1=P(grads[1]), bn_w=P(grads[2]), bn_b=P(grads[3]),
                                  w2=P(grads[4])) if want_grad else None
        nbytes = L.cp_cls_workspace_bytes(n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(L.cp_cls_forward_backward(ctypes.byref(params), P(a7), P(labels), n, cfg["bn_mode"], 0.1, 1e-5,
                                             P(feats), P(loss), P(pred), P(ncor), P(d_a7),
                                             ctypes.byref(gstruct) if want_grad else None, P(ws), nbytes, _lib.stream()),
                   "cp_cls_forward_backward")
        ctx.saved = (d_a7, grads)
        if cfg.get("tap") is not None:          # parity tap: the workspace starts with relu(linear1(a7)), (n,128) fp32
            cfg["tap"]["relu_head"] = ws[:n * 128 * 4].view(torch.float32).reshape(n, 128).clone()
        ctx.mark_non_differentiable(feats, pred, ncor)
        return feats, loss, pred, ncor

    @staticmethod
    def backward(ctx, _gf, g_loss, _gp, _gn):
        d_a7, grads = ctx.saved
        if d_a7 is None:
            raise RuntimeError("prediction head was run without gradients")
        return (d_a7 * g_loss, None, None) + tuple(g * g_loss for g in grads)


class FusedLogits:
    """What `Model.forward` returns in training: a handle to the fused head's results.  The 41x41
    similarity tiles never leave shared memory; `.materialize()` re-runs the head with a logits
    output for callers that really want the tensor."""

    def __init__(self, model, emb, B, W, loss, pred, ncor):
        self._model, self._emb, self.B, self.W = model, emb, B, W
        self.loss, self.pred, self.ncor = loss, pred, ncor
        self.shape = (B * W, MAX_TASKS_TRAIN, MAX_TASKS_TRAIN)

    def materialize(self):
        w, b = self._model.glove_net.table_params()
        with torch.no_grad():
            return _HeadFn.apply(self._emb.detach(), w, b, self.B, self.W, False, True)[3]


# ------------------------------------------------------------------------------------ Model
class Model(nn.Module):
    """CLIP-style contrastive classifier over 41 grasp classes (models.py:66-228)."""

    def __init__(self, params, adabn=True, train_model=True, prediction=False, glove=False, device="cuda"):
        super().__init__()
        if glove:
            raise NotImplementedError("--glove is broken in the reference (models.py:417 vs 451: the 20-dim glove row "
                                      "is fed to Linear(256,128)); the glove tower is available through clip.ClipModel")
        self.params = params
        self.train_model = train_model
        self.adabn = adabn
        self.prediction = prediction
        self.glove = glove
        self.device = torch.device(device)

        self.emg_net = EMGNet(d_e=params['d_e'], dp=params['dp_emg'], adabn=adabn, prediction=prediction,
                              device=device)
        self.glove_net = GLOVENet(d_e=params['d_e'], dp=params['dp_glove'], adabn=adabn, prediction=prediction,
                                  device=device)
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1) / 0.07)   # never applied (models.py:129)
        self.to(self.device)

        self.correct_tr = []
        self.correct_v = []
        self.materialize_logits = False     # training: return a real logits tensor from forward()
        self.reset()

    # -- mode switches (models.py:87-104)
    def set_train(self):
        self.train_model = True
        self.train()
        self.reset()

    def set_test(self):
        self.train_model = False
        self.eval()
        self.reset()

    def set_val(self):
        self.set_test()

    def reset(self):
        self._pending = []          # per-loss() device results, resolved lazily (no per-step host sync)
        self._corrects = []
        self.voting = []
        self.y_pred = []
        self.y_true = []

    def encode_emg(self, EMG):
        return self.emg_net(EMG)

    def encode_glove(self, GLOVE, labels):
        return self.glove_net(GLOVE, labels)

    # -- forward (models.py:112-130)
    def forward(self, EMG, GLOVE, labels, subjects=None):
        """subjects (optional, beyond the reference's signature): (B,41) subject of every class row -> per-subject
        AdaBN (EMGNet.per_subject)."""
        if self.prediction:
            return self._forward_prediction(EMG, labels)
        B, T = EMG.shape[0], EMG.shape[1]
        W = EMG.shape[2]
        if T != MAX_TASKS_TRAIN:
            raise RuntimeError(f"expected {MAX_TASKS_TRAIN} class rows per group, got {T}")
        emb = self.emg_net.encode_flat(EMG, subjects)             # (B*41*W, 16), order (b, class, w)
        w, b = self.glove_net.table_params()
        want_grad = torch.is_grad_enabled() and self.training
        want_logits = (not self.training) or self.materialize_logits
        loss, pred, ncor, logits = _HeadFn.apply(emb, w, b, B, W, want_grad, want_logits)
        handle = FusedLogits(self, emb, B, W, loss, pred, ncor)
        if want_logits:
            logits._cp_handle = handle
            return logits
        return handle

    # -- --prediction mode (models.py:113-119: features = emg_net(EMG) normalised per row; 175-196: CE + accuracy)
    def _forward_prediction(self, EMG, labels):
        net = self.emg_net
        a7 = net.encode_flat(EMG)                                  # (N,512) trunk output
        if not self.training and VOTE:
            # the reference's own evaluation of this mode dies on `assert len(shape)==3, "wrong logit shape for val
            # time"` (models.py:178: EMGNet.forward returns 2-D logits in prediction mode, models.py:337)
            raise AssertionError("wrong logit shape for val time")
        lin1, bn, lin2 = net.last[0], _bn_leaf(net.last[2]), net.last[3]
        if self.adabn:
            bn_mode, rm, rv = _lib.BN_BATCH, None, None
        else:
            bn_mode = _lib.BN_BATCH_UPDATE if self.training else _lib.BN_RUNNING
            rm, rv = bn.running_mean, bn.running_var
            if self.training:
                bn.num_batches_tracked += 1
        cfg = {"bn_mode": bn_mode, "bn_rm": rm, "bn_rv": rv, "want_grad": torch.is_grad_enabled() and self.training,
               "tap": net.debug_tap}
        feats, loss, pred, ncor = _ClsHeadFn.apply(a7, labels, cfg, lin1.weight, lin1.bias, bn.weight, bn.bias, lin2.weight)
        feats._cp_cls = (loss, pred, ncor, a7.shape[0])
        return feats

    # -- loss + accuracy (+ vote) (models.py:132-208)
    def loss(self, logits, labels):
        if self.prediction:
            h = getattr(logits, "_cp_cls", None)
            if h is None:
                raise RuntimeError("prediction mode: pass the tensor Model.forward returned (the loss is fused with it)")
            loss, pred, ncor, n = h
            self._pending.append(("cls", ncor, n))
            return loss
        handle = logits if isinstance(logits, FusedLogits) else getattr(logits, "_cp_handle", None)
        if handle is not None:
            loss, pred, ncor, B, W = handle.loss, handle.pred, handle.ncor, handle.B, handle.W
        else:
            # foreign logits tensor: same loss from the materialised values
            W = self.emg_net.shape[2] if (not self.training and VOTE) else 1
            B = logits.shape[0] // W
            want_grad = torch.is_grad_enabled() and logits.requires_grad
            loss, pred, ncor = _LogitsLossFn.apply(logits, want_grad)
        if not self.training and VOTE:
            L = _lib.lib()
            votes = torch.empty((B, N_VOTES), dtype=torch.int32, device=pred.device)
            y_pred = torch.empty((B, MAX_TASKS_TRAIN), dtype=torch.int64, device=pred.device)
            _lib.check(L.cp_vote_eval(_lib.ptr(pred), B, W, N_VOTES, _lib.ptr(votes), _lib.ptr(y_pred),
                                      _lib.stream()), "cp_vote_eval")
            self._pending.append(("vote", votes, y_pred))
        else:
            self._pending.append(("train", ncor, None))
        return loss

    def assemble_last(self, lo, n_global):
        """Sharded evaluation (train._evaluate under torchrun): the result of the last loss() call covers groups
        [lo, lo + B_local) of a global batch of n_global groups; replace it by the whole batch's integer arrays
        (every rank then resolves the same accuracy / voting / y_pred a single GPU would)."""
        from . import dist as cpdist
        kind, a, b = self._pending[-1]
        a = cpdist.assemble_rows(a, lo, n_global)
        if b is not None:
            b = cpdist.assemble_rows(b, lo, n_global)
        self._pending[-1] = (kind, a, b)

    # -- accumulators.  The kernels return INTEGER counts; the float arithmetic of the reference
    #    (float32 running sum of count/41 per batch, models.py:134,166,170-172) is reproduced here.
    def _resolve(self):
        T = MAX_TASKS_TRAIN
        for kind, a, b in self._pending:
            if kind == "cls":                    # (prediction == labels).mean()  (models.py:190-195): float64 count / n
                self._corrects.append(float(np.float64(int(a.item())) / np.float64(b)))
                continue
            if kind == "train":
                counts = a.cpu().numpy().astype(np.float64)
            else:
                votes = a.cpu().numpy()
                y_pred = b.cpu().numpy()
                for g in range(votes.shape[0]):
                    self.voting.append(list(votes[g].astype(np.float64) / T))
                    self.y_pred.append(y_pred[g])
                    self.y_true.append(np.arange(T, dtype=np.int64))
                counts = votes[:, -1].astype(np.float64)
            per_group = (counts / T).astype(np.float32)
            acc = np.cumsum(per_group, dtype=np.float32)[-1]         # sequential float32 adds
            self._corrects.append(float(np.float32(acc / np.float32(len(per_group)))))
        self._pending = []

    @property
    def corrects(self):
        self._resolve()
        return self._corrects

    def correct(self):
        return np.array(self.corrects).mean()

    def correct_raw(self):
        return np.array(self.corrects)

    def voting_raw(self):
        self._resolve()
        return np.array(self.voting)

    def y_pred_raw(self):
        self._resolve()
        return np.array(self.y_pred)

    def y_true_raw(self):
        self._resolve()
        return np.array(self.y_true)

    def l2(self):
        """models.py:225-228: reg * sum of un-squared Frobenius norms (tiny; plain torch ops)."""
        if self.prediction:
            return self.emg_net.l2() * self.params['reg_emg']
        return self.glove_net.l2() * self.params['reg_glove'] + self.emg_net.l2() * self.params['reg_emg']


class _L2Fn(torch.autograd.Function):
    """sum_t ||W_t||_2 over a parameter list through cp_l2_forward / cp_l2_backward (K5): 3 launches per step
    instead of the ~8 per tensor of `reg = reg + torch.norm(p)` and its autograd graph."""

    @staticmethod
    def forward(ctx, *params):
        L = _lib.lib()
        n = len(params)
        dev = params[0].device
        ptrs = (ctypes.c_void_p * n)(*[_lib.ptr(p.detach(), torch.float32).value for p in params])
        sizes = (ctypes.c_int64 * n)(*[p.numel() for p in params])
        norms = torch.empty(n, dtype=torch.float32, device=dev)
        total = torch.empty((), dtype=torch.float32, device=dev)
        nb = L.cp_l2_workspace_bytes(n)
        if nb == 0:
            raise RuntimeError("cp_l2_workspace_bytes rejected the parameter list (1..32 tensors)")
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        _lib.check(L.cp_l2_forward(ptrs, sizes, n, _lib.ptr(norms), _lib.ptr(total), _lib.ptr(ws), nb, _lib.stream()),
                   "cp_l2_forward")
        ctx.save_for_backward(norms, *params)
        return total

    @staticmethod
    def backward(ctx, g_total):
        L = _lib.lib()
        norms, *params = ctx.saved_tensors
        n = len(params)
        grads = [torch.empty_like(p, memory_format=torch.contiguous_format) for p in params]
        ptrs = (ctypes.c_void_p * n)(*[_lib.ptr(p.detach(), torch.float32).value for p in params])
        gptrs = (ctypes.c_void_p * n)(*[_lib.ptr(g).value for g in grads])
        sizes = (ctypes.c_int64 * n)(*[p.numel() for p in params])
        g_total = g_total.contiguous()
        _lib.check(L.cp_l2_backward(ptrs, sizes, n, _lib.ptr(norms), _lib.ptr(g_total, torch.float32), 1.0, gptrs,
                                    _lib.stream()), "cp_l2_backward")
        return tuple(grads)


def _l2_of(module):
    """models.py:344-349 / 467-472: parameters whose name has neither 'bn' nor 'bias'."""
    ps = [p for name, p in module.named_parameters() if 'bn' not in name and 'bias' not in name]
    return _L2Fn.apply(*ps) if ps else 0


class EMGNet(nn.Module):
    """12-channel instantaneous sEMG -> d_e embedding (models.py:230-349)."""

    def __init__(self, d_e, dp=.5, adabn=True, train=True, prediction=False, device="cuda"):
        super().__init__()
        if d_e != 16 and not prediction:
            raise NotImplementedError("libcpros is built for d_e = 16 (train.py:182 des=[16])")
        self.device = torch.device(device)
        self.d_e = d_e
        self.dp = dp
        self.prediction = prediction
        self.adabn = adabn
        if adabn:
            self.bn1d_func, self.bn2d_func = AdaBatchNorm1d, AdaBatchNorm2d
            bn1 = lambda f: AdaBatchNorm1d(f, device=device)     # noqa: E731
            bn2 = lambda f: AdaBatchNorm2d(f, device=device)     # noqa: E731
        else:
            self.bn1d_func, self.bn2d_func = nn.BatchNorm1d, nn.BatchNorm2d
            bn1, bn2 = nn.BatchNorm1d, nn.BatchNorm2d

        # Same module tree (and therefore the same state-dict keys and RNG consumption at init)
        # as models.py:248-315.
        self.conv_emg = nn.Sequential(
            nn.Conv2d(1, 64, (3, 3), padding=(1, 1)), nn.ReLU(), bn2(64),
            nn.Conv2d(64, 64, (3, 3), padding=(1, 1)), nn.ReLU(), bn2(64),
            nn.Flatten())
        blocks = [nn.Linear(EMG_DIM * 64, 512), nn.ReLU(), bn1(512)]
        for i in range(6):
            blocks += [nn.Linear(512, 512), nn.ReLU(), bn1(512)]
            if i >= 2:
                blocks.append(nn.Dropout(self.dp))
        self.linear = nn.Sequential(*blocks)
        if prediction:                       # models.py:300-309
            self.bits = MAX_TASKS_TRAIN
            self.last = nn.Sequential(nn.Linear(512, 128), nn.ReLU(), bn1(128), nn.Linear(128, self.bits, bias=False))
        else:
            self.bits = self.d_e
            self.last = nn.Sequential(nn.Linear(512, self.d_e, bias=False))
        self.to(self.device)

        self.engine = _lib.ENGINE_TC         # tcgen05 GEMMs on the 3-product fp16 split (fp32-level accuracy); ENGINE_SIMT = fp32 FFMA
        self.dropout_seed = 0x5EED
        self._step = 0
        self.dropout_step = None             # device int64 counter mixed into the Philox key (graph.GraphedTrainStep)
        self.sync_bn = False                 # True (+ torch.distributed initialised): BatchNorm statistics over the
        self.process_group = None            # rows of EVERY rank (global-batch semantics) instead of rank-local ones
        self.ext_dropout_masks = None        # (4, N, 512) uint8 keep masks injected by parity tests
        self.debug_tap = None                # set to {} to keep the workspace for read_activation()
        self.shape = None
        # Per-subject AdaBN ("momentum = 0 and batch per subject in order to have adaptive normalization",
        # models.py:245 -- described there, never implemented): with per_subject = True and a subject id per class
        # row (EMG._cp_subjects from TaskWrapper(with_subjects=True), or the `subjects` argument) the batch statistics
        # of every BatchNorm are taken over the windows of ONE subject at a time; gamma / beta stay shared.
        self.per_subject = False
        self.segment_streams = 4             # subject segments are independent encoder passes: run them on this many streams
        self._seg_pool = None

    # ordered views of the module tree for the kernels
    def _convs(self):
        return [self.conv_emg[0], self.conv_emg[3]]

    def _linears(self):
        return [m for m in self.linear if isinstance(m, nn.Linear)]

    def _bns(self):
        out = [_bn_leaf(self.conv_emg[2]), _bn_leaf(self.conv_emg[5])]
        out += [_bn_leaf(m) for m in self.linear
                if isinstance(m, (AdaBatchNorm1d, nn.BatchNorm1d))]
        return out

    def kernel_params(self):
        c, l, b = self._convs(), self._linears(), self._bns()
        proj = None if self.prediction else self.last[0].weight        # trunk-only: the classifier head follows (cls.cuh)
        return ([c[0].weight, c[0].bias, c[1].weight, c[1].bias] + [m.weight for m in l] +
                [m.bias for m in l] + [proj] + [m.weight for m in b] + [m.bias for m in b])

    def encode_flat(self, EMG, subjects=None, raw=False):
        """(B,41,W,1,12) or anything reshapeable to (-1,12) -> (N,16) embeddings, row order unchanged
        (prediction mode: the (N,512) output of the 7th linear block).  subjects: see `per_subject`."""
        self.shape = EMG.shape
        if subjects is None and self.per_subject:
            subjects = getattr(EMG, "_cp_subjects", None)
            if subjects is None:
                raise RuntimeError("per-subject AdaBN needs the subject of every class row: TaskWrapper.with_subjects "
                                   "= True, or pass subjects=")
        x = EMG.reshape(-1, EMG_DIM)
        bns = self._bns()
        if self.adabn:
            bn_mode, rm, rv = _lib.BN_BATCH, None, None
        else:
            bn_mode = _lib.BN_BATCH_UPDATE if self.training else _lib.BN_RUNNING
            rm, rv = [m.running_mean for m in bns], [m.running_var for m in bns]
            if self.training:
                for m in bns:
                    m.num_batches_tracked += 1
        dp = float(self.dp) if self.training else 0.0
        self._step += 1
        # sample-sharded ranks hold different rows of the global batch: mix the rank into the Philox key so that they
        # do not all apply the same keep mask to their local rows
        from . import dist as cpdist
        seed = (self.dropout_seed * 1000003 + self._step) ^ (cpdist.rank() << 48)
        cfg = {"bn_mode": bn_mode, "engine": self.engine, "dropout_p": dp,
               "seed": seed & 0xFFFFFFFFFFFFFFFF,
               "ext_masks": self.ext_dropout_masks if dp > 0 else None,
               "dropout_step": self.dropout_step, "bn_rm": rm, "bn_rv": rv,
               "need_bwd": torch.is_grad_enabled() and self.training,
               "tap": self.debug_tap, "sync_bn": self._sync_active(), "group": self.process_group,
               "trunk_only": self.prediction,
               "ws_alloc": getattr(self, "ws_alloc", None)}
        if raw:                                  # step.LeanTrainStep: no autograd node, the caller keeps the context
            if subjects is not None:
                raise NotImplementedError("per-subject AdaBN runs through autograd (one encoder pass per subject)")
            ctx = PlainCtx()
            cfg["need_bwd"] = True
            with torch.no_grad():
                emb = _EncoderFn.forward(ctx, x, cfg, *self.kernel_params())
            return emb, ctx
        if subjects is not None:
            return self._encode_per_subject(x, subjects, cfg)
        return _EncoderFn.apply(x, cfg, *self.kernel_params())

    def _encode_per_subject(self, x, subjects, cfg):
        """Per-subject AdaBN: BatchNorm is the only coupling between the rows of a batch, so statistics per subject ==
        one independent encoder pass per subject segment with the shared weights (autograd sums the segments' weight
        gradients).  Rows are sorted by subject (stable), every segment runs the ordinary kernels on its contiguous
        slice -- segments round-robin over `segment_streams` streams, they are small -- and the embeddings return in
        the caller's row order.  One host sync per call (the segment sizes)."""
        if not self.adabn:
            raise RuntimeError("per-subject statistics are an AdaBN mode (batch statistics in train and eval); "
                               "running-statistics BatchNorm has one set of statistics")
        if cfg["sync_bn"] or cfg["tap"] is not None:
            raise RuntimeError("per-subject AdaBN does not combine with sync_bn / debug_tap")
        n = x.shape[0]
        subjects = torch.as_tensor(subjects, device=x.device).to(torch.long)
        if subjects.numel() != n:              # one id per class row (B,41) -> one per window (B,41,W)
            if n % subjects.numel() != 0:
                raise RuntimeError(f"{subjects.numel()} subject ids for {n} windows")
            subjects = subjects.reshape(-1, 1).expand(-1, n // subjects.numel())
        subjects = subjects.reshape(-1)
        order = torch.argsort(subjects, stable=True)
        sizes = [c for c in torch.bincount(subjects).tolist() if c > 0]
        if min(sizes) < 2:
            # nn.BatchNorm1d: "Expected more than 1 value per channel when training"
            raise ValueError("per-subject AdaBN: a subject has a single window in this batch")
        xs = x[order]
        masks = cfg["ext_masks"][:, order] if cfg["ext_masks"] is not None else None
        dev = x.device
        main = torch.cuda.current_stream(dev)
        P = max(1, min(int(self.segment_streams), len(sizes)))
        if P > 1 and (self._seg_pool is None or len(self._seg_pool) < P):
            self._seg_pool = [torch.cuda.Stream(device=dev) for _ in range(P)]
        ready = main.record_event() if P > 1 else None
        params = self.kernel_params()
        parts, a = [], 0
        for i, c in enumerate(sizes):
            seg_cfg = dict(cfg)
            seg_cfg["seed"] = (cfg["seed"] + 0x9E3779B97F4A7C15 * (i + 1)) & 0xFFFFFFFFFFFFFFFF   # own dropout stream
            if P > 1:
                s = self._seg_pool[i % P]
                s.wait_event(ready)
                with torch.cuda.stream(s):               # (the mask slice is copied on the segment's stream too)
                    seg_cfg["ext_masks"] = masks[:, a:a + c].contiguous() if masks is not None else None
                    out = _EncoderFn.apply(xs[a:a + c], seg_cfg, *params)
                out.record_stream(main)
            else:
                seg_cfg["ext_masks"] = masks[:, a:a + c].contiguous() if masks is not None else None
                out = _EncoderFn.apply(xs[a:a + c], seg_cfg, *params)
            parts.append(out)
            a += c
        if P > 1:
            for s in self._seg_pool[:P]:                 # every pool stream reads slices of xs / masks
                xs.record_stream(s)
                if masks is not None:
                    masks.record_stream(s)
                main.wait_stream(s)
        inv = torch.empty_like(order)
        inv[order] = torch.arange(n, device=dev)
        return torch.cat(parts)[inv]

    def _sync_active(self):
        import torch.distributed as dist
        return bool(self.sync_bn and dist.is_available() and dist.is_initialized()
                    and dist.get_world_size(self.process_group) > 1)

    def read_activation(self, stage, which=0):
        """Parity tap (tests): saved activation of BN stage 0..8 of the last training forward, in the
        reference's layout ((N,64,1,12) for the conv stages, (N,512) for the linear ones).  Needs
        `self.debug_tap = {}` before the forward."""
        t = self.debug_tap
        n = t["n"]
        shape = (n * 12, 64) if stage < 2 else (n, 512)
        dst = torch.empty(shape, dtype=torch.float32, device=t["ws"].device)
        _lib.check(_lib.lib().cp_encoder_read_activation(_lib.ptr(t["ws"]), t["ws"].numel(), n, ctypes.byref(t["opts"]),
                                                         stage, which, _lib.ptr(dst), _lib.stream()),
                   "cp_encoder_read_activation")
        if stage < 2:
            return dst.reshape(n, 12, 64).permute(0, 2, 1).reshape(n, 64, 1, 12)
        return dst

    def forward(self, EMG):
        """models.py:319-342 incl. the (B,41,W) -> (B*W,41) regrouping."""
        if self.prediction:
            raise RuntimeError("prediction mode: the classifier head is fused with its loss (cp_cls_forward_backward) "
                               "and needs the labels; call Model.forward(EMG, GLOVE, labels)")
        out = self.encode_flat(EMG)
        shape = self.shape
        out = out.reshape((shape[0], shape[1], shape[2], self.bits)).transpose(1, 2)
        return out.reshape((-1, shape[1], self.bits))

    def l2(self):
        return _l2_of(self)


class GLOVENet(nn.Module):
    """Class tower.  Default branch of the reference (models.py:457-458): a learnable 41 x d_e table,
    `Linear(41->d_e)(one_hot(label))`; the glove angles themselves are unused."""

    def __init__(self, d_e, dp=.5, adabn=True, train=True, prediction=False, device="cuda"):
        super().__init__()
        self.device = torch.device(device)
        self.d_e = d_e
        self.dp = dp
        self.prediction = prediction
        self.conv_glove = nn.Sequential(nn.Flatten())
        self.linear = nn.Sequential(nn.Flatten())
        self.bits = MAX_TASKS_TRAIN if prediction else self.d_e
        self.easy = nn.Sequential(nn.Linear(MAX_TASKS_TRAIN, self.d_e))
        if prediction:                       # models.py:413-421 (parameters only: never called without --glove)
            bn = AdaBatchNorm1d(128, device=device) if adabn else nn.BatchNorm1d(128)
            self.last = nn.Sequential(nn.Linear(512 // 2, 128), nn.ReLU(), bn, nn.Dropout(self.dp),
                                      nn.Linear(128, self.bits, bias=False))
        else:
            self.last = nn.Sequential(nn.Linear(512 // 2, self.bits, bias=False))   # unused in forward, still regularised
        self.to(self.device)

    def table_params(self):
        return self.easy[0].weight, self.easy[0].bias

    def forward(self, GLOVE, labels):
        """models.py:432-465.  Returns the class embeddings (B or B*25, 41, d_e); Model.forward does
        not call this (the table is read directly by the fused head)."""
        w, b = self.table_params()
        shape = GLOVE.shape
        out = (w.t() + b)[labels.reshape(-1)].reshape((shape[0], -1, self.bits))
        if not self.training and VOTE:
            out = out.reshape((shape[0], 1, shape[1], self.bits)).expand(-1, PREDICTION_WINDOW_SIZE, -1, -1)
            out = out.reshape(-1, shape[1], self.bits)
        return out

    def l2(self):
        return _l2_of(self)
