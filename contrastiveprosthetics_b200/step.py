"""The loop body of the reference's train_loop (code/train.py:95-108) without autograd:

    logits = model.forward(EMG, GLOVE, label)          cp_encoder_forward + cp_head_forward_backward
    loss   = model.loss(logits, label) + model.l2()    (loss and its gradients come out of the fused head; the
    loss.backward()                                     regulariser only contributes its gradient, below)
                                                        cp_encoder_backward
    optimizer_emg.step(); optimizer_glove.step()        cp_adam_step: ONE launch for both Adams (train.py:72-73),
                                                        with d(reg * ||W||_2)/dW added on the fly

Through torch's autograd and torch.optim the same step carries ~35 extra launches of a few microseconds (norm / mul /
add / fill kernels, one gradient accumulation per regularised tensor, five multi-tensor Adam launches): a quarter of
the CUDA-graph nodes of the batch_size-8 step of go.sh:6.  `LeanTrainStep` replaces them by `cp_step_prologue` (parameter
norms + step counters) and `cp_adam_step`; the parameter gradients live in ONE flat bucket that the library writes in
place, so sample-sharded training (one process per GPU) all-reduces that bucket directly -- no pack / unpack.

Same results as the autograd step: identical loss and gradients (the regulariser's gradient is added with the same two
roundings), Adam in the arithmetic of torch's single-kernel implementation (tests/test_gpu_step.py).
"""
import ctypes

import torch

from . import _lib
from . import dist as cpdist
from .constants import MAX_TASKS_TRAIN
from .models import _encoder_backward_into, _head_launch

_ALIGN = 128            # elements: every tensor of the flat buckets starts on a 512-byte boundary, like a torch allocation


def _is_regularised(name):
    """models.py:344-349 / 467-472: parameters whose name has neither 'bn' nor 'bias'."""
    return 'bn' not in name and 'bias' not in name


class LeanTrainStep:
    def __init__(self, model, lr_emg, lr_glove, betas=(0.9, 0.999), eps=1e-8, sync_grads=False, group=None):
        """model: models.Model in the contrastive (default) mode; lr_*, betas, eps: the two torch.optim.Adam of
        train.py:72-73 (weight_decay 0).  sync_grads: average the gradient bucket over the ranks of `group` every step
        (sample-sharded training; NCCL's AVG, or SUM + divide on gloo)."""
        if model.prediction:
            raise NotImplementedError("LeanTrainStep covers the contrastive head; --prediction steps through autograd")
        self.model, self.group = model, group
        self.sync_grads = bool(sync_grads) and cpdist.world_size() > 1
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        net, glove = model.emg_net, model.glove_net
        dev = next(net.parameters()).device
        named = [("emg", n, p) for n, p in net.named_parameters()] + [("glove", n, p) for n, p in glove.named_parameters()]
        self.params = [p for _, _, p in named]
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("LeanTrainStep needs contiguous float32 parameters")
        n = len(self.params)
        if n > 48:
            raise RuntimeError("more than CP_STEP_MAX_TENSORS parameter tensors")
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        self.grad_flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self.grad_views = [self.grad_flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, self.params)]
        for p, g in zip(self.params, self.grad_views):
            p.grad = g                                   # what a caller inspecting .grad (or clipping it) sees
        self._view_of = {id(p): g for p, g in zip(self.params, self.grad_views)}
        # counters[0]: dropout step (cp_encoder_opts.dropout_step), counters[1]: Adam's t
        self.counters = torch.zeros(2, dtype=torch.int64, device=dev)
        net.dropout_step = self.counters[0:1]
        self.lr = torch.tensor([lr_emg, lr_glove], dtype=torch.float64, device=dev)
        # regularised tensors (both nets, one norm pass)
        reg_idx = [i for i, (_, name, _) in enumerate(named) if _is_regularised(name)]
        self.norms = torch.zeros(max(1, len(reg_idx)), dtype=torch.float32, device=dev)
        L = _lib.lib()
        nb = L.cp_step_workspace_bytes(len(reg_idx))
        self._ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
        self._n_reg = len(reg_idx)
        self._reg_ptrs = (ctypes.c_void_p * len(reg_idx))(*[self.params[i].data_ptr() for i in reg_idx])
        self._reg_sizes = (ctypes.c_int64 * len(reg_idx))(*[self.params[i].numel() for i in reg_idx])
        self._p_ptrs = (ctypes.c_void_p * n)(*[p.data_ptr() for p in self.params])
        self._sizes = (ctypes.c_int64 * n)(*[p.numel() for p in self.params])
        self._offs = (ctypes.c_int64 * n)(*offs)
        self._lr_index = (ctypes.c_int32 * n)(*[0 if which == "emg" else 1 for which, _, _ in named])
        self._which = [which for which, _, _ in named]
        self._norm_index = (ctypes.c_int32 * n)(*[reg_idx.index(i) if i in reg_idx else -1 for i in range(n)])
        self._reg = (ctypes.c_float * n)()
        self.set_reg(model.params['reg_emg'], model.params['reg_glove'])

    # -- hyper-parameters a scheduler / a fold may change between steps
    def set_lr(self, lr_emg, lr_glove):
        """lr lives on the device (read by cp_adam_step at run time): effective for graph replays too."""
        self.lr.copy_(torch.tensor([lr_emg, lr_glove], dtype=torch.float64))

    def set_reg(self, reg_emg, reg_glove):
        """reg_* are launch arguments: changing them needs a new capture of a graph that contains the step."""
        for i, which in enumerate(self._which):
            self._reg[i] = float(reg_emg if which == "emg" else reg_glove)

    def state_tensors(self):
        return [self.exp_avg, self.exp_avg_sq, self.counters]

    def zero_state(self):
        for t in self.state_tensors():
            t.zero_()

    # -- the step
    def body(self, EMG):
        """One training step on a (B,41,1,1,12) batch.  Returns (loss, per-group correct counts) like
        graph.GraphedTrainStep.  Capturable: no host synchronisation, no allocation outside torch's caching allocator."""
        m = self.model
        net = m.emg_net
        if not m.training:
            raise RuntimeError("LeanTrainStep: model.set_train() first")
        if EMG.dim() != 5 or EMG.shape[1] != MAX_TASKS_TRAIN:
            raise RuntimeError(f"expected a (B, {MAX_TASKS_TRAIN}, W, 1, 12) batch, got {tuple(EMG.shape)}")
        L = _lib.lib()
        P = _lib.ptr
        _lib.check(L.cp_step_prologue(self._reg_ptrs, self._reg_sizes, self._n_reg, P(self.norms), P(self.counters), 2,
                                      P(self._ws), self._ws.numel(), _lib.stream()), "cp_step_prologue")
        B, W = EMG.shape[0], EMG.shape[2]
        emb, ctx = net.encode_flat(EMG, raw=True)
        w, b = m.glove_net.table_params()
        loss, pred, ncor, _, d_emb, _, _ = _head_launch(emb, w, b, B, W, True, False,
                                                        d_w=self._view_of[id(w)], d_b=self._view_of[id(b)])
        grads = [self._view_of[id(p)] if p is not None else None for p in net.kernel_params()]
        _encoder_backward_into(ctx, d_emb, grads)
        if self.sync_grads:
            self._all_reduce()
        _lib.check(L.cp_adam_step(self._p_ptrs, self._sizes, self._offs, len(self.params), P(self.grad_flat),
                                  P(self.exp_avg), P(self.exp_avg_sq), P(self.lr), self._lr_index, self._reg,
                                  self._norm_index, P(self.norms), P(self.counters[1:2]), self.betas[0], self.betas[1],
                                  self.eps, _lib.stream()), "cp_adam_step")
        m._pending.append(("train", ncor, None))
        return loss, ncor

    __call__ = body

    def _all_reduce(self):
        import torch.distributed as dist
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.grad_flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(self.grad_flat, op=dist.ReduceOp.SUM, group=self.group)
            self.grad_flat.div_(dist.get_world_size(self.group))


def from_optimizers(model, optimizers, sync_grads=False, group=None):
    """LeanTrainStep with the hyper-parameters of train.py:72-73's two torch.optim.Adam instances (emg, glove).  The
    torch optimizers themselves are not stepped afterwards."""
    if len(optimizers) != 2:
        raise RuntimeError("expected [optimizer_emg, optimizer_glove]")
    hp = []
    for o in optimizers:
        if not isinstance(o, torch.optim.Adam) or len(o.param_groups) != 1:
            raise RuntimeError("LeanTrainStep mirrors torch.optim.Adam with one parameter group per optimizer")
        g = o.param_groups[0]
        if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
            raise RuntimeError("LeanTrainStep: Adam(weight_decay=0, amsgrad=False, maximize=False) only (train.py:72-73)")
        hp.append((float(g["lr"]), tuple(g["betas"]), float(g["eps"])))
    if hp[0][1:] != hp[1][1:]:
        raise RuntimeError("the two Adams must share betas / eps")
    return LeanTrainStep(model, hp[0][0], hp[1][0], betas=hp[0][1], eps=hp[0][2], sync_grads=sync_grads, group=group)
