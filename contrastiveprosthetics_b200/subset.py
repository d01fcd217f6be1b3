"""Test-time class-subset evaluator (reference: README.md:11,15; inputs as dumped by
results.py:42-61; outputs as data/{mean,std,min,max}_grasp.xlsx).

The user picks a subset of grasp classes; the predicted label is the arg-max of the inner
products restricted to that subset, majority-voted over the 250 ms window.  The reference never
committed this stage; the semantics implemented here are the restated spec
of SURVEY.md section 8 (a13): subset = k random grasps + the rest class (label 40).

GPU path: cp_rank_rows once per logits tensor, then cp_subset_eval for any number of trials.
Trials are independent: with world_size > 1 they are split per rank with no communication
(`shard_trials`) and the integer counts are summed by the caller."""
import numpy as np
import torch

from . import _lib
from .constants import MAX_TASKS

REST_LABEL = MAX_TASKS - 1


def make_trials(sizes=range(1, MAX_TASKS), trials_per_size=144, seed=0):
    """masks (len(sizes)*trials_per_size, 41) uint8: k random grasps (labels 0..39) + rest."""
    rs = np.random.RandomState(seed)
    sizes = list(sizes)
    masks = np.zeros((len(sizes) * trials_per_size, MAX_TASKS), dtype=np.uint8)
    t = 0
    for k in sizes:
        if not 1 <= k <= MAX_TASKS - 1:
            raise ValueError("subset size counts grasps: 1..40")
        for _ in range(trials_per_size):
            masks[t, rs.choice(MAX_TASKS - 1, size=k, replace=False)] = 1
            masks[t, REST_LABEL] = 1
            t += 1
    return masks, np.repeat(np.array(sizes), trials_per_size)


def shard_trials(n_trials, rank, world_size):
    """Contiguous trial slice of this rank (144 trials over 8 GPUs = 18 each)."""
    per = (n_trials + world_size - 1) // world_size
    lo = min(rank * per, n_trials)
    return lo, min(lo + per, n_trials)


class SubsetEvaluator:
    """rank once, evaluate many.  logits: (G, 41, 41) CUDA fp32 with G = items*W in (item, w) order
    (what Model.forward returns in eval mode), or (items, W, 41, 41)."""

    def __init__(self, logits, window):
        L = _lib.lib()
        if logits.dim() == 3:
            logits = logits.reshape(-1, window, MAX_TASKS, MAX_TASKS)
        self.items, self.window = logits.shape[0], logits.shape[1]
        logits = logits.contiguous()
        n_rows = logits.numel() // MAX_TASKS
        self.order = torch.empty((n_rows, MAX_TASKS), dtype=torch.uint8, device=logits.device)
        _lib.check(L.cp_rank_rows(_lib.ptr(logits, torch.float32), n_rows, _lib.ptr(self.order), _lib.stream()),
                   "cp_rank_rows")

    def evaluate(self, masks):
        """masks (n_trials, 41) uint8 (numpy or tensor) -> (correct, total) int64 CUDA tensors."""
        L = _lib.lib()
        masks = torch.as_tensor(masks, dtype=torch.uint8).to(self.order.device).contiguous()
        n = masks.shape[0]
        correct = torch.empty(n, dtype=torch.int64, device=masks.device)
        total = torch.empty(n, dtype=torch.int64, device=masks.device)
        _lib.check(L.cp_subset_eval(_lib.ptr(self.order), self.items, self.window, _lib.ptr(masks), n,
                                    _lib.ptr(correct), _lib.ptr(total), _lib.stream()), "cp_subset_eval")
        return correct, total


def summarize(correct, total, sizes):
    """Per subset size: mean / std / min / max accuracy over trials (layout of data/*_grasp.xlsx)."""
    acc = np.asarray(correct, dtype=np.float64) / np.maximum(np.asarray(total, dtype=np.float64), 1)
    out = {}
    for k in np.unique(sizes):
        a = acc[sizes == k]
        out[int(k)] = {"mean": a.mean(), "std": a.std(), "min": a.min(), "max": a.max()}
    return out
