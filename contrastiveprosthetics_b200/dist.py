"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in
CPU tests).  Only what the hot path needs:

  * sample-sharded training: every rank holds a replica of the model and of the 54 MB sEMG tensor,
    takes a disjoint slice of each global batch (TaskWrapper.batches(rank=, world_size=)) and the
    gradients are averaged with ONE flat all-reduce per step (2,027,616 fp32 = 8.1 MB);
    BatchNorm statistics stay local to the rank ("local-BN": equals the average of world_size
    independent reference replicas at batch B/world_size).
  * subset trials / cross-validation folds: split per rank, no data-path collective; integer
    counts are summed at the end.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Read RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world_size, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world)
    return rank, world, torch.device(f"cuda:{local}" if use_cuda else "cpu")


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


class FlatGradAllReduce:
    """Average (or, average=False, sum) the gradients of `params` across ranks with one all-reduce over a
    flat fp32 bucket.  Sum is for losses that are already normalised by the GLOBAL batch (clip.clip_head)."""

    def __init__(self, params, average=True):
        self.average = average
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self._flat = None

    def __call__(self):
        if world_size() == 1:
            return
        ps = [p for p in self.params if p.grad is not None]
        if not ps:
            return
        n = sum(p.numel() for p in ps)
        if self._flat is None or self._flat.numel() != n or self._flat.device != ps[0].device:
            self._flat = torch.empty(n, dtype=torch.float32, device=ps[0].device)
        off = 0
        views = []
        for p in ps:
            v = self._flat[off:off + p.numel()].view_as(p)
            views.append(v)
            off += p.numel()
        torch._foreach_copy_(views, [p.grad for p in ps])
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        if self.average:
            self._flat.div_(world_size())
        torch._foreach_copy_([p.grad for p in ps], views)


def sum_counts(*tensors):
    """Exact integer reduction of per-rank counts (subset trials, vote counts)."""
    if world_size() == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors


def shard_range(n, r=None, w=None):
    r = rank() if r is None else r
    w = world_size() if w is None else w
    per = (n + w - 1) // w
    lo = min(r * per, n)
    return lo, min(lo + per, n)
