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
    backend = backend or os.environ.get("CP_DIST_BACKEND") or ("nccl" if use_cuda else "gloo")
    if use_cuda:
        ndev = torch.cuda.device_count()
        if local >= ndev:
            # more local ranks than GPUs: only gloo can put two ranks on one device (NCCL refuses); this is how the
            # multi-rank paths are exercised on a 1-GPU box
            if backend == "nccl":
                raise RuntimeError(f"LOCAL_RANK {local} but {ndev} GPU(s): NCCL needs one GPU per rank")
            local %= ndev
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, torch.device(f"cuda:{local}" if use_cuda else "cpu")


def world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def _needs_staging(t, group=None):
    """NCCL moves CUDA tensors only; a CPU tensor (a host-side permutation) is staged through the device."""
    return dist.get_backend(group) == "nccl" and not t.is_cuda


def broadcast_(t, src=0, group=None):
    """In-place broadcast of `t` from rank `src` (no-op in a single process).  Returns t."""
    if world_size() == 1:
        return t
    if _needs_staging(t, group):
        d = t.cuda()
        dist.broadcast(d, src=src, group=group)
        t.copy_(d)
    else:
        dist.broadcast(t, src=src, group=group)
    return t


def broadcast_module(module, src=0, group=None):
    """Every parameter and buffer of `module` takes rank `src`'s value: the replicas of sample-sharded training
    must start (and, after load_state_dict, restart) from ONE model.  Integer buffers (num_batches_tracked) included."""
    if world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            broadcast_(t.data, src=src, group=group)


def even_shard(n, r=None, w=None):
    """[lo, hi) of rank r when n items are dealt as evenly as possible to w ranks: sizes differ by at most one and,
    for n >= w, every rank gets at least one item (every rank then runs the same number of steps and collectives)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    return (r * n) // w, ((r + 1) * n) // w


def assemble_rows(local, lo, total, group=None):
    """Rows [lo, lo + len(local)) of a (total, ...) array live on this rank: returns the whole array on every rank
    (exact for integers: each row is written by exactly one rank, the rest contribute zeros to the sum)."""
    if world_size() == 1:
        return local
    full = torch.zeros((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    full[lo:lo + local.shape[0]] = local
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full


def _has_tensor_collectives(group=None):
    """all_gather_into_tensor / reduce_scatter_tensor exist for NCCL; gloo offers all_reduce / broadcast only."""
    return dist.get_backend(group) == "nccl"


def all_gather_rows(x, group=None):
    """(n, ...) per rank -> (world*n, ...) on every rank, rank-major."""
    w, r = dist.get_world_size(group), dist.get_rank(group)
    n = x.shape[0]
    out = torch.empty((w * n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if _has_tensor_collectives(group):
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    else:
        out.zero_()
        out[r * n:(r + 1) * n] = x
        dist.all_reduce(out, group=group)
    return out


def reduce_scatter_rows(full, group=None):
    """(world*n, ...) partials on every rank -> this rank's (n, ...) block of the sum."""
    w, r = dist.get_world_size(group), dist.get_rank(group)
    n = full.shape[0] // w
    if _has_tensor_collectives(group):
        out = torch.empty((n,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        dist.reduce_scatter_tensor(out, full.contiguous(), group=group)
        return out
    dist.all_reduce(full, group=group)
    return full[r * n:(r + 1) * n].clone()


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
        if self.average and dist.get_backend() == "nccl":
            dist.all_reduce(self._flat, op=dist.ReduceOp.AVG)          # the division happens inside the collective
        else:
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
