"""Drop-in `TaskWrapper`, `Glover` (runtime half), `RunningStats.normalize`, `torchize`
(reference: code/utils.py).  Row gathers go through cp_gather_norm (libcpros.so); there is a
batched fast path (`TaskWrapper.batches`) next to the per-item `__getitem__` the reference's
DataLoader uses."""
import ctypes

import numpy as np
import torch

from . import _lib
from .constants import GLOVE_DIM

_DEVICE = "cuda"


def set_default_device(device):
    """The reference hard-codes "cuda" (utils.py:19); host-logic tests redirect it to "cpu"."""
    global _DEVICE
    _DEVICE = device


def default_device():
    return torch.device(_DEVICE)


def torchize(X):
    """utils.py:18-19."""
    return torch.from_numpy(np.array(X)).to(default_device())


def gather_rows(src2d, idx, mean=None, std=None, n_ch=1):
    """dst[r] = (src2d[idx[r]] - mean) / std through cp_gather_norm.  src2d (rows, row_len) fp32 CUDA,
    idx any-shape int64 CUDA.  Returns (idx.numel(), row_len)."""
    L = _lib.lib()
    if src2d.dim() != 2:
        raise RuntimeError("gather_rows expects a 2-D source")
    idx = idx.reshape(-1).contiguous()
    if idx.dtype != torch.int64:
        idx = idx.to(torch.int64)
    n, row_len = idx.numel(), src2d.shape[1]
    dst = torch.empty((n, row_len), dtype=torch.float32, device=src2d.device)
    stat_len = 0 if mean is None else int(mean.numel())
    _lib.check(L.cp_gather_norm(_lib.ptr(src2d, torch.float32), src2d.shape[0], row_len, _lib.ptr(idx), n,
                                _lib.ptr(dst), _lib.ptr(mean), _lib.ptr(std), stat_len, n_ch,
                                _lib.ptr(_err_flag(src2d.device)), _lib.stream()), "cp_gather_norm")
    return dst


_ERR_FLAGS = {}


def _err_flag(device):
    """One sticky device-side out-of-range flag per device (the kernel only ever ORs into it): no per-call
    allocation or fill launch, no per-call host sync.  `check_gather_errors` reads it where a sync is affordable."""
    f = _ERR_FLAGS.get(device)
    if f is None:
        f = _ERR_FLAGS[device] = torch.zeros(1, dtype=torch.int32, device=device)
    return f


def check_gather_errors(device=None):
    """Raise IndexError if any gather since the last check saw an out-of-range row index (the reference's
    advanced indexing raises at once, load.py:256-273; here the check costs a sync, so it runs at epoch /
    split boundaries: TaskWrapper.batches end, set_train / set_val / set_test)."""
    for dev, f in list(_ERR_FLAGS.items()):
        if device is not None and dev != torch.device(device):
            continue
        if int(f.item()) != 0:
            f.zero_()
            raise IndexError("cp_gather_norm: a row index was out of range (rows outside the source table were "
                             "read as row 0)")


class RunningStats:
    """Only the part of utils.py:79-130 that is on the hot path: holding (mean, std) and the
    normalisation `(X - mean) / std` (utils.py:129-130), which the gather kernel fuses."""

    def __init__(self, mean, std, device=None):
        device = device or default_device()
        self._mean = torch.as_tensor(np.asarray(mean, dtype=np.float32)).reshape(-1).to(device)
        self._std = torch.as_tensor(np.asarray(std, dtype=np.float32)).reshape(-1).to(device)
        if self._mean.numel() != self._std.numel():
            # the shipped data/emg_mean.npy is a scalar and emg_std.npy per-channel
            n = max(self._mean.numel(), self._std.numel())
            self._mean = self._mean.expand(n).contiguous()
            self._std = self._std.expand(n).contiguous()

    def mean(self):
        return self._mean

    def std(self):
        return self._std

    def mean_std(self):
        return self._mean, self._std

    def normalize(self, X):
        """(X - mean)/std over the last (channel) axis, on the GPU via cp_gather_norm (identity gather)."""
        flat = X.reshape(-1, X.shape[-1]).contiguous()
        idx = torch.arange(flat.shape[0], device=flat.device, dtype=torch.int64)
        return gather_rows(flat, idx, self._mean, self._std, n_ch=X.shape[-1]).reshape(X.shape)


class Glover:
    """GPU-resident glove tensor (41, Dg, 20) and its row gather (utils.py:185-254, runtime half)."""

    def __init__(self, device=None):
        self.device = torch.device(device) if device is not None else default_device()
        self.GLOVE = None
        self.GLOVE_use = None
        self.D = 0

    def load_stored(self, path=None):
        from .constants import PATH_DIR
        self.GLOVE = torch.load(path or (PATH_DIR + 'data/glove.pt'), map_location=self.device)
        return self.GLOVE

    def load_valid(self, tasks_mask):
        if self.GLOVE is None:
            self.D, self.GLOVE_use = 0, None
            return
        tensor = self.GLOVE[tasks_mask]
        self.D = self.GLOVE.shape[1]
        self.GLOVE_use = tensor.reshape(-1, self.GLOVE.shape[-1]).contiguous()

    def __getitem__(self, idx):
        shape = tuple(idx.shape)
        return gather_rows(self.GLOVE_use, idx).reshape(shape + (self.GLOVE_use.shape[1],))


class TaskWrapper:
    """utils.py:21-76: item i -> one random window of EACH of the 41 classes.

    `emg_rand[t, i]` is a per-class random permutation plus the class offset t*D; the reference's
    DataLoader calls `__getitem__` once per item (B tiny index launches per step).  `batches()`
    does the same sampling with ONE gather launch per batch."""

    def __init__(self, dataset, with_glove=True):
        self.__dict__["dataset"] = dataset
        self.device = dataset.device
        self.with_glove = with_glove
        self.with_subjects = False         # True: batches carry the subject of every class row (EMG._cp_subjects)

    def return_rand(self, D):
        T = self.dataset.TASKS
        base = torch.arange(T, device=self.device, dtype=torch.long).reshape(T, 1) * D
        return torch.rand((T, D), device=self.device).argsort(dim=-1) + base

    def reset(self):
        from . import dist as cpdist
        self.emg_rand = self.return_rand(self.dataset.D)
        gd = self.dataset.glover.D
        self.glove_rand = self.return_rand(gd) if (self.with_glove and gd > 0) else None
        self.idx = torch.randperm(self.dataset.TASKS * self.dataset.D, device=self.device, dtype=torch.long)
        if cpdist.world_size() > 1:
            # sample sharding: every rank must index the SAME per-class permutations (rank 0's)
            for t in (self.emg_rand, self.glove_rand, self.idx):
                if t is not None:
                    cpdist.broadcast_(t)

    def __getattr__(self, name):
        return getattr(self.__dict__["dataset"], name)

    def __len__(self):
        return self.dataset.D

    def _labels(self, B=None):
        lab = torch.arange(self.dataset.TASKS, device=self.device, dtype=torch.long)
        return lab if B is None else lab.unsqueeze(0).expand(B, -1).contiguous()

    def __getitem__(self, idx):
        tensor_emg = self.dataset[self.emg_rand[:, idx]]
        if self.glove_rand is not None:
            tensor_glove = self.dataset.glover[self.glove_rand[:, idx % self.dataset.glover.D]]
        else:
            tensor_glove = torch.zeros((self.dataset.TASKS, GLOVE_DIM), device=self.device)
        return tensor_emg, tensor_glove, self._labels()

    def _rand_by_item(self, name):
        """(D,41) item-major copy of the (41,D) permutation table `name`, rebuilt whenever the table object changes
        (reset(), or a test installing its own): a batch then gathers B contiguous rows instead of 41 x B strided ids
        (the strided gather + transpose was 35 us per step at B = 4096)."""
        table = getattr(self, name)
        cache = self.__dict__.setdefault("_by_item", {})
        hit = cache.get(name)
        if hit is None or hit[0] is not table:
            hit = cache[name] = (table, table.t().contiguous())
        return hit[1]

    def get_batch(self, items):
        """items: (B,) int64 tensor of item ids -> (EMG (B,41,W,1,12), GLOVE (B,41,20), label (B,41))."""
        items = items.to(self.device)
        B = items.numel()
        rows = self._rand_by_item("emg_rand")[items]                        # (B,41): one contiguous 41-id row per item
        EMG = self.dataset[rows]                                            # one launch
        if self.with_subjects:
            EMG._cp_subjects = self.dataset.subjects_of(rows)              # (B,41) -> per-subject AdaBN (models.py:245)
        if self.glove_rand is not None:
            grow = self._rand_by_item("glove_rand")[items % self.dataset.glover.D]
            GLOVE = self.dataset.glover[grow]
        else:
            GLOVE = torch.zeros((B, self.dataset.TASKS, GLOVE_DIM), device=self.device)
        return EMG, GLOVE, self._labels(B)

    def get_flat_batch(self, start, n):
        """Flat sampling of the batch x batch variant (utils.py:56-59, commented in the reference): `n`
        consecutive entries of the epoch's flat row permutation `self.idx` -> one window per sample,
        label = row // D, and a glove row of the SAME class.  Returns (EMG (n,1,1,12), GLOVE (n,dim), label (n,))."""
        rows = self.idx[start:start + n]
        D = self.dataset.D
        label = rows // D
        EMG = self.dataset.slice_batch(rows)
        gl = self.dataset.glover
        if gl.GLOVE_use is not None:
            GLOVE = gl[label * gl.D + (rows % D) % gl.D]
        else:
            GLOVE = torch.zeros((rows.numel(), GLOVE_DIM), device=self.device)
        return EMG, GLOVE, label

    def batch_plan(self, batch_size, shuffle=True, generator=None, rank=0, world_size=1):
        """Item ids of every batch of one epoch for `rank`: list of (items, lo, n_global) with `items` this rank's
        slice [lo, lo + len) of a global batch of n_global items.  Every rank runs the SAME number of batches (equal
        collective counts): a global batch is dealt out evenly (sizes differ by at most one), and a ragged last
        batch with fewer items than ranks is dropped on all of them."""
        from . import dist as cpdist
        D = len(self)
        order = torch.randperm(D, generator=generator) if shuffle else torch.arange(D)
        if world_size > 1 and shuffle:
            cpdist.broadcast_(order)                # one permutation for the job, whatever the ranks' RNG states
        plan = []
        for s in range(0, D, batch_size):
            chunk = order[s:s + batch_size]
            n = chunk.numel()
            if world_size > 1:
                if n < world_size:
                    continue
                lo, hi = cpdist.even_shard(n, rank, world_size)
                plan.append((chunk[lo:hi], lo, n))
            else:
                plan.append((chunk, 0, n))
        return plan

    def batches(self, batch_size, shuffle=True, generator=None, rank=0, world_size=1, with_span=False):
        """Equivalent of `DataLoader(self, batch_size, shuffle)` (train.py:86): a permutation of the
        D items cut into batches (last one ragged).  With world_size > 1 every rank uses the same
        permutation and takes a disjoint slice of each global batch (sample sharding).  with_span: also
        yield (lo, n_global), the position of the slice inside its global batch."""
        for items, lo, n in self.batch_plan(batch_size, shuffle, generator, rank, world_size):
            b = self.get_batch(items)
            yield (b + (lo, n)) if with_span else b
        check_gather_errors(self.device)

    def set_train(self):
        self.dataset.set_train()
        self.reset()

    def set_val(self):
        self.dataset.set_val()
        self.reset()

    def set_test(self):
        self.dataset.set_test()
        self.reset()
