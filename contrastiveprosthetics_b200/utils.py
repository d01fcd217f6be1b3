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
    err = torch.zeros(1, dtype=torch.int32, device=src2d.device)
    _lib.check(L.cp_gather_norm(_lib.ptr(src2d, torch.float32), src2d.shape[0], row_len, _lib.ptr(idx), n,
                                _lib.ptr(dst), _lib.ptr(mean), _lib.ptr(std), stat_len, n_ch,
                                _lib.ptr(err), _lib.stream()), "cp_gather_norm")
    dst._cp_err = err        # checked lazily by callers that can afford a sync
    return dst


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

    def return_rand(self, D):
        T = self.dataset.TASKS
        base = torch.arange(T, device=self.device, dtype=torch.long).reshape(T, 1) * D
        return torch.rand((T, D), device=self.device).argsort(dim=-1) + base

    def reset(self):
        self.emg_rand = self.return_rand(self.dataset.D)
        gd = self.dataset.glover.D
        self.glove_rand = self.return_rand(gd) if (self.with_glove and gd > 0) else None
        self.idx = torch.randperm(self.dataset.TASKS * self.dataset.D, device=self.device, dtype=torch.long)

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

    def get_batch(self, items):
        """items: (B,) int64 tensor of item ids -> (EMG (B,41,W,1,12), GLOVE (B,41,20), label (B,41))."""
        items = items.to(self.device)
        B = items.numel()
        rows = self.emg_rand[:, items].t().contiguous()                     # (B,41)
        EMG = self.dataset[rows]                                            # one launch
        if self.glove_rand is not None:
            grow = self.glove_rand[:, items % self.dataset.glover.D].t().contiguous()
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

    def batches(self, batch_size, shuffle=True, generator=None, rank=0, world_size=1):
        """Equivalent of `DataLoader(self, batch_size, shuffle)` (train.py:86): a permutation of the
        D items cut into batches (last one ragged).  With world_size > 1 every rank draws the same
        permutation and takes a disjoint slice of each global batch (sample sharding)."""
        D = len(self)
        order = torch.randperm(D, generator=generator) if shuffle else torch.arange(D)
        for s in range(0, D, batch_size):
            chunk = order[s:s + batch_size]
            if world_size > 1:
                per = (chunk.numel() + world_size - 1) // world_size
                chunk = chunk[rank * per:(rank + 1) * per]
                if chunk.numel() == 0:
                    continue
            yield self.get_batch(chunk)

    def set_train(self):
        self.dataset.set_train()
        self.reset()

    def set_val(self):
        self.dataset.set_val()
        self.reset()

    def set_test(self):
        self.dataset.set_test()
        self.reset()
