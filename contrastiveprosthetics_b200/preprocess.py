"""Offline sEMG preprocessing on the GPU (reference: code/load.py:85-155, code/utils.py:79-156; SURVEY.md
section 8f row 4 -- the step BEFORE the hot path: raw 2 kHz recordings -> the resident `emg.pt` tensor).

    raw segment (2010 samples x 12 ch, float32)  ->  x 2**10  ->  Butterworth order-4 band-pass 20-450 Hz (lfilter)
      ->  moving RMS over 11 samples  ->  rms[time_mask]  (100 samples)          cp_emg_preprocess, bit-exact with scipy
    all segments  ->  EMG (46, 41, 6, 100, 12)  ->  RunningStats over the training subset  ->  (EMG - mean) / std

`.mat` parsing (load.py:78-83, the `restimulus` / `rerepetition` masks of get_stim_rep) stays outside: the dataset is
not available offline, so callers hand in the extracted segments (`synthetic_raw` makes NinaPro-shaped ones).  The
filter design (`scipy.signal.butter`, nine coefficients) is host-side plumbing exactly as in utils.py:143.

Reference quirks kept (and switchable): `time_mask` is a uint8 arange, so it wraps at 256 (load.py:116) -- pass
`wrap=False` for the evidently intended arange(0, 2000, 20); `RunningStats(complete=True)` averages the mean to a
scalar but returns the per-channel std (utils.py:99-124), which is what the shipped data/emg_{mean,std}.npy hold.
"""
import ctypes

import numpy as np
import torch
from scipy import signal

from . import _lib
from .constants import (EMG_DIM, FACTOR, Hz, MAX_REPS, MAX_TASKS, RMS_WINDOW, TOTAL_WINDOW_SIZE, WINDOW_EDGE)

SEG_LEN = TOTAL_WINDOW_SIZE + 2 * WINDOW_EDGE          # load.py:93
GAIN = 2.0 ** 10                                       # load.py:96


def butter_bandpass(f=(20, 450), order=4):
    """utils.py:134-143."""
    nyq = Hz / 2
    return signal.butter(order, [f[0] / nyq, f[1] / nyq], btype="bandpass")


def time_mask(wrap=True):
    """load.py:116 (uint8 arange, wraps at 256) or, wrap=False, the un-wrapped index list."""
    return np.arange(0, TOTAL_WINDOW_SIZE, FACTOR, dtype=np.uint8 if wrap else np.int64).astype(np.int32)


def preprocess_segments(raw, wrap=True, idx=None):
    """raw (..., SEG_LEN, n_ch) float32 CUDA -> (..., len(idx), n_ch) float32: filter -> rms -> rms[idx]."""
    L = _lib.lib()
    if raw.dtype != torch.float32:
        raise RuntimeError("raw segments must be float32 (the dtype of the NinaPro .mat recordings)")
    lead, (seg_len, n_ch) = raw.shape[:-2], raw.shape[-2:]
    raw = raw.reshape(-1, seg_len, n_ch).contiguous()
    n_seg = raw.shape[0]
    idx = time_mask(wrap) if idx is None else np.asarray(idx, dtype=np.int32)
    n_rms = int(idx.max()) + 1
    b, a = butter_bandpass()
    cb = (ctypes.c_double * len(b))(*b)
    ca = (ctypes.c_double * len(a))(*a)
    idx_dev = torch.from_numpy(idx).to(raw.device)
    out = torch.empty((n_seg, len(idx), n_ch), dtype=torch.float32, device=raw.device)
    n_scratch = L.cp_emg_preprocess_scratch_elems(n_seg, n_ch, n_rms)
    scratch = torch.empty(max(n_scratch, 1), dtype=torch.float32, device=raw.device)
    _lib.check(L.cp_emg_preprocess(_lib.ptr(raw), n_seg, seg_len, n_ch, cb, ca, len(b), GAIN, RMS_WINDOW, n_rms,
                                   _lib.ptr(idx_dev, torch.int32), len(idx), _lib.ptr(out), _lib.ptr(scratch),
                                   scratch.numel(), _lib.stream()), "cp_emg_preprocess")
    return out.reshape(*lead, len(idx), n_ch)


def fit_stats(EMG, train_mask, complete=False):
    """utils.py:79-124 over the windows selected by `train_mask` (bool, EMG.shape[:-2]): Welford on the per-window
    means == mean / unbiased std of those means.  complete=True: scalar mean, per-channel std (reference quirk)."""
    means = EMG[train_mask].mean(-2)                    # (n_windows, n_ch): X.mean(0) of every pushed window
    mean, std = means.mean(0), means.std(0, unbiased=True)
    if complete:
        mean = mean.mean()
    return mean, std


def build_emg_tensor(raw, train_mask, wrap=True, complete=False):
    """load_dataset (load.py:103-147): raw (people, tasks, reps, SEG_LEN, 12) float32 CUDA -> normalised EMG
    (people, tasks, reps, 100, 12) float32 + (mean, std), ready for `torch.save` as emg.pt / DB23.load_tensors."""
    EMG = preprocess_segments(raw, wrap=wrap)
    mean, std = fit_stats(EMG, train_mask, complete)
    return (EMG - mean) / std, mean, std


def synthetic_raw(people=2, tasks=MAX_TASKS, reps=MAX_REPS, seed=0, device="cuda"):
    """Seeded NinaPro-shaped raw recordings (volts-scale broadband noise + 50 Hz line + drift), float32."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    shape = (people, tasks, reps, SEG_LEN, EMG_DIM)
    t = torch.arange(SEG_LEN, dtype=torch.float64)[:, None] / Hz
    amp = 1e-5 * (1 + 4 * torch.rand(shape[:3] + (1, EMG_DIM), generator=g, dtype=torch.float64))
    x = amp * torch.randn(shape, generator=g, dtype=torch.float64) + 2e-5 * torch.sin(2 * np.pi * 50 * t)
    return x.to(torch.float32).to(device)
