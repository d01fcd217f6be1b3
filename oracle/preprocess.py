"""Oracle (TEST INFRASTRUCTURE): offline sEMG preprocessing of one (subject, stimulus, repetition) segment,
numpy restatement of /root/reference/code:

  load.py:85-101   get_stim_rep: first TOTAL_WINDOW_SIZE + 2*WINDOW_EDGE = 2010 samples x 12 channels, x 2**10,
                   band-pass, moving RMS, `emg_[self.time_mask]`
  utils.py:134-147 filter: scipy.signal.butter(4, (20, 450)/nyquist, 'bandpass') + lfilter per channel, the result
                   stored back INTO the input array (so it is rounded to the input dtype, float32 for NinaPro .mat)
  utils.py:151-156 moving_rms / rms: sqrt(uniform_filter1d(square(x), size=11, mode='nearest'))[5:-5]
  load.py:116      time_mask = np.arange(0, 2000, 20, dtype=np.uint8): the uint8 WRAPS, so the 100 "downsampled"
                   samples are rms[(20*j) % 256] -- positions 0..252 only, 36 of them taken twice.  Reference
                   quirk, reproduced (pass wrap=False for the evidently intended arange(0, 2000, 20)).
  utils.py:79-130  RunningStats: Welford over the per-window means of the training subset; normalize = (X-mean)/std

scipy's two C loops are restated sample by sample -- lfilter's direct-form-II-transposed update
`z[n] = (z[n+1] + x*b[n+1]) - y*a[n+1]` and uniform_filter1d's running sum `tmp += e[l+size-1] - e[l-1]; out = tmp/size`
over the edge-replicated line -- because the CUDA kernel is sequential in time in the same way and is compared
BIT-EXACTLY.  Pinned by tests/golden/preprocess.npz (outputs of the reference's own filter / rms / RunningStats on
seeded raw segments, written by tests/golden/make_golden.py).
"""
import numpy as np
from scipy import signal

HZ, FACTOR, RMS_WINDOW, WINDOW_EDGE, TOTAL_WINDOW_SIZE = 2000, 20, 11, 5, 2000
SEG_LEN = TOTAL_WINDOW_SIZE + 2 * WINDOW_EDGE
GAIN = 2.0 ** 10


def butter_bandpass(f=(20, 450), order=4, hz=HZ):
    """utils.py:135-143 -> (b, a) float64, len 2*order+1."""
    nyq = hz / 2
    return signal.butter(order, [f[0] / nyq, f[1] / nyq], btype="bandpass")


def time_mask(wrap=True):
    """load.py:116.  wrap=True: the reference's uint8 arange (indices mod 256)."""
    return np.arange(0, TOTAL_WINDOW_SIZE, FACTOR, dtype=np.uint8 if wrap else np.int64).astype(np.int64)


def lfilter_df2t(b, a, x):
    """scipy.signal.lfilter for 1-D x (float64 arithmetic, a[0] == 1), sample by sample."""
    nb = len(b)
    z = np.zeros(nb - 1)
    y = np.empty(len(x))
    for t, xv in enumerate(np.asarray(x, dtype=np.float64)):
        yv = z[0] + b[0] * xv
        for n in range(nb - 2):
            z[n] = (z[n + 1] + xv * b[n + 1]) - yv * a[n + 1]
        z[nb - 2] = xv * b[nb - 1] - yv * a[nb - 1]
        y[t] = yv
    return y


def uniform_filter1d_nearest(x, size):
    """scipy.ndimage.uniform_filter1d(x, size, mode='nearest') for 1-D x: double running sum, output in x.dtype."""
    L, h = len(x), size // 2
    e = np.concatenate([np.repeat(x[:1], h), x, np.repeat(x[-1:], size - 1 - h)]).astype(np.float64)
    out = np.empty(L)
    tmp = 0.0
    for k in range(size):
        tmp += e[k]
    out[0] = tmp / size
    for l in range(1, L):
        tmp += e[l + size - 1] - e[l - 1]
        out[l] = tmp / size
    return out.astype(x.dtype)


def preprocess_segment(raw, idx=None, n_samples=None):
    """raw (SEG_LEN, C) float32/float64 -> (len(idx), C) in raw.dtype.  n_samples: only the first n_samples rms
    positions are evaluated (the filters are causal / local, so later samples never influence earlier ones)."""
    idx = time_mask() if idx is None else np.asarray(idx)
    b, a = butter_bandpass()
    n_rms = (int(idx.max()) + 1) if n_samples is None else n_samples
    need = n_rms + 2 * WINDOW_EDGE                      # filtered samples that enter rms[0 .. n_rms)
    x = raw * raw.dtype.type(GAIN)
    out = np.empty((len(idx), raw.shape[1]), dtype=raw.dtype)
    for c in range(raw.shape[1]):
        f = lfilter_df2t(b, a, x[:need, c]).astype(raw.dtype)        # stored back into the input array's dtype
        sq = np.square(f)
        # full-line semantics: the right edge replication never reaches positions < SEG_LEN - 2*WINDOW_EDGE, so
        # evaluating a prefix with a padded tail gives the same values on that prefix
        pad = np.concatenate([sq, np.repeat(sq[-1:], RMS_WINDOW)])
        uf = uniform_filter1d_nearest(pad, RMS_WINDOW)[:need]
        rms = np.sqrt(uf)[WINDOW_EDGE:]
        out[:, c] = rms[idx]
    return out


def running_stats(windows):
    """utils.py:79-130 on a sequence of (W, C) windows: Welford over the per-window means.  -> (mean, std)."""
    counter = 0
    for X in windows:
        counter += 1
        m = X.mean(0)
        if counter == 1:
            old_mean = new_mean = m
            old_s = np.zeros_like(m)
            new_s = old_s
        else:
            new_mean = old_mean + (m - old_mean) / counter
            new_s = old_s + (m - old_mean) * (m - new_mean)
            old_mean, old_s = new_mean, new_s
    return new_mean, np.sqrt(new_s / (counter - 1))
