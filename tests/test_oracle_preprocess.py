"""Offline-preprocessing oracle (oracle/preprocess.py) against tests/golden/preprocess.npz = the outputs of the
reference's own filter / rms / time_mask / RunningStats (load.py:85-116, utils.py:79-156).  CPU only."""
import os

import numpy as np
import pytest

from oracle import preprocess as OP


@pytest.fixture(scope="module")
def gp(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess.npz"))


def test_time_mask_wraps_like_the_reference(gp):
    assert np.array_equal(OP.time_mask(True), gp["time_mask"])            # uint8 arange: 20*j mod 256
    assert OP.time_mask(True).max() == 252 and len(np.unique(OP.time_mask(True))) == 64
    assert np.array_equal(OP.time_mask(False), np.arange(0, 2000, 20))


def test_segments_bit_exact(gp):
    raw = gp["raw"]
    for s in range(raw.shape[0]):
        assert np.array_equal(OP.preprocess_segment(raw[s]), gp["emg_wrap"][s]), s
    for s in (0, 3):                                                       # un-wrapped indices: the whole 2010 samples
        assert np.array_equal(OP.preprocess_segment(raw[s], idx=OP.time_mask(False)), gp["emg_full"][s]), s
    assert np.array_equal(OP.preprocess_segment(raw[0].astype(np.float64)), gp["emg_wrap_f64"])


def test_restated_scipy_loops_match_scipy():
    from scipy import signal
    from scipy.ndimage import uniform_filter1d
    rs = np.random.RandomState(3)
    x = rs.randn(400)
    b, a = OP.butter_bandpass()
    assert np.array_equal(OP.lfilter_df2t(b, a, x), signal.lfilter(b, a, x))
    for dt in (np.float32, np.float64):
        v = np.square(rs.randn(300)).astype(dt)
        assert np.array_equal(OP.uniform_filter1d_nearest(v, 11), uniform_filter1d(v, size=11, mode="nearest"))


def test_running_stats(gp):
    mean, std = OP.running_stats(list(gp["emg_wrap"][:5]))
    np.testing.assert_allclose(mean, gp["stats_mean_perch"], rtol=1e-5)
    np.testing.assert_allclose(std, gp["stats_std_perch"], rtol=1e-5)
    # complete=True: scalar mean, but the std stays per channel (utils.py:113-124) -- like the shipped data/emg_*.npy
    assert gp["stats_mean_complete"].shape == () and gp["stats_std_complete"].shape == (12,)
    np.testing.assert_allclose(mean.mean(), gp["stats_mean_complete"], rtol=1e-5)
    np.testing.assert_allclose((gp["emg_wrap"] - mean) / std, gp["normalized_perch"], rtol=1e-4, atol=1e-5)


def test_host_module_agrees_with_the_oracle_on_the_cpu_side():
    """contrastiveprosthetics_b200/preprocess.py's host-side pieces (filter design, index list, constants) -- no GPU."""
    from contrastiveprosthetics_b200 import preprocess as PP
    b0, a0 = OP.butter_bandpass()
    b1, a1 = PP.butter_bandpass()
    assert np.array_equal(b0, b1) and np.array_equal(a0, a1) and len(b1) == 9 and a1[0] == 1.0
    assert np.array_equal(PP.time_mask(True), OP.time_mask(True)) and np.array_equal(PP.time_mask(False), OP.time_mask(False))
    assert PP.SEG_LEN == OP.SEG_LEN == 2010 and PP.GAIN == OP.GAIN == 1024.0
    import torch
    with pytest.raises(RuntimeError):                       # no CPU fallback: host tensors are refused
        PP.preprocess_segments(torch.zeros(1, PP.SEG_LEN, 12))
