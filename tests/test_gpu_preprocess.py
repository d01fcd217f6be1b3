"""cp_emg_preprocess (offline band-pass -> moving RMS -> time_mask) bit-exact against the reference's own outputs
(tests/golden/preprocess.npz) and the scipy-restating oracle; statistics / normalisation within fp32 tolerance."""
import os

import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200 import preprocess as PP
from oracle import preprocess as OP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gp(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess.npz"))


def test_matches_reference_outputs_bit_exact(gp):
    raw = torch.from_numpy(gp["raw"]).cuda()
    assert np.array_equal(PP.preprocess_segments(raw, wrap=True).cpu().numpy(), gp["emg_wrap"])
    assert np.array_equal(PP.preprocess_segments(raw, wrap=False).cpu().numpy(), gp["emg_full"])
    assert np.array_equal(PP.time_mask(True), gp["time_mask"])


@pytest.mark.parametrize("n_seg,n_ch", [(1, 12), (7, 12), (300, 12), (5, 3), (2, 1)])
def test_matches_oracle_bit_exact(n_seg, n_ch):
    rs = np.random.RandomState(n_seg * 31 + n_ch)
    raw = (rs.randn(n_seg, OP.SEG_LEN, n_ch) * 10.0 ** rs.uniform(-6, -3, (n_seg, 1, n_ch))).astype(np.float32)
    raw[0, :, 0] = 0.0                                                    # a dead channel: rms == 0 everywhere
    got = PP.preprocess_segments(torch.from_numpy(raw).cuda()).cpu().numpy()
    picks = range(n_seg) if n_seg <= 7 else rs.choice(n_seg, 6, replace=False)
    for s in picks:
        assert np.array_equal(got[s], OP.preprocess_segment(raw[s])), s
    assert np.count_nonzero(got[0, :, 0]) == 0
    # arbitrary (unordered, repeated) index lists and the empty batch
    idx = np.array([7, 0, 7, 1999, 3], dtype=np.int32)
    got2 = PP.preprocess_segments(torch.from_numpy(raw[:1]).cuda(), idx=idx).cpu().numpy()
    assert np.array_equal(got2[0], OP.preprocess_segment(raw[0], idx=idx))
    assert PP.preprocess_segments(torch.empty(0, OP.SEG_LEN, n_ch, device="cuda")).shape == (0, 100, n_ch)


def test_leading_dims_and_linearity():
    """(people, tasks, reps, 2010, 12) in, (people, tasks, reps, 100, 12) out; the pipeline is positively homogeneous:
    scaling the raw signal by a power of two scales the RMS by the same factor exactly."""
    raw = PP.synthetic_raw(people=2, tasks=3, reps=2, seed=5)
    out = PP.preprocess_segments(raw)
    assert out.shape == (2, 3, 2, 100, 12) and torch.isfinite(out).all() and (out >= 0).all()
    assert torch.equal(PP.preprocess_segments(raw * 4.0), out * 4.0)
    assert torch.equal(PP.preprocess_segments(raw[1, 2, 0]), out[1, 2, 0])


def test_stats_and_normalisation_match_reference(gp):
    EMG = torch.from_numpy(gp["emg_wrap"]).cuda()
    mask = torch.tensor([True] * 5 + [False], device="cuda")             # the fixture's "training subset"
    mean, std = PP.fit_stats(EMG, mask)
    np.testing.assert_allclose(mean.cpu().numpy(), gp["stats_mean_perch"], rtol=1e-5)
    np.testing.assert_allclose(std.cpu().numpy(), gp["stats_std_perch"], rtol=1e-5)
    mean_c, std_c = PP.fit_stats(EMG, mask, complete=True)
    assert mean_c.dim() == 0 and std_c.shape == (12,)
    np.testing.assert_allclose(mean_c.item(), gp["stats_mean_complete"], rtol=1e-5)
    np.testing.assert_allclose(std_c.cpu().numpy(), gp["stats_std_complete"], rtol=1e-5)
    raw = torch.from_numpy(gp["raw"]).cuda()
    norm, _, _ = PP.build_emg_tensor(raw, mask)
    np.testing.assert_allclose(norm.cpu().numpy(), gp["normalized_perch"], rtol=1e-4, atol=1e-5)


def test_rejects_bad_input():
    with pytest.raises(RuntimeError):
        PP.preprocess_segments(torch.zeros(1, OP.SEG_LEN, 12, dtype=torch.float64, device="cuda"))
    with pytest.raises(RuntimeError):
        PP.preprocess_segments(torch.zeros(1, 100, 12, device="cuda"))   # segment shorter than the indices need
