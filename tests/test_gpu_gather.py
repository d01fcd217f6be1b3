"""K1 parity: cp_gather_norm through DB23 / TaskWrapper vs the oracle and the reference fixtures."""
import os

import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.synthetic import fixed_perm, synth_emg, synth_glove
from contrastiveprosthetics_b200.utils import RunningStats, TaskWrapper, check_gather_errors, gather_rows
from oracle import dataset as OD

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gd(golden_dir):
    return np.load(os.path.join(golden_dir, "dataset.npz"))


@pytest.fixture(scope="module")
def emg():
    return synth_emg()


@pytest.mark.parametrize("db2", [False, True])
@pytest.mark.parametrize("split", ["train", "val", "test"])
def test_items_match_reference_fixture(gd, emg, db2, split):
    ds = DB23(db2=db2, device="cuda")
    ds.load_tensors(emg, synth_glove())
    tw = TaskWrapper(ds)
    getattr(tw, "set_" + split)()
    tag = f"db2{int(db2)}_{split}"
    tw.emg_rand = torch.from_numpy(fixed_perm(41, ds.D, 11)).cuda()
    tw.glove_rand = torch.from_numpy(fixed_perm(41, ds.glover.D, 12)).cuda()
    e, g, l = tw[int(gd[tag + "_items"][1])]
    assert np.array_equal(e.cpu().numpy(), gd[tag + "_item_emg"])          # bit-exact copy
    assert np.array_equal(g.cpu().numpy(), gd[tag + "_item_glove"])
    assert np.array_equal(l.cpu().numpy(), gd[tag + "_item_label"])
    # batched fast path == stacking the per-item path == oracle
    items = torch.tensor([0, 5, ds.D - 1, 7])
    EMG, GLOVE, lab = tw.get_batch(items)
    EMG_use, tensor, D = OD.load_valid(emg.transpose(0, 1).contiguous().numpy(), db2, split)
    ref, ref_lab = OD.get_items(EMG_use, tensor, fixed_perm(41, D, 11), items.numpy(), train=(split == "train"))
    assert EMG.shape == ref.shape
    assert np.array_equal(EMG.cpu().numpy(), ref)
    assert np.array_equal(lab.cpu().numpy(), ref_lab)
    assert torch.equal(EMG[1], tw[5][0])


def test_fused_normalisation_bit_exact(gd):
    x = torch.from_numpy(gd["norm_x"]).cuda()
    rs = RunningStats(gd["norm_mean"], gd["norm_std"], device="cuda")
    assert np.array_equal(rs.normalize(x).cpu().numpy(), gd["norm_y"])     # reference RunningStats.normalize
    # scalar statistics (the shipped emg_mean.npy is a scalar)
    rs1 = RunningStats(np.float32(0.25), np.float32(1.75), device="cuda")
    ref = (gd["norm_x"] - np.float32(0.25)) / np.float32(1.75)
    assert np.array_equal(rs1.normalize(x).cpu().numpy(), ref)


def test_gather_with_stats_through_dataset(emg):
    mean = np.linspace(-0.3, 0.4, 12).astype(np.float32)
    std = np.linspace(0.5, 2.0, 12).astype(np.float32)
    ds = DB23(db2=False, device="cuda", emg_stats=RunningStats(mean, std, device="cuda"))
    ds.load_tensors(emg)
    for split in ("train", "test"):
        getattr(ds, "set_" + split)()
        rows = torch.from_numpy(fixed_perm(41, ds.D, 3)[:, :6].T.copy()).cuda()       # (6,41)
        out = ds[rows].cpu().numpy()
        src = (ds.EMG_use if split == "train" else ds.tensor).cpu().numpy()
        ref = OD.normalize(src[rows.cpu().numpy()], mean, std)
        assert np.array_equal(out.reshape(ref.shape), ref)


def test_edge_cases():
    src = torch.arange(7 * 5, dtype=torch.float32, device="cuda").reshape(7, 5)       # odd row length -> scalar path
    idx = torch.tensor([6, 0, 3, 3], device="cuda")
    out = gather_rows(src, idx)
    assert torch.equal(out, src[idx])
    empty = gather_rows(src, torch.zeros(0, dtype=torch.int64, device="cuda"))
    assert empty.shape == (0, 5)
    check_gather_errors("cuda:0")                                                     # nothing flagged so far
    gather_rows(src, torch.tensor([1, 99, -1], device="cuda"))
    with pytest.raises(IndexError):                                                   # out-of-range flagged (sticky)
        check_gather_errors("cuda:0")
    check_gather_errors("cuda:0")                                                     # ... and cleared by the check


def test_full_size_round_trip(emg):
    """DB2-shaped train split (820,000 rows): gathering with a permutation then with its inverse is
    the identity; the sum over a full per-class permutation equals the table sum."""
    ds = DB23(db2=True, device="cuda")
    ds.load_tensors(emg)
    ds.set_train()
    n = ds.EMG_use.shape[0]
    perm = torch.randperm(n, device="cuda")
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device="cuda")
    a = gather_rows(ds.EMG_use, perm)
    b = gather_rows(a, inv)
    assert torch.equal(b, ds.EMG_use)
