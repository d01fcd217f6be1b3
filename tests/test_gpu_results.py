"""Reporting stage (code/results.py:24-97): confusion-matrix kernel bit-exact vs the oracle, the files
results.test() writes (names / shapes / known-answer relations of the reference's artefacts)
and the class-subset tables."""
import os

import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200 import results as cpres, train as cptrain
from oracle import vote_subset as OV

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,C", [(0, 41), (1, 41), (1968, 41), (1_000_003, 41), (5000, 64), (777, 1)])
def test_confusion_matrix_bit_exact(n, C):
    rs = np.random.RandomState(n % 97 + C)
    yt, yp = rs.randint(0, C, n), rs.randint(0, C, n)
    got = cpres.confusion_matrix(torch.from_numpy(yt).cuda(), torch.from_numpy(yp).cuda(), C).cpu().numpy()
    assert got.dtype == np.int64 and np.array_equal(got, OV.confusion_counts(yt, yp, C))
    assert got.sum() == n and np.array_equal(got.sum(1), np.bincount(yt, minlength=C))


def test_confusion_matrix_reference_artifact(golden_dir):
    d = os.path.join(golden_dir, "ref_data")
    y_true, y_pred = np.load(os.path.join(d, "y_true.npy")), np.load(os.path.join(d, "y_pred.npy"))
    got = cpres.confusion_matrix(y_true, y_pred, 41, device="cuda").cpu().numpy()
    assert np.array_equal(got / 48, np.load(os.path.join(d, "confusion_matrix.npy")))


def test_confusion_matrix_rejects_bad_labels():
    with pytest.raises(ValueError):
        cpres.confusion_matrix(torch.tensor([0, 41]).cuda(), torch.tensor([0, 0]).cuda(), 41)
    with pytest.raises(RuntimeError):
        cpres.confusion_matrix(torch.tensor([0]).cuda(), torch.tensor([0]).cuda(), 65)


def test_results_main_after_train_main(tmp_path):
    """train.py --test writes the checkpoint + cross_val tables; results.py then reloads them and dumps the report."""
    common = ["--synthetic", f"--data_dir={tmp_path}/data/", f"--checkpoint_dir={tmp_path}/ckpt/"]
    cptrain.main(cptrain.build_parser().parse_args(
        ["--final_epochs=1", "--crossval_size=2", "--crossval_epochs=1", "--batch_size=64", "--no_verbose"] + common))
    out = f"{tmp_path}/report/"
    loss, acc = cpres.main(cpres.build_parser().parse_args(["--batch_size=16", f"--out_dir={out}"] + common))
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0
    G = 48                                                   # DB3-shaped test split: 6 subjects x 2 reps x 4 windows
    logs, y_pred, y_true = (np.load(out + f) for f in ("logs.npy", "y_pred.npy", "y_true.npy"))
    voting, cm, counts = (np.load(out + f) for f in ("voting.npy", "confusion_matrix.npy", "confusion_counts.npy"))
    assert logs.shape == (G * 25, 41, 41) and logs.dtype == np.float32
    assert y_pred.shape == y_true.shape == (G * 41,) and y_pred.dtype == np.int64
    assert voting.shape == (G, 249) and cm.shape == (41, 41)
    # the known-answer relations of the reference's own artefacts
    assert np.array_equal(np.sort(y_true.reshape(G, 41), 1), np.tile(np.arange(41), (G, 1)))
    assert np.array_equal((y_pred == y_true).reshape(G, 41).mean(1), voting[:, -1])
    assert np.allclose(voting * 41, np.round(voting * 41), atol=1e-9)
    assert np.array_equal(counts, OV.confusion_counts(y_true, y_pred, 41)) and np.array_equal(cm, counts / G)
    assert abs(np.trace(counts) / (G * 41) - acc) < 1e-6
    # subset tables: one row per size 1..40; 40 grasps + rest = the full class set -> every trial identical
    mean, std, mn, mx = (np.load(out + f"{s}_grasp.npy") for s in ("mean", "std", "min", "max"))
    assert mean.shape == std.shape == mn.shape == mx.shape == (40,)
    assert mn[-1] == mx[-1] and abs(mean[-1] - mn[-1]) < 1e-12 and std[-1] < 1e-12
    assert abs(mean[-1] - np.trace(counts) / (G * 41)) < 1e-12
    assert np.all(mn <= mean + 1e-12) and np.all(mean <= mx + 1e-12)      # (a mean of equal values may round up an ulp)
    # the full-set subset decision equals the voted y_pred (restricted argmax == argmax, same vote)
    tables1 = cpres.subset_tables(torch.from_numpy(logs).cuda(), 25, sizes=[40], trials_per_size=2)
    assert abs(tables1["mean"][0] - mean[-1]) < 1e-12
