"""Vote / subset oracle: reference artefact known-answer relations, numpy vs
plain-C twin, and the only pin the subset evaluator has (full set == models.py vote).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import model as OM
from oracle import vote_subset as OV
from oracle import cvote


@pytest.fixture(scope="module")
def ref(golden_dir):
    d = os.path.join(golden_dir, "ref_data")
    return {k: np.load(os.path.join(d, k + ".npy")) for k in ("y_true", "y_pred", "voting", "confusion_matrix")}


def test_reference_artifact_relations(ref):
    y_true, y_pred, voting, cm = ref["y_true"], ref["y_pred"], ref["voting"], ref["confusion_matrix"]
    assert np.array_equal(y_true, np.tile(np.arange(41), 48))           # y_true = arange(41) per group
    acc = (y_pred == y_true).reshape(48, 41).mean(1)
    assert np.array_equal(acc, voting[:, -1])                           # last vote column == final acc
    assert np.allclose(voting * 41, np.round(voting * 41), atol=1e-9)   # every entry is count/41
    # confusion matrix == counts / 48 rows per class
    cm2 = np.zeros((41, 41))
    np.add.at(cm2, (y_true, y_pred), 1)
    assert np.array_equal(cm2 / 48, cm)
    # the oracle's accumulator reproduces the stored per-group accuracies from the stored y_pred
    counts = (y_pred == y_true).reshape(48, 41).sum(1)
    assert np.array_equal(counts / 41.0, voting[:, -1])
    assert abs(OM.correct_float(counts) - 0.33943089) < 1e-6


def test_prefix_mode_tie_rule_matches_torch():
    rs = np.random.RandomState(0)
    pred = rs.randint(0, 6, size=(25, 41))             # few labels -> many ties
    modes = OM.prefix_mode(pred)
    for w in range(1, 26):
        t = torch.from_numpy(pred[:w]).mode(0)[0].numpy()
        assert np.array_equal(modes[w - 1], t)


@pytest.mark.parametrize("labels", [3, 41])
def test_c_twin_matches_numpy_vote(labels):
    rs = np.random.RandomState(1)
    preds = rs.randint(0, labels, size=(7, 25, 41))
    v1, y1 = OV.vote(preds)
    v2, y2 = cvote.vote(preds)
    assert np.array_equal(v1, v2) and np.array_equal(y1, y2)


def _masks(rs, n, k_max=40):
    m = np.zeros((n, 41), dtype=np.uint8)
    for t in range(n):
        k = rs.randint(1, k_max + 1)
        m[t, rs.choice(40, size=k, replace=False)] = 1
        m[t, 40] = 1                                    # rest is always in the subset
    return m


def test_c_twin_matches_numpy_subset():
    rs = np.random.RandomState(2)
    logits = rs.randn(5, 25, 41, 41).astype(np.float32)
    logits[0, 0, 3, :] = 0.25                           # exact ties -> first max in label order
    masks = _masks(rs, 24)
    c1, t1 = OV.subset_eval(logits, masks)
    c2, t2 = cvote.subset_eval(logits, masks)
    assert np.array_equal(c1, c2) and np.array_equal(t1, t2)
    assert np.array_equal(t1, 5 * masks.sum(1))


def test_full_subset_reproduces_reference_vote(golden_dir):
    """The subset evaluator's only pin: with S = all 41 classes it must equal the reference's own
    y_pred / final vote counts, computed from the reference's own eval logits."""
    gm = np.load(os.path.join(golden_dir, "model.npz"))
    for tag in ("adabn", "stockbn"):
        logits = np.concatenate([gm[f"{tag}|eval_logits0"], gm[f"{tag}|eval_logits1"]])
        logits = logits.reshape(-1, 25, 41, 41)
        full = np.ones((1, 41), dtype=np.uint8)
        c, t = OV.subset_eval(logits, full)
        ref_counts = np.rint(gm[f"{tag}|eval_voting"][:, -1] * 41).astype(np.int64)
        assert c[0] == ref_counts.sum() and t[0] == 41 * logits.shape[0]
        c2, _ = cvote.subset_eval(logits, full)
        assert c2[0] == c[0]


def test_grasp_table_structure(golden_dir):
    """data/*_grasp.xlsx: size-40 subsets are the full class set (min == max), mean decreases."""
    t = np.load(os.path.join(golden_dir, "ref_data", "grasp_tables.npz"))
    assert t["mean_grasp"].shape == (40,)
    assert abs(t["min_grasp"][-1] - t["max_grasp"][-1]) < 1e-4
    assert t["mean_grasp"][0] > t["mean_grasp"][9] > t["mean_grasp"][39]


def test_confusion_counts_match_sklearn_and_reference_artifact(ref):
    """results.py:58: the oracle's confusion counts == sklearn's, and reproduce data/confusion_matrix.npy."""
    import sklearn.metrics as me
    y_true, y_pred = ref["y_true"], ref["y_pred"]
    counts = OV.confusion_counts(y_true, y_pred, 41)
    assert np.array_equal(counts, me.confusion_matrix(y_true, y_pred, labels=np.arange(41)))
    assert np.array_equal(counts / 48, ref["confusion_matrix"])
    rs = np.random.RandomState(0)
    a, b = rs.randint(0, 7, 1000), rs.randint(0, 7, 1000)
    assert np.array_equal(OV.confusion_counts(a, b, 7), me.confusion_matrix(a, b, labels=np.arange(7)))
    assert OV.confusion_counts(np.zeros(0, int), np.zeros(0, int), 5).sum() == 0
