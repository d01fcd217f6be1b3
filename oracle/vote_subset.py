"""Oracle (TEST INFRASTRUCTURE): windowed majority vote and class-subset evaluator, numpy.

vote:   /root/reference/code/models.py:149-166 (argmax -> prefix mode -> equality count).
subset: NO runnable reference (README.md:11,15; results.py:42-61 dumps the raw logits the
        offline step consumed).  Restated spec: for a class subset S
        (always containing the rest label 40), for every group and every row i in S:
        pred_w = the label j in S with the largest logits[g,w,i,j] (first max in ascending
        label order), decision = mode of pred_w over the W=25 window (smallest label on
        ties), correct iff decision == i.  Output integer (correct, total) per trial.
        PARITY UNPINNED beyond: S = all 41 classes reproduces models.py's y_pred / counts.
"""
import numpy as np

from .model import prefix_mode, VOTE_LOOP


def vote(preds, n_votes=VOTE_LOOP - 1):
    """preds (B,W,T) ints -> votes (B,n_votes) #correct per prefix window, y_pred (B,T)."""
    B, W, T = preds.shape
    votes = np.empty((B, n_votes), dtype=np.int64)
    y_pred = np.empty((B, T), dtype=np.int64)
    tgt = np.arange(T)
    idx = np.minimum(np.arange(1, n_votes + 1), W) - 1
    for b in range(B):
        modes = prefix_mode(preds[b])
        votes[b] = (modes == tgt[None, :]).sum(1)[idx]
        y_pred[b] = modes[-1]
    return votes, y_pred


def subset_eval(logits, masks):
    """logits (B,W,T,T) float32; masks (n_trials,T) {0,1}.  Returns (correct, total) int64 (n_trials,)."""
    B, W, T, _ = logits.shape
    n = masks.shape[0]
    correct = np.zeros(n, dtype=np.int64)
    total = np.zeros(n, dtype=np.int64)
    for t in range(n):
        S = np.flatnonzero(masks[t])
        if len(S) == 0:
            continue
        sub = logits[:, :, S][:, :, :, S]                  # (B,W,|S|,|S|)
        pred = S[np.argmax(sub, axis=-1)]                   # first max in ascending label order
        for b in range(B):
            modes = prefix_mode(pred[b])[-1]                # (|S|,) full-window vote
            correct[t] += int((modes == S).sum())
        total[t] = B * len(S)
    return correct, total


def confusion_counts(y_true, y_pred, n_classes):
    """/root/reference/code/results.py:58 (sklearn.metrics.confusion_matrix with labels 0..C-1): counts[t, p].
    Pinned by tests/test_oracle_vote.py against sklearn and the reference's data/confusion_matrix.npy."""
    y_true = np.asarray(y_true).reshape(-1)
    y_pred = np.asarray(y_pred).reshape(-1)
    counts = np.zeros((n_classes, n_classes), dtype=np.int64)
    np.add.at(counts, (y_true, y_pred), 1)
    return counts
