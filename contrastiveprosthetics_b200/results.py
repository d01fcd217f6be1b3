#!/usr/bin/env python
"""Reporting stage after the evaluator (reference: code/results.py:24-97).

`test()` keeps the reference's outputs and file names -- `logs.npy` (raw logits of every test group,
results.py:42-43), `y_pred.npy` / `y_true.npy` (voted decisions, flattened, 50-53), `voting.npy` (accuracy per
voting-window length, 56-57) -- and adds the file the reference meant to write at results.py:60 (it saves
`voting.npy` twice; the shipped artefact `data/confusion_matrix.npy` is the row-normalised matrix):
`confusion_matrix.npy`.  The confusion counts come from `cp_confusion_matrix` (integer kernel) instead of
sklearn; the class-subset accuracy tables of README.md:11-17 (`data/{mean,std,min,max}_grasp.xlsx`: one row per
subset size, statistics over 144 random trials) come from `SubsetEvaluator` on the same logits.

Differences from the reference are plumbing only: batches come from `TaskWrapper.batches` (one gather per
batch), `--synthetic` replaces the unavailable `emg.pt`, tables are written as `.npy` + `.csv` (openpyxl is not
a dependency), and under torchrun the subset trials are split per rank with the integer counts summed.
"""
import argparse
import os

import numpy as np
import torch

from . import _lib, dist as cpdist, subset as cps
from .constants import *  # noqa: F401,F403
from .load import DB23
from .models import Model
from .utils import TaskWrapper

shuff = True


def confusion_matrix(y_true, y_pred, n_classes=MAX_TASKS, device=None):  # noqa: F405
    """(C, C) int64 counts[t, p] on the GPU (sklearn.metrics.confusion_matrix with labels 0..C-1)."""
    dev = torch.device(device) if device is not None else (y_true.device if torch.is_tensor(y_true) else torch.device("cuda"))
    yt = torch.as_tensor(y_true, dtype=torch.int64).reshape(-1).to(dev).contiguous()
    yp = torch.as_tensor(y_pred, dtype=torch.int64).reshape(-1).to(dev).contiguous()
    if yt.numel() != yp.numel():
        raise ValueError("y_true and y_pred differ in length")
    counts = torch.empty((n_classes, n_classes), dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().cp_confusion_matrix(_lib.ptr(yt), _lib.ptr(yp), yt.numel(), n_classes, _lib.ptr(counts),
                                              _lib.ptr(err), _lib.stream()), "cp_confusion_matrix")
    if int(err.item()):
        raise ValueError(f"label outside [0, {n_classes})")
    return counts


def subset_tables(logits, window, sizes=range(1, MAX_TASKS), trials_per_size=144, seed=0):  # noqa: F405
    """Accuracy over random class subsets (README.md:11-17).  Returns {"sizes", "mean", "std", "min", "max"} with one
    entry per subset size -- the rows of data/{mean,std,min,max}_grasp.xlsx.  Trials are sharded over ranks."""
    masks, trial_sizes = cps.make_trials(sizes=sizes, trials_per_size=trials_per_size, seed=seed)
    lo, hi = cps.shard_trials(len(masks), cpdist.rank(), cpdist.world_size())
    ev = cps.SubsetEvaluator(logits, window)
    correct = torch.zeros(len(masks), dtype=torch.int64, device=logits.device)
    total = torch.zeros(len(masks), dtype=torch.int64, device=logits.device)
    if hi > lo:
        correct[lo:hi], total[lo:hi] = ev.evaluate(masks[lo:hi])
    correct, total = cpdist.sum_counts(correct, total)
    summary = cps.summarize(correct.cpu().numpy(), total.cpu().numpy(), trial_sizes)
    ks = sorted(summary)
    out = {"sizes": np.array(ks)}
    for stat in ("mean", "std", "min", "max"):
        out[stat] = np.array([summary[k][stat] for k in ks])
    return out


def test(model, dataset, save="../data/", batch_size=32, subsets=True, trials_per_size=144):
    """results.py:24-64 + the confusion matrix and subset tables.  Returns (mean_loss, acc)."""
    dataset.set_test()
    model.set_test()
    total_loss, logs = [], []
    for (EMG, GLOVE, label) in dataset.batches(batch_size, shuffle=shuff):
        label = label.reshape(-1)
        with torch.no_grad():
            logits = model.forward(EMG, GLOVE, label)
            total_loss.append(model.loss(logits, label))
            logs.append(logits)
    logs = torch.cat(logs)
    acc = model.correct()
    mean_loss = torch.stack(total_loss).cpu().numpy().mean()

    y_pred = model.y_pred_raw().flatten()
    y_true = model.y_true_raw().flatten()
    voting = model.voting_raw()
    counts = confusion_matrix(y_true, y_pred, device=logs.device).cpu().numpy()
    tables = subset_tables(logs, PREDICTION_WINDOW_SIZE, trials_per_size=trials_per_size) if subsets else None  # noqa: F405
    if save is not None and cpdist.rank() == 0:
        os.makedirs(save, exist_ok=True)
        np.save(save + "logs.npy", logs.cpu().numpy())
        np.save(save + "y_pred.npy", y_pred)
        np.save(save + "y_true.npy", y_true)
        np.save(save + "voting.npy", voting)
        # the shipped artefact is the row-normalised matrix (counts / test groups per class)
        np.save(save + "confusion_matrix.npy", counts / np.maximum(counts.sum(1, keepdims=True), 1))
        np.save(save + "confusion_counts.npy", counts)
        if tables is not None:
            for stat in ("mean", "std", "min", "max"):
                np.save(save + f"{stat}_grasp.npy", tables[stat])
                np.savetxt(save + f"{stat}_grasp.csv", tables[stat], header="0", comments="")
    print(counts, voting)
    return mean_loss, acc


def main(args):
    rank, world, device = cpdist.init_from_env()
    dataset23 = DB23(db2=args.db2, device=device)
    print("Loading dataset")
    if args.synthetic:
        dataset23.load_synthetic()
    else:
        dataset23.load_stored()
    print("Dataset loaded")
    dataset23 = TaskWrapper(dataset23)

    values = np.load(args.data_dir + "cross_val_values.npy")
    keys = np.load(args.data_dir + "cross_val_keys.npy")
    best_key = keys[np.nanargmax(values[:, 1])]
    d_e, lr_e, reg_e, dp_e, lr_g, reg_g, dp_g = best_key            # best model during validation
    scale = 1 / 10 if args.load_model else 1
    params = {'d_e': int(d_e), 'epochs': args.final_epochs, 'lr_emg': lr_e * scale, 'dp_emg': dp_e, 'reg_emg': reg_e,
              'lr_glove': lr_g * scale, 'dp_glove': dp_g, 'reg_glove': reg_g}
    model = Model(params=params, train_model=True, adabn=args.no_adabn, prediction=args.prediction, glove=args.glove,
                  device=str(device)).to(torch.float32)
    checkpoint = os.path.join(args.checkpoint_dir, "contrastive.pt")
    model.load_state_dict(torch.load(checkpoint, map_location=device))

    final_stats = test(model, dataset23, save=args.out_dir, batch_size=args.batch_size)
    print("loss,\t\t\tcorrect")
    print(final_stats)
    return final_stats


def build_parser():
    parser = argparse.ArgumentParser(description='Test-split report on ninapro dataset')
    # --- the reference's flags, verbatim (results.py:128-141)
    parser.add_argument('--crossval_size', type=int, default=10)
    parser.add_argument('--crossval_epochs', type=int, default=1)
    parser.add_argument('--batch_size', type=int, default=32)
    parser.add_argument('--final_epochs', type=int, default=10)
    parser.add_argument('--glove', action='store_true')
    parser.add_argument('--db2', action='store_true')
    parser.add_argument('--load_model', action='store_true')
    parser.add_argument('--crossval_load', action='store_true')
    parser.add_argument('--prediction', action='store_true')
    parser.add_argument('--no_adabn', action='store_false')
    parser.add_argument('--no_checkpoint', action='store_false')
    parser.add_argument('--no_verbose', action='store_false')
    parser.add_argument('--test', action='store_true')
    # --- additions (plumbing only)
    parser.add_argument('--synthetic', action='store_true', help='seeded NinaPro-shaped data instead of emg.pt')
    parser.add_argument('--data_dir', default="../data/")
    parser.add_argument('--checkpoint_dir', default="../checkpoints/")
    parser.add_argument('--out_dir', default="../data/")
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
