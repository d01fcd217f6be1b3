#!/usr/bin/env python
"""Experiment driver with the reference's flags (code/train.py:253-265, verbatim, including the
inverted store_false flags --no_adabn / --no_checkpoint / --no_verbose) and loop structure
(train_loop 65-138, validate 46-63, test 27-44, cross_validate 140-166, main 168-222).

Differences from the reference are confined to plumbing:
  * batches come from TaskWrapper.batches (one gather launch per batch) instead of a DataLoader
    that calls __getitem__ once per item; `--item_loader` restores the DataLoader path;
  * `--synthetic` installs seeded NinaPro-shaped tensors (emg.pt is unavailable offline);
  * under torchrun (WORLD_SIZE > 1) batches are sample-sharded and gradients all-reduced once per
    step (dist.FlatGradAllReduce); cross-validation folds are split per rank.
Reference quirks are kept on purpose: both StepLRs wrap the glove optimizer
(train.py:79-80), the checkpoint condition is always true (train.py:122), test() uses
batch_size*8 (train.py:32).
"""
import argparse
import os

import numpy as np
import torch
import torch.optim as optim
import torch.utils.data as data

from . import dist as cpdist
from .constants import *  # noqa: F401,F403
from .load import DB23
from .models import Model
from .utils import TaskWrapper

shuff = True
args = None


def _loader(dataset, batch_size, with_span=False):
    if getattr(args, "item_loader", False):
        if cpdist.world_size() > 1:
            raise RuntimeError("--item_loader is the reference's single-process DataLoader path; drop it under torchrun")
        return data.DataLoader(dataset, batch_size=batch_size, shuffle=shuff)
    return dataset.batches(batch_size, shuffle=shuff, rank=cpdist.rank(), world_size=cpdist.world_size(),
                           with_span=with_span)


def _evaluate(model, dataset, batch_size):
    """validate / test body (train.py:27-63).  Under torchrun every global eval batch is sharded over the ranks and
    put back together before anything is reported (SURVEY.md section 8e "voting eval: shard groups; integer counts
    summed"): BatchNorm statistics are taken over the rows of EVERY rank (the AdaBN eval batch is the global batch,
    as on one GPU), the per-group vote / y_pred arrays of each batch are assembled across ranks (exact integers),
    and the batch loss is the row-weighted mean -- so every rank returns the numbers a single GPU would."""
    world = cpdist.world_size()
    if world == 1:
        total_loss = []
        for (EMG, GLOVE, label) in _loader(dataset, batch_size):
            label = label.reshape(-1)
            with torch.no_grad():
                logits = model.forward(EMG, GLOVE, label)
                total_loss.append(model.loss(logits, label))
        mean_loss = torch.stack(total_loss).mean().item()
        return mean_loss, model.correct()
    saved_sync = model.emg_net.sync_bn
    model.emg_net.sync_bn = True
    total_loss = []
    try:
        for (EMG, GLOVE, label, lo, n_global) in _loader(dataset, batch_size, with_span=True):
            label = label.reshape(-1)
            with torch.no_grad():
                logits = model.forward(EMG, GLOVE, label)
                loss = model.loss(logits, label)
            model.assemble_last(lo, n_global)
            part = loss.detach().double() * (EMG.shape[0] / n_global)
            torch.distributed.all_reduce(part)
            total_loss.append(part)
    finally:
        model.emg_net.sync_bn = saved_sync
    mean_loss = torch.stack(total_loss).mean().item()
    return mean_loss, model.correct()


def test(model, dataset):
    dataset.set_test()
    model.set_test()
    return _evaluate(model, dataset, args.batch_size * 8)


def validate(model, dataset):
    dataset.set_val()
    model.set_val()
    return _evaluate(model, dataset, args.batch_size)


def train_loop(dataset, params, checkpoint=False, checkpoint_dir="../checkpoints/model", annealing=False,
               load=None, verbose=False):
    model = Model(params=params, train_model=True, adabn=args.no_adabn, prediction=args.prediction,
                  glove=args.glove, device=str(dataset.device)).to(torch.float32)
    model.emg_net.sync_bn = getattr(args, "sync_bn", False)      # global-batch BatchNorm statistics under torchrun
    if getattr(args, "per_subject_adabn", False):                # models.py:245: BatchNorm statistics per subject
        model.emg_net.per_subject = True
        dataset.with_subjects = True
    if load is not None:
        print("Loading model")
        model.load_state_dict(torch.load(load + ".pt"))
    # sample-sharded replicas must be ONE model: rank 0's initial values (or checkpoint) everywhere
    cpdist.broadcast_module(model)

    fused = True if getattr(args, "fused_adam", False) else None     # None: torch's default implementation
    optimizer_emg = optim.Adam(model.emg_net.parameters(), lr=params['lr_emg'], weight_decay=0, fused=fused)
    optimizer_glove = optim.Adam(model.glove_net.parameters(), lr=params['lr_glove'], weight_decay=0, fused=fused)
    if annealing:
        scheduler_emg = optim.lr_scheduler.CosineAnnealingLR(optimizer_emg, T_max=args.final_epochs, eta_min=0)
        scheduler_glove = optim.lr_scheduler.CosineAnnealingLR(optimizer_glove, T_max=args.final_epochs, eta_min=0)
    else:
        # reference quirk: BOTH schedulers drive the glove optimizer
        scheduler_emg = optim.lr_scheduler.StepLR(optimizer_glove, step_size=5, gamma=.2)
        scheduler_glove = optim.lr_scheduler.StepLR(optimizer_glove, step_size=5, gamma=.2)
    sync_grads = cpdist.FlatGradAllReduce(list(model.emg_net.parameters()) + list(model.glove_net.parameters()))
    # --lean_step: the loop body below without autograd / torch.optim (step.LeanTrainStep: the same kernels, the
    # regulariser's gradient and both Adams in one launch, the gradient bucket all-reduced in place).  The torch
    # optimizers stay as the schedulers' handle on the learning rates.
    lean = None
    if getattr(args, "lean_step", False) and not args.prediction and not getattr(args, "per_subject_adabn", False):
        from . import step as cpstep
        lean = cpstep.from_optimizers(model, [optimizer_emg, optimizer_glove], sync_grads=True)
        optimizer_emg._opt_called = optimizer_glove._opt_called = True    # (the schedulers' "step() before optimizer.step()" check)

    dataset.set_train()
    model.set_train()
    val_losses = {}
    final_val_acc = None
    print("Training...")
    for e in range(params['epochs']):
        loss_train = []
        for (EMG, GLOVE, label) in _loader(dataset, args.batch_size):
            if lean is not None:
                loss_train.append(lean(EMG)[0])
                continue
            label = label.reshape(-1)
            logits = model.forward(EMG, GLOVE, label)
            loss = model.loss(logits, label)
            loss_train.append(loss.detach())
            loss = loss + model.l2()
            optimizer_emg.zero_grad(set_to_none=True)
            optimizer_glove.zero_grad(set_to_none=True)
            loss.backward()
            sync_grads()
            optimizer_emg.step()
            optimizer_glove.step()
        acc_train = model.correct()
        scheduler_emg.step()
        scheduler_glove.step()
        if lean is not None:
            lean.set_lr(optimizer_emg.param_groups[0]['lr'], optimizer_glove.param_groups[0]['lr'])
        loss_train = torch.stack(loss_train).mean().item()
        if cpdist.world_size() > 1:         # logging only: mean over the ranks' local shards
            t = torch.tensor([loss_train, acc_train], dtype=torch.float64, device=dataset.device)
            torch.distributed.all_reduce(t)
            loss_train, acc_train = (t / cpdist.world_size()).tolist()

        if verbose:
            loss_val, acc_val = validate(model, dataset)
            final_val_acc = (loss_val, acc_val)
            val_losses[e] = loss_val
            print("Epoch %d. Train loss: %.4f\tVal loss: %.4f\tVal acc: %.6f\tTrain acc: %.4f"
                  % (e, loss_train, loss_val, acc_val, acc_train))
        if checkpoint and verbose and loss_val <= max(val_losses.values()) and cpdist.rank() == 0:
            print("Checkpointing model...")
            os.makedirs(os.path.dirname(checkpoint_dir) or ".", exist_ok=True)
            torch.save(model.state_dict(), checkpoint_dir + ".pt")
        model.set_train()
        dataset.set_train()

    if not verbose:
        loss_val, acc_val = validate(model, dataset)
        print("Epoch %d. Train loss: %.4f\tVal loss: %.4f\tVal acc: %.6f\tTrain acc: %.4f"
              % (e, loss_train, loss_val, acc_val, acc_train))
        final_val_acc = (loss_val, acc_val)
        if checkpoint and cpdist.rank() == 0:
            os.makedirs(os.path.dirname(checkpoint_dir) or ".", exist_ok=True)
            torch.save(model.state_dict(), checkpoint_dir + ".pt")
    return final_val_acc, model


def cross_validate(des, hyperparams, dataset, id_, epochs=6, save=True, load=False, load_dir=None,
                   data_dir="../data/"):
    """Random hyper-parameter search (train.py:140-166).  Folds are independent: under torchrun fold i
    runs on rank i % world_size (each with a 1-GPU train_loop) and results are gathered."""
    if load:
        return np.load(data_dir + "cross_val_values%s.npy" % id_), np.load(data_dir + "cross_val_keys%s.npy" % id_)
    combos = [(d_e,) + tuple(h) for d_e in des for h in zip(*hyperparams.values())]
    names = list(hyperparams.keys())
    r, w = cpdist.rank(), cpdist.world_size()
    mine = {}
    saved = (cpdist.rank, cpdist.world_size)
    K = int(getattr(args, "concurrent_folds", 1) or 1)
    try:
        cpdist.rank, cpdist.world_size = (lambda: 0), (lambda: 1)       # folds train un-sharded
        my_folds = [i for i in range(len(combos)) if i % w == r]
        # the learning rates are constant for the first 5 epochs (StepLR(step_size=5), train.py:79-80), which is
        # what a captured graph needs; longer folds, checkpoints to resume from or the item loader run one by one
        if K > 1 and epochs <= 5 and load_dir is None and not getattr(args, "item_loader", False):
            from .folds import ConcurrentFolds
            for g0 in range(0, len(my_folds), K):
                group = my_folds[g0:g0 + K]
                plist = []
                for i in group:
                    params = {'d_e': combos[i][0], 'epochs': epochs}
                    params.update(dict(zip(names, combos[i][1:])))
                    print(params)
                    plist.append(params)
                dataset.set_train()
                folds = ConcurrentFolds(dataset, plist, args.batch_size, adabn=args.no_adabn)
                for _ in range(epochs):
                    loss_train = folds.run_epoch()
                for i, model, lt, acc_train in zip(group, folds.models, loss_train, folds.train_accuracy()):
                    loss_val, acc_val = validate(model, dataset)
                    print("Epoch %d. Train loss: %.4f\tVal loss: %.4f\tVal acc: %.6f\tTrain acc: %.4f"
                          % (epochs - 1, lt, loss_val, acc_val, acc_train))
                    mine[i] = (loss_val, acc_val)
                    dataset.set_train()
                del folds
            my_folds = []
        for i in my_folds:
            key = combos[i]
            params = {'d_e': key[0], 'epochs': epochs}
            params.update(dict(zip(names, key[1:])))
            print(params)
            (loss_t, acc_t), _ = train_loop(dataset, params, checkpoint=False, verbose=False, load=load_dir)
            mine[i] = (loss_t, acc_t)
    finally:
        cpdist.rank, cpdist.world_size = saved
    if w > 1:
        gathered = [None] * w
        torch.distributed.all_gather_object(gathered, mine)
        mine = {k: v for g in gathered for k, v in g.items()}
    values = np.array([mine[i] for i in range(len(combos))])
    keys = np.array(combos)
    if save and r == 0:
        os.makedirs(data_dir, exist_ok=True)
        np.save(data_dir + "cross_val_values%s.npy" % id_, values)
        np.save(data_dir + "cross_val_keys%s.npy" % id_, keys)
    return values, keys


def main(a):
    global args
    args = a
    rank, world, device = cpdist.init_from_env()
    np.random.seed(42)                       # train.py:22: the hyper-parameter draws depend on it
    # the reference seeds torch at import (models.py:12, utils.py:14, load.py:18): same initial weights / sampling
    # streams on every run -- and on every rank
    torch.manual_seed(42)
    if device.type == "cuda":
        torch.cuda.manual_seed(42)
    dataset23 = DB23(db2=args.db2, device=device, mixed=getattr(args, "mixed", False))
    print("Loading dataset")
    if args.synthetic:
        dataset23.load_synthetic()
    else:
        dataset23.load_stored()
    print("Dataset loaded")
    dataset23 = TaskWrapper(dataset23)

    n = args.crossval_size
    hyperparams = {
        'lr_emg': 10 ** np.random.uniform(low=-6, high=-1, size=(n,)),
        'reg_emg': 10 ** np.random.uniform(low=-9, high=-1, size=(n,)),
        'dp_emg': np.random.uniform(low=.4, high=.6, size=(n,)),
        'lr_glove': 10 ** np.random.uniform(low=-6, high=-1, size=(n,)),
        'reg_glove': 10 ** np.random.uniform(low=-9, high=-1, size=(n,)),
        'dp_glove': np.random.uniform(low=0, high=.9, size=(n,)),
    }
    values, keys = cross_validate([16], hyperparams, dataset23, id_="", epochs=args.crossval_epochs, save=True,
                                  load=args.crossval_load, data_dir=args.data_dir)
    best_key = keys[np.nanargmax(values[:, 1])]
    print("Best combination: %s" % str(best_key))
    d_e, lr_e, reg_e, dp_e, lr_g, reg_g, dp_g = best_key
    scale = 1 / 10 if args.load_model else 1
    params = {'d_e': int(d_e), 'epochs': args.final_epochs, 'lr_emg': lr_e * scale, 'dp_emg': dp_e,
              'reg_emg': reg_e, 'lr_glove': lr_g * scale, 'dp_glove': dp_g, 'reg_glove': reg_g}
    checkpoint_dir = os.path.join(args.checkpoint_dir, "contrastive")
    final_vals, model = train_loop(dataset23, params, checkpoint=args.no_checkpoint, annealing=True,
                                   checkpoint_dir=checkpoint_dir, verbose=args.no_verbose,
                                   load=checkpoint_dir if args.load_model else None)
    print("Final validation model statistics")
    print(final_vals)
    if args.no_checkpoint and os.path.exists(checkpoint_dir + ".pt"):
        model.load_state_dict(torch.load(checkpoint_dir + ".pt"))
    if args.test:
        final_stats = test(model, dataset23)
        print("loss,\t\t\tcorrect")
        print(final_stats)
        return final_stats
    return final_vals


def build_parser():
    parser = argparse.ArgumentParser(description='Training on ninapro dataset')
    # --- the reference's flags, verbatim (train.py:253-265)
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
    parser.add_argument('--item_loader', action='store_true', help="reference-style per-item DataLoader")
    parser.add_argument('--sync_bn', action='store_true',
                        help='under torchrun: BatchNorm statistics over the rows of every rank (global-batch parity) '
                             'instead of rank-local ones')
    parser.add_argument('--per_subject_adabn', action='store_true',
                        help='AdaBN statistics per subject (models.py:245 "batch per subject"): every BatchNorm '
                             'normalises the windows of one subject at a time, in training and evaluation')
    parser.add_argument('--lean_step', action='store_true',
                        help="train step without autograd / torch.optim (step.LeanTrainStep): same kernels and update rule, "
                             "~35 fewer launches per step; what bench.py's cuda_graph_lean mode and --concurrent_folds use")
    parser.add_argument('--fused_adam', action='store_true',
                        help="torch.optim.Adam(fused=True): same update rule in one kernel per optimizer (what bench.py uses)")
    parser.add_argument('--concurrent_folds', type=int, default=1,
                        help='cross-validation: train this many hyper-parameter folds at a time on one GPU, each as a '
                             'CUDA graph on its own stream (folds.ConcurrentFolds); 1 = one after the other')
    parser.add_argument('--mixed', action='store_true',
                        help='DB2 + DB3 subjects mixed (46 people, DB3 repetition split, DB3 channel 10 zeroed)')
    parser.add_argument('--data_dir', default="../data/")
    parser.add_argument('--checkpoint_dir', default="../checkpoints/")
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
