"""Config C1 (the reference's own run: --batch_size=8 -> 328 windows/step): eager vs CUDA-graph step."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def c1_small_batch(dev, B=8, steps=200):
    from contrastiveprosthetics_b200.graph import GraphedTrainStep
    from contrastiveprosthetics_b200.load import DB23
    from contrastiveprosthetics_b200.models import Model
    from contrastiveprosthetics_b200.utils import TaskWrapper
    params = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.5, 'reg_emg': 1e-5, 'reg_glove': 1e-5}
    ds = DB23(db2=True, device=dev)
    ds.load_synthetic(with_glove=False)
    tw = TaskWrapper(ds, with_glove=False)
    tw.set_train()
    out = {}
    for mode in ("eager", "graph", "graph_lean"):
        torch.manual_seed(42)
        model = Model(dict(params), adabn=True, device=str(dev))
        model.set_train()
        gm = mode != "eager"            # graph modes: torch's single-kernel Adam (fused=True), 2 nodes instead of ~14
        opts = [torch.optim.Adam(model.emg_net.parameters(), lr=1e-3, capturable=gm, fused=gm or None),
                torch.optim.Adam(model.glove_net.parameters(), lr=1e-3, capturable=gm, fused=gm or None)]
        items = [torch.randperm(tw.D)[:B].to(dev) for _ in range(16)]
        if gm:                          # graph_lean: step.LeanTrainStep inside the graph (no autograd / torch.optim nodes)
            step = GraphedTrainStep(model, opts, tw.get_batch(items[0])[0], lean=mode == "graph_lean")

            def one(i):
                step(tw.get_batch(items[i % 16])[0])
        else:
            def one(i):
                EMG, GLOVE, label = tw.get_batch(items[i % 16])
                label = label.reshape(-1)
                lg = model.forward(EMG, GLOVE, label)
                total = model.loss(lg, label) + model.l2()
                for o in opts:
                    o.zero_grad(set_to_none=True)
                total.backward()
                for o in opts:
                    o.step()
        for i in range(20):
            one(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[mode] = {"ms_per_step": ms, "windows_per_s": B * 41 / (ms / 1e3),
                     "wall_ms_per_step": 1e3 * (time.perf_counter() - t0) / steps}
    out["workload"] = f"C1: batch_size {B} groups x 41 = {B * 41} windows/step (go.sh:6), AdaBN, dropout 0.5, {steps} steps"
    return out


if __name__ == "__main__":
    import json
    print(json.dumps(c1_small_batch(torch.device("cuda:0"))))


def c1_concurrent_folds(dev, B=8, steps=200, ks=(1, 4, 8, 16, 32)):
    """The reference's cross-validation workload (150 folds x 2,500 steps at batch_size 8, train.py:140-166): K folds
    in lockstep on one GPU, ONE CUDA graph with K branches (folds.ConcurrentFolds, lean steps).  Aggregate windows/s."""
    from contrastiveprosthetics_b200.folds import ConcurrentFolds
    from contrastiveprosthetics_b200.load import DB23
    from contrastiveprosthetics_b200.utils import TaskWrapper
    ds = DB23(db2=True, device=dev)
    ds.load_synthetic(with_glove=False)
    tw = TaskWrapper(ds, with_glove=False)
    tw.set_train()
    params = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.5, 'reg_emg': 1e-5, 'reg_glove': 1e-5, 'lr_emg': 1e-3, 'lr_glove': 1e-3}
    items = [torch.randperm(tw.D)[:B].to(dev) for _ in range(16)]
    out = {}
    for K in ks:
        folds = ConcurrentFolds(tw, [dict(params) for _ in range(K)], B)

        def one(i):
            folds.step(tw.get_batch(items[i % 16])[0])
            if i % 8 == 7:
                folds.losses = [[] for _ in range(K)]
                folds.accs = [[] for _ in range(K)]
        for i in range(20):
            one(i)
        folds.join()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            one(i)
        folds.join()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0) / steps
        out[f"K={K}"] = {"ms_per_lockstep": ms, "fold_steps_per_s": K / (ms / 1e3), "windows_per_s": K * B * 41 / (ms / 1e3)}
        del folds
    out["workload"] = f"batch_size {B} groups x 41 = {B * 41} windows per fold-step, AdaBN, dropout 0.5, {steps} lock-steps, wall clock"
    return out
