"""Concurrent hyper-parameter folds on one GPU (reference: cross_validate, code/train.py:140-166; SURVEY.md
section 8f row 1).

The reference's workflow is 150 independent `train_loop(epochs=1)` runs at `--batch_size=8` (go.sh:6): 328 windows
per step, every kernel a few microseconds long, the GPU almost idle.  Across GPUs the folds are split per rank
(`train.cross_validate`); WITHIN a GPU `ConcurrentFolds` advances K folds in lockstep: each fold owns its model,
its two Adam optimizers, ONE CUDA graph of the whole step (`graph.GraphedTrainStep`) and ONE stream, so K graphs
are in flight at a time and the SMs one fold leaves idle run the others.  The folds draw the same batches (one
gather per step for all K; each fold would otherwise draw its own permutation of the same split, train.py:83-86) and
stay statistically independent through their own initialisation-free hyper-parameters and dropout streams.

Every kernel is deterministic and the folds share no state, so a fold trained concurrently ends bit-identical to
the same fold trained alone on the same batches (tests/test_gpu_folds.py).
"""
import torch
import torch.optim as optim

from .graph import GraphedTrainStep
from .models import Model


class ConcurrentFolds:
    def __init__(self, dataset, params_list, batch_size, adabn=True, seeds=None, dropout_seeds=None):
        """dataset: TaskWrapper in train mode; params_list: one hyper-parameter dict per fold (d_e, dp_emg, dp_glove,
        reg_emg, reg_glove, lr_emg, lr_glove); batch_size: groups per step (the captured shape, so the ragged last
        batch of an epoch is run eagerly)."""
        self.dataset, self.batch_size = dataset, batch_size
        self.device = dataset.device
        self.models, self.opts, self.steps, self.streams = [], [], [], []
        example = dataset.get_batch(torch.arange(batch_size))[0]
        self.label_full = torch.arange(example.shape[1], device=self.device).repeat(batch_size)
        for k, params in enumerate(params_list):
            torch.manual_seed(42 if seeds is None else seeds[k])             # models.py:12
            model = Model(params=dict(params), adabn=adabn, device=str(self.device)).to(torch.float32)
            model.emg_net.dropout_seed = 1000 + k if dropout_seeds is None else dropout_seeds[k]   # independent streams
            model.set_train()
            # train.py:72-73's two Adams in torch's single-kernel implementation: 2 graph nodes per fold instead of ~14
            opts = [optim.Adam(model.emg_net.parameters(), lr=params['lr_emg'], weight_decay=0, capturable=True, fused=True),
                    optim.Adam(model.glove_net.parameters(), lr=params['lr_glove'], weight_decay=0, capturable=True, fused=True)]
            self.models.append(model)
            self.opts.append(opts)
            self.steps.append(GraphedTrainStep(model, opts, example))
            self.streams.append(torch.cuda.Stream(device=self.device))
        self.losses = [[] for _ in params_list]
        self.accs = [[] for _ in params_list]

    def _eager_step(self, k, EMG):
        model, opts = self.models[k], self.opts[k]
        label = torch.arange(EMG.shape[1], device=self.device).repeat(EMG.shape[0])
        logits = model.forward(EMG, None, label)
        loss = model.loss(logits, label)
        total = loss + model.l2()
        for o in opts:
            o.zero_grad(set_to_none=True)
        total.backward()
        for o in opts:
            o.step()
        handle = logits if not torch.is_tensor(logits) else logits._cp_handle
        return loss.detach(), handle.ncor

    def step(self, EMG):
        """One training step of EVERY fold on the batch EMG (B, 41, 1, 1, 12)."""
        main = torch.cuda.current_stream(self.device)
        ready = main.record_event()
        for k, s in enumerate(self.streams):
            s.wait_event(ready)
            EMG.record_stream(s)
            with torch.cuda.stream(s):
                if EMG.shape[0] == self.batch_size:
                    loss, ncor = self.steps[k](EMG)
                    loss = loss.clone()                       # the graph's static outputs are overwritten by the next replay
                else:
                    loss, ncor = self._eager_step(k, EMG)
                self.losses[k].append(loss)
                # per-batch accuracy = mean over groups of (#correct / 41)  (models.py:166-172)
                self.accs[k].append(ncor.float().mean() / EMG.shape[1])

    def join(self):
        main = torch.cuda.current_stream(self.device)
        for s in self.streams:
            main.wait_stream(s)

    def run_epoch(self, shuffle=True, generator=None):
        """One pass over the training split for all folds.  Returns the mean training loss of every fold."""
        for m in self.models:
            m.set_train()
        self.losses = [[] for _ in self.models]
        self.accs = [[] for _ in self.models]
        for (EMG, _, _) in self.dataset.batches(self.batch_size, shuffle=shuffle, generator=generator):
            self.step(EMG)
        self.join()
        return [torch.stack(l).mean().item() for l in self.losses]

    def train_accuracy(self):
        """Model.correct() of the epoch (mean of the per-batch accuracies, models.py:210-211) for every fold."""
        return [torch.stack(a).mean().item() for a in self.accs]
