"""Concurrent hyper-parameter folds on one GPU (reference: cross_validate, code/train.py:140-166; SURVEY.md
section 8f row 1).

The reference's workflow is 150 independent `train_loop(epochs=1)` runs at `--batch_size=8` (go.sh:6): 328 windows
per step, every kernel a few microseconds long, the GPU almost idle.  Across GPUs the folds are split per rank
(`train.cross_validate`); WITHIN a GPU `ConcurrentFolds` advances K folds in lockstep: each fold owns its model and
its two Adam optimizers, and the whole steps of ALL K folds are captured into ONE CUDA graph with K parallel
branches (one capture stream per fold): a lockstep is one graph launch -- the host pays for one launch instead of
K x ~150 kernel nodes -- and the SMs one fold leaves idle run the others.  The folds draw the same batches (one
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
    def __init__(self, dataset, params_list, batch_size, adabn=True, seeds=None, dropout_seeds=None, lean=True):
        """dataset: TaskWrapper in train mode; params_list: one hyper-parameter dict per fold (d_e, dp_emg, dp_glove,
        reg_emg, reg_glove, lr_emg, lr_glove); batch_size: groups per step (the captured shape, so the ragged last
        batch of an epoch is run eagerly).  lean (default): every fold steps through step.LeanTrainStep -- no
        autograd / torch.optim nodes in the graph (~110 nodes per fold-step instead of ~145); lean=False keeps the
        autograd + torch.optim.Adam(fused=True) step."""
        self.dataset, self.batch_size, self.lean = dataset, batch_size, lean
        self.device = dataset.device
        self.models, self.opts, self.steps, self.streams = [], [], [], []
        example = dataset.get_batch(torch.arange(batch_size))[0]
        self.static_emg = example.clone()                     # the one input tensor of the K-fold graph
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
            self.steps.append(GraphedTrainStep(model, opts, example, capture=False, static_emg=self.static_emg,
                                               lean=lean))
            self.streams.append(torch.cuda.Stream(device=self.device))
        # ONE graph: fork a branch per fold from the capture stream, join them, stack the folds' results
        T = example.shape[1]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            cs = torch.cuda.current_stream(self.device)
            outs = []
            for step, s in zip(self.steps, self.streams):
                s.wait_stream(cs)
                with torch.cuda.stream(s):
                    outs.append(step.body())
            for s in self.streams:
                cs.wait_stream(s)
            self.loss_all = torch.stack([o[0] for o in outs])                         # (K,)
            # per-batch accuracy = mean over groups of (#correct / 41)  (models.py:166-172)
            self.acc_all = torch.stack([o[1] for o in outs]).float().mean(1) / T        # (K,)
        for step in self.steps:
            step.restore()
        self._hist = []                 # per lockstep: (loss (K,), acc (K,)) clones

    def _eager_step(self, k, EMG):
        model, opts = self.models[k], self.opts[k]
        if self.lean:
            return self.steps[k].lean.body(EMG)
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
        if EMG.shape[0] == self.batch_size:
            self.static_emg.copy_(EMG)
            self.graph.replay()                               # the steps of all K folds: one launch
            # the graph's static outputs are overwritten by the next replay
            self._hist.append((self.loss_all.clone(), self.acc_all.clone()))
            return
        # ragged last batch of an epoch: eagerly, fold by fold on the folds' streams
        main = torch.cuda.current_stream(self.device)
        ready = main.record_event()
        losses, accs = [], []
        for k, s in enumerate(self.streams):
            s.wait_event(ready)
            EMG.record_stream(s)
            with torch.cuda.stream(s):
                loss, ncor = self._eager_step(k, EMG)
                losses.append(loss)
                accs.append(ncor.float().mean() / EMG.shape[1])
        self.join()
        self._hist.append((torch.stack(losses), torch.stack(accs)))

    def join(self):
        main = torch.cuda.current_stream(self.device)
        for s in self.streams:
            main.wait_stream(s)

    # per-fold views of the lockstep history (fold k: list of 0-d tensors, one per step since the last reset)
    @property
    def losses(self):
        return [[h[0][k] for h in self._hist] for k in range(len(self.models))]

    @losses.setter
    def losses(self, _value):
        self._hist = []

    @property
    def accs(self):
        return [[h[1][k] for h in self._hist] for k in range(len(self.models))]

    @accs.setter
    def accs(self, _value):
        pass

    def run_epoch(self, shuffle=True, generator=None):
        """One pass over the training split for all folds.  Returns the mean training loss of every fold."""
        for m in self.models:
            m.set_train()
        self._hist = []
        for (EMG, _, _) in self.dataset.batches(self.batch_size, shuffle=shuffle, generator=generator):
            self.step(EMG)
        self.join()
        return torch.stack([h[0] for h in self._hist]).mean(0).tolist()

    def train_accuracy(self):
        """Model.correct() of the epoch (mean of the per-batch accuracies, models.py:210-211) for every fold."""
        return torch.stack([h[1] for h in self._hist]).mean(0).tolist()
