"""CUDA-graph capture of the training step for launch-bound batch sizes.

The reference's own workflow (go.sh:6: --batch_size=8, 150 one-epoch cross-validation folds of 2,500
steps each, train.py:140-166) runs 328 windows per step: every kernel of the step is a few
microseconds long and the step is bound by launch latency (~600 launches: ours + Adam / l2 / index
kernels of torch), not by the GPU.  `GraphedTrainStep` captures

    forward -> fused head/loss (+ l2) -> backward -> Adam x2            (train.py:96-108)

once for a fixed batch size and replays it as ONE graph launch per step.  The batch gather stays an eager
launch in front of the graph (its source table is re-sliced every epoch, load.py:233-251, so its address is
not capturable); the gathered batch is copied into the graph's static input.

Dropout: the Philox key of the in-kernel mask generator reads a device-resident step counter
(cp_encoder_opts.dropout_step) that the graph itself increments, so every replay draws a fresh mask.
"""
import torch


def _opt_tensors(opt):
    return [v for st in opt.state.values() for v in st.values() if torch.is_tensor(v)]


class GraphedTrainStep:
    def __init__(self, model, optimizers, example_emg, sync_grads=None, warmup=3, capture=True, static_emg=None,
                 lean=False):
        """model: models.Model (training mode); optimizers: Adam(..., capturable=True) instances;
        example_emg: a (B,41,1,1,12) CUDA batch that fixes the captured shape.  The capture runs `warmup`
        + 1 real steps on it; parameters, BatchNorm buffers and optimizer state are restored afterwards.
        sync_grads: dist.FlatGradAllReduce for sample-sharded training (one process per GPU): the gradient
        all-reduce (and, with model.emg_net.sync_bn, the BatchNorm-statistics all-reduces) are captured INSIDE the
        graph -- NCCL collectives are capturable -- so every rank replays one graph per step and the ranks' host
        threads stop being a source of skew.  Every rank must construct and call the step in lockstep.
        capture=False: only the warm-up; the caller captures `body()` itself (folds.ConcurrentFolds puts the steps of
        K folds into ONE graph) and calls `restore()` afterwards.  static_emg: share the graph's input tensor.
        lean=True: the step runs WITHOUT autograd and torch.optim (step.LeanTrainStep: one prologue launch, the fused
        kernels, one Adam launch for both optimizers, gradients in one flat bucket that is all-reduced in place) --
        ~35 fewer graph nodes per step.  `optimizers` then only supply lr / betas / eps (they are not stepped; their
        state stays empty) and `sync_grads` only says WHETHER to average the gradients over the ranks."""
        self.lean = None
        if lean:
            from . import step as cpstep
            self.lean = cpstep.from_optimizers(model, optimizers, sync_grads=sync_grads is not None,
                                               group=getattr(model.emg_net, "process_group", None))
        self.sync_grads = sync_grads
        if sync_grads is not None:
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_backend() != "nccl":
                raise RuntimeError("capturing the gradient all-reduce in a CUDA graph needs the NCCL backend")
        for o in optimizers:
            if not lean and not all(g.get("capturable", False) for g in o.param_groups):
                raise RuntimeError("GraphedTrainStep needs optim.Adam(..., capturable=True)")
        self.model, self.optimizers = model, list(optimizers)
        dev = example_emg.device
        self.B = example_emg.shape[0]
        self.static_emg = example_emg.clone() if static_emg is None else static_emg
        self.static_label = torch.arange(example_emg.shape[1], device=dev).repeat(self.B)
        if self.lean is not None:
            self._step_dev = self.lean.counters[0:1]          # advanced by cp_step_prologue
        else:
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        model.emg_net.dropout_step = self._step_dev

        self._saved_model = {k: v.clone() for k, v in model.state_dict().items()}
        # optimizer state is created lazily by the first step: snapshot it if it exists, else it is reset to zero
        self._saved_opt = [[t.clone() for t in _opt_tensors(o)] if len(o.state) else None for o in self.optimizers]
        self._saved_lean = [t.clone() for t in self.lean.state_tensors()] if self.lean is not None else None
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(stream)
        self.graph = None
        self.steps = 0
        if capture:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss, self.n_correct = self._body()
            self.restore()

    def close(self):
        """Drop the captured graph.  With sync_grads the graph holds NCCL collectives: release it BEFORE
        torch.distributed.destroy_process_group(), which otherwise waits on the communicator forever."""
        self.graph = None

    def body(self):
        """One step on `static_emg` (what a capture records).  Returns (loss, per-group correct counts)."""
        return self._body()

    def restore(self):
        """Undo the warm-up / capture steps IN PLACE (a graph holds the addresses)."""
        with torch.no_grad():
            sd = self.model.state_dict()
            for k, v in self._saved_model.items():
                sd[k].copy_(v)
            for o, saved in zip(self.optimizers, self._saved_opt):
                for i, t in enumerate(_opt_tensors(o)):
                    if saved is None:
                        t.zero_()
                    else:
                        t.copy_(saved[i])
            self._step_dev.zero_()
            if self.lean is not None:
                for t, saved in zip(self.lean.state_tensors(), self._saved_lean):
                    t.copy_(saved)
        self.model.reset()
        self._saved_model = self._saved_opt = self._saved_lean = None
        self.steps = 0

    def _body(self):
        m = self.model
        if self.lean is not None:
            return self.lean.body(self.static_emg)
        self._step_dev.add_(1)
        logits = m.forward(self.static_emg, None, self.static_label)
        loss = m.loss(logits, self.static_label)
        total = loss + m.l2()
        for o in self.optimizers:
            o.zero_grad(set_to_none=True)
        total.backward()
        if self.sync_grads is not None:
            self.sync_grads()
        for o in self.optimizers:
            o.step()
        handle = logits if not torch.is_tensor(logits) else logits._cp_handle
        return loss.detach(), handle.ncor

    def __call__(self, EMG):
        """One training step on a (B,41,1,1,12) batch.  Returns the graph's static (loss, per-group correct
        counts) tensors: valid until the next call -- copy what must be kept."""
        if EMG.shape != self.static_emg.shape:
            raise RuntimeError(f"graph was captured for {tuple(self.static_emg.shape)}, got {tuple(EMG.shape)}")
        self.static_emg.copy_(EMG)
        self.graph.replay()
        self.steps += 1
        return self.loss, self.n_correct
