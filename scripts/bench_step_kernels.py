"""K6 alone: cp_step_prologue and cp_adam_step at the model's size (40 tensors, 2,027,616 parameters), CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrastiveprosthetics_b200 import _lib
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.step import LeanTrainStep

PARAMS = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.0, 'reg_emg': 1e-5, 'reg_glove': 1e-5}
torch.manual_seed(42)
model = Model(dict(PARAMS), adabn=True, device="cuda")
model.set_train()
s = LeanTrainStep(model, 1e-3, 1e-3)
s.grad_flat.normal_(generator=torch.Generator(device="cuda").manual_seed(0))
L, P = _lib.lib(), _lib.ptr


def prologue():
    _lib.check(L.cp_step_prologue(s._reg_ptrs, s._reg_sizes, s._n_reg, P(s.norms), P(s.counters), 2, P(s._ws), s._ws.numel(),
                                  _lib.stream()))


def adam():
    _lib.check(L.cp_adam_step(s._p_ptrs, s._sizes, s._offs, len(s.params), P(s.grad_flat), P(s.exp_avg), P(s.exp_avg_sq),
                              P(s.lr), s._lr_index, s._reg, s._norm_index, P(s.norms), P(s.counters[1:2]), 0.9, 0.999, 1e-8,
                              _lib.stream()))


reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for name, fn, nbytes in (("step_prologue_kernel", prologue, 4 * sum(s._reg_sizes)),
                         ("adam_step_kernel", adam, 28 * sum(s._sizes))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name}: {us:.2f} us per launch (back to back, L2-resident: {nbytes / 1e6:.1f} MB of algorithmic traffic = "
          f"{nbytes / us / 1e3:.0f} GB/s)")
