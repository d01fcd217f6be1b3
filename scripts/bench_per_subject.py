"""Per-subject AdaBN (models.py:245) at the C3 shape: 4096 groups x 41 windows, 46 subjects, forward + backward."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.utils import TaskWrapper

PARAMS = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.0, 'reg_emg': 1e-5, 'reg_glove': 1e-5, 'lr_emg': 1e-3, 'lr_glove': 1e-3}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ds = DB23(device="cuda", mixed=True); ds.load_synthetic(with_glove=False)
tw = TaskWrapper(ds, with_glove=False); tw.with_subjects = True; tw.set_train()
torch.manual_seed(42)
model = Model(dict(PARAMS), adabn=True, device="cuda"); model.set_train()
EMG, GLOVE, label = tw.get_batch(torch.randperm(tw.D)[:B])
label = label.reshape(-1)


def step():
    model.zero_grad(set_to_none=True)
    loss = model.loss(model.forward(EMG, GLOVE, label), label) + model.l2()
    loss.backward()


def timed(k=5, w=2):
    for _ in range(w): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


model.emg_net.per_subject = False
print(f"pooled statistics: {timed():.2f} ms fwd+bwd for {B * 41} windows")
model.emg_net.per_subject = True
for s in (1, 4, 16, 46):
    model.emg_net.segment_streams = s
    model.emg_net._seg_pool = None
    print(f"per-subject ({int(EMG._cp_subjects.unique().numel())} subjects), {s:2d} streams: {timed():.2f} ms")
