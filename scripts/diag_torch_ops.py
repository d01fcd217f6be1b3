"""Which torch ops (and kernels) run inside one C2 train step besides libcpros' own launches?  torch.profiler, 3 steps.
python scripts/diag_torch_ops.py [batch_groups]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from torch.profiler import profile, ProfilerActivity
from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.utils import TaskWrapper
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
PARAMS = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.5, 'reg_emg': 1e-5, 'reg_glove': 1e-5, 'lr_emg': 1e-3, 'lr_glove': 1e-3, 'epochs': 1}
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = Model(dict(PARAMS), adabn=True, device="cuda:0")
opt_e = torch.optim.Adam(model.emg_net.parameters(), lr=1e-3, fused=True)
opt_g = torch.optim.Adam(model.glove_net.parameters(), lr=1e-3, fused=True)
ds = DB23(db2=True, device=dev); ds.load_synthetic(with_glove=False)
tw = TaskWrapper(ds, with_glove=False); tw.set_train(); model.set_train()
items = torch.randperm(tw.D)[:B].to(dev)
def step():
    EMG, GLOVE, label = tw.get_batch(items)
    label = label.reshape(-1)
    logits = model.forward(EMG, GLOVE, label)
    loss = model.loss(logits, label) + model.l2()
    opt_e.zero_grad(set_to_none=True); opt_g.zero_grad(set_to_none=True)
    loss.backward()
    opt_e.step(); opt_g.step()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=70, max_name_column_width=70))
