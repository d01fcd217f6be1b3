"""Which host call launches the small torch kernels of one training step?  (torch.profiler, with stacks)"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.utils import TaskWrapper

PARAMS = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.0, 'reg_emg': 1e-5, 'reg_glove': 1e-5, 'lr_emg': 1e-3, 'lr_glove': 1e-3}
dev = torch.device("cuda")
torch.manual_seed(42)
model = Model(dict(PARAMS), adabn=True, device="cuda")
opt_e = torch.optim.Adam(model.emg_net.parameters(), lr=1e-3, fused=True)
opt_g = torch.optim.Adam(model.glove_net.parameters(), lr=1e-3, fused=True)
ds = DB23(db2=True, device=dev); ds.load_synthetic(with_glove=False)
tw = TaskWrapper(ds, with_glove=False); tw.set_train(); model.set_train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8

def step():
    EMG, GLOVE, label = tw.get_batch(torch.randperm(tw.D)[:B].to(dev))
    label = label.reshape(-1)
    logits = model.forward(EMG, GLOVE, label)
    loss = model.loss(logits, label) + model.l2()
    opt_e.zero_grad(set_to_none=True); opt_g.zero_grad(set_to_none=True)
    loss.backward()
    opt_e.step(); opt_g.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
# CPU ops that launched kernels: count by op name + first repo frame
cnt = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CPU and ev.name.startswith("aten::") and ev.cpu_parent is not None:
        pass
ka = prof.key_averages(group_by_stack_n=6)
rows = []
for e in ka:
    if e.device_time_total > 0 and e.key.startswith("aten::"):
        stack = [s for s in e.stack if "contrastiveprosthetics_b200" in s or "diag_launch" in s or "torch/optim" in s or "autograd" in s][:3]
        rows.append((e.count, e.key, e.device_time_total, " <- ".join(s.split("/")[-1] for s in stack)))
rows.sort(reverse=True)
for r in rows[:60]:
    print(f"{r[0]:4d} {r[1]:32s} {r[2]:8.1f}us  {r[3]}")
