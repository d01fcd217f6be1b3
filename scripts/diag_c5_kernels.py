import sys, os, collections
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import profile, ProfilerActivity
from contrastiveprosthetics_b200.clip import ClipModel
from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.utils import TaskWrapper
PARAMS = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.5, 'reg_emg': 1e-5, 'reg_glove': 1e-5, 'lr_emg': 1e-3, 'lr_glove': 1e-3}
dev = torch.device("cuda"); B = 65536
torch.manual_seed(42)
model = ClipModel(dict(PARAMS), glove_dim=22, device="cuda"); model.train()
opt_e = torch.optim.Adam(model.emg_net.parameters(), lr=1e-3); opt_g = torch.optim.Adam(model.glove_net.parameters(), lr=1e-3)
ds = DB23(db2=True, device=dev); ds.load_synthetic(with_glove=True, glove_dim=22)
tw = TaskWrapper(ds, with_glove=True); tw.set_train()
def step(i):
    EMG, GLOVE, _ = tw.get_flat_batch(i * B, B)
    e, g = model(EMG, GLOVE)
    total = model.loss(e, g) + model.l2()
    opt_e.zero_grad(set_to_none=True); opt_g.zero_grad(set_to_none=True)
    total.backward(); opt_e.step(); opt_g.step()
for i in range(3): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(3); torch.cuda.synchronize()
ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ks:
    agg[e.name[:80]][0] += 1; agg[e.name[:80]][1] += e.time_range.end - e.time_range.start
tot = sum(v[1] for v in agg.values())
print("total kernel time ms", tot/1e3, "launches", len(ks))
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]:
    print(f"  {t/1e3:7.3f} ms n={c:3d} {n}")
