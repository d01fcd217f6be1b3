import sys, os, collections
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import profile, ProfilerActivity
from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.utils import TaskWrapper
from contrastiveprosthetics_b200.graph import GraphedTrainStep
PARAMS = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.0, 'reg_emg': 1e-5, 'reg_glove': 1e-5, 'lr_emg': 1e-3, 'lr_glove': 1e-3}
dev = torch.device("cuda")
torch.manual_seed(42)
model = Model(dict(PARAMS), adabn=True, device="cuda")
opts = [torch.optim.Adam(model.emg_net.parameters(), lr=1e-3, fused=True, capturable=True),
        torch.optim.Adam(model.glove_net.parameters(), lr=1e-3, fused=True, capturable=True)]
ds = DB23(db2=True, device=dev); ds.load_synthetic(with_glove=False)
tw = TaskWrapper(ds, with_glove=False); tw.set_train(); model.set_train()
EMG = tw.get_batch(torch.randperm(tw.D)[:8].to(dev))[0]
gs = GraphedTrainStep(model, opts, EMG)
for _ in range(3): gs(EMG)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gs(EMG); torch.cuda.synchronize()
ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ks)
print("nodes", len(iv), "span us", iv[-1][1]-iv[1][0], "sum dur", sum(e-s for s,e,_ in iv[1:]))
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in iv[1:]:
    agg[n[:70]][0] += 1; agg[n[:70]][1] += e - s
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  n={c:3d} {t:7.1f} us  {n}")
