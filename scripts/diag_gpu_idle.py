"""GPU busy time vs span of one C2 training step (eager and CUDA-graph replay): how much is launch gap / tail?"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.utils import TaskWrapper
from contrastiveprosthetics_b200.graph import GraphedTrainStep

PARAMS = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.0, 'reg_emg': 1e-5, 'reg_glove': 1e-5, 'lr_emg': 1e-3, 'lr_glove': 1e-3}
dev = torch.device("cuda")
torch.manual_seed(42)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = Model(dict(PARAMS), adabn=True, device="cuda")
opts = [torch.optim.Adam(model.emg_net.parameters(), lr=1e-3, fused=True, capturable=True),
        torch.optim.Adam(model.glove_net.parameters(), lr=1e-3, fused=True, capturable=True)]
ds = DB23(db2=True, device=dev); ds.load_synthetic(with_glove=False)
tw = TaskWrapper(ds, with_glove=False); tw.set_train(); model.set_train()
EMG = tw.get_batch(torch.randperm(tw.D)[:B].to(dev))[0]
gs = GraphedTrainStep(model, opts, EMG)
for _ in range(3):
    gs(EMG)
torch.cuda.synchronize()

def report(tag, prof):
    ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ks)
    span = iv[-1][1] - iv[0][0]
    busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    gaps = []
    for s, e, n in iv[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, n))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print(f"{tag}: kernels {len(iv)}  span {span/1e3:.3f} ms  busy {busy/1e3:.3f} ms  idle {(span-busy)/1e3:.3f} ms  "
          f"sum of durations {sum(e-s for s,e,_ in iv)/1e3:.3f} ms")
    # context of the largest in-graph gap
    big = sorted(((iv[i][0] - max(e for _, e, _ in iv[:i]), i) for i in range(1, len(iv))), reverse=True)[:3]
    for gsz, i in big:
        print(f"   gap {gsz:.1f} us before #{i}:")
        for j in range(max(0, i - 4), min(len(iv), i + 3)):
            print(f"      #{j} start {iv[j][0]-iv[0][0]:9.1f} dur {iv[j][1]-iv[j][0]:7.1f}  {iv[j][2][:70]}")
    gaps.sort(reverse=True)
    print("   largest gaps (us, before kernel):", [(round(g, 1), n[:40]) for g, n in gaps[:8]])
    import statistics
    print("   median gap us:", statistics.median(g for g, _ in gaps), " gaps:", len(gaps))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n in iv:
        agg[n[:60]][0] += 1; agg[n[:60]][1] += e - s
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"   {t/1e3:7.3f} ms n={c:3d} {n}")

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gs(EMG)
    torch.cuda.synchronize()
report("graph replay", prof)
