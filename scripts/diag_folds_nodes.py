"""Per-kernel durations inside ONE lockstep of folds.ConcurrentFolds at K = 1 and K = 16 (torch profiler / CUPTI):
which kernels stretch when the folds' branches run next to each other."""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrastiveprosthetics_b200.folds import ConcurrentFolds
from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.utils import TaskWrapper

PARAMS = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.0, 'reg_emg': 1e-5, 'reg_glove': 1e-5, 'lr_emg': 1e-3, 'lr_glove': 1e-3}
dev = torch.device("cuda")
ds = DB23(db2=True, device=dev)
ds.load_synthetic(with_glove=False)
tw = TaskWrapper(ds, with_glove=False)
tw.set_train()
EMG = tw.get_batch(torch.randperm(tw.D)[:8].to(dev))[0]
tables = {}
for K in (1, int(sys.argv[1]) if len(sys.argv) > 1 else 16):
    folds = ConcurrentFolds(tw, [dict(PARAMS) for _ in range(K)], 8)
    for _ in range(5):
        folds.step(EMG)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        folds.step(EMG)
        torch.cuda.synchronize()
    ks = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
    iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in ks)
    span = iv[-1][1] - iv[0][0]
    # busy time: union of the kernel intervals; concurrency = sum of durations / busy
    busy, cur_s, cur_e = 0.0, None, None
    for s, e, _ in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    tot = sum(e - s for s, e, _ in iv)
    print(f"K={K}: kernels {len(iv)} span {span:.1f} us, busy {busy:.1f} us, sum of durations {tot:.1f} us, mean concurrency {tot / busy:.2f}")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n in iv:
        agg[n[:60]][0] += 1
        agg[n[:60]][1] += e - s
    tables[K] = agg
    del folds
K = max(tables)
print(f"{'kernel':60s} n/fold  us@K=1  us@K={K}  stretch  total us/fold@K={K}")
for n, (c, t) in sorted(tables[K].items(), key=lambda kv: -kv[1][1]):
    c1, t1 = tables[1].get(n, [0, 0.0])
    m1 = t1 / c1 if c1 else float('nan')
    mk = t / c
    print(f"{n:60s} {c / K:5.1f} {m1:8.2f} {mk:8.2f} {mk / m1 if c1 else float('nan'):7.2f} {t / K:9.1f}")
