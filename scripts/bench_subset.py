import sys, os, time, torch, numpy as np
sys.path.insert(0, "/root/repo")
from contrastiveprosthetics_b200 import subset as cps
items, W = 160, 25
g = torch.Generator().manual_seed(0)
lg = torch.randn(items, W, 41, 41, generator=g).cuda()
masks, sizes = cps.make_trials()
ev = cps.SubsetEvaluator(lg, W)
m = torch.from_numpy(masks).cuda()
for _ in range(3): ev.evaluate(m)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): c, t = ev.evaluate(m)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"subset eval {len(masks)} trials x {items*W*41} windows: {ms:.3f} ms = {items*W*41*len(masks)/ms/1e9*1e3/1e3:.3f} T preds/s  checksum {int(c.sum())}")
