"""Time the batch x batch (CLIP) head alone (config 5 shape): python scripts/bench_clip.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from contrastiveprosthetics_b200 import clip as C  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = torch.Generator().manual_seed(0)
E = torch.randn(B, 16, generator=g).cuda().requires_grad_(True)
G = torch.randn(B, 16, generator=g).cuda().requires_grad_(True)
ops = C._CudaOps
eh, _ = ops.normalize(E.detach())
gh, _ = ops.normalize(G.detach())
ght = ops.transpose(gh)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rowsum, _ = ops.sums(eh, ght, B, 1.0, True)
ms_sums = timeit(lambda: ops.sums(eh, ght, B, 1.0, True))
ms_grad = timeit(lambda: ops.grad(eh, ght, B, 1.0, rowsum, rowsum, 1.0 / B))


def full():
    E.grad = G.grad = None
    loss, _, _ = C.clip_head(E, G, 0.0)
    loss.backward()


ms_full = timeit(full)
fl_s = 2.0 * B * B * 16
print(f"B={B}: sums pass {ms_sums:.3f} ms ({fl_s / ms_sums / 1e9:.1f} TFLOP/s fp32, {B * B / ms_sums / 1e9:.2f} T exp/s); "
      f"grad pass {ms_grad:.3f} ms ({2 * fl_s / ms_grad / 1e9:.1f} TFLOP/s); "
      f"head fwd+bwd (2 sums + 2 grad passes) {ms_full:.3f} ms = {B / ms_full / 1e3:.2f} M pairs/s")
