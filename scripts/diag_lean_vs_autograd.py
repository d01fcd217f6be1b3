"""Per-step losses and per-tensor parameter differences: lean step vs autograd step (torch Adam fused / foreach)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_step as T
from contrastiveprosthetics_b200.step import LeanTrainStep

batches = T._batches(6)
runs = {}
for name, fused in (("autograd_fused", True), ("autograd_foreach", False)):
    m = T._model()
    snaps = []
    opts = [torch.optim.Adam(m.emg_net.parameters(), lr=T.PARAMS['lr_emg'], fused=fused),
            torch.optim.Adam(m.glove_net.parameters(), lr=T.PARAMS['lr_glove'], fused=fused)]
    label = torch.arange(41, device="cuda").repeat(8)
    losses = []
    for EMG in batches:
        lg = m.forward(EMG, None, label)
        loss = m.loss(lg, label)
        total = loss + m.l2()
        for o in opts:
            o.zero_grad(set_to_none=True)
        total.backward()
        if len(snaps) == 0:
            grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        for o in opts:
            o.step()
        losses.append(loss.item())
        snaps.append({k: v.clone() for k, v in m.state_dict().items()})
    runs[name] = (losses, snaps, grads)
m = T._model()
lean = LeanTrainStep(m, T.PARAMS['lr_emg'], T.PARAMS['lr_glove'])
losses, snaps = [], []
for EMG in batches:
    losses.append(lean(EMG)[0].item())
    snaps.append({k: v.clone() for k, v in m.state_dict().items()})
runs["lean"] = (losses, snaps, None)
for k, (l, _, _) in runs.items():
    print(f"{k:18s}", " ".join(f"{x:.7f}" for x in l))
for a, b in (("lean", "autograd_fused"), ("autograd_foreach", "autograd_fused")):
    for step in (0, 1, 5):
        worst = sorted(((float((runs[a][1][step][k] - runs[b][1][step][k]).abs().max()), k) for k in runs[a][1][step]
                        if runs[a][1][step][k].dtype.is_floating_point), reverse=True)[:4]
        print(a, "vs", b, "after step", step + 1, ["%s %.3g" % (k, v) for v, k in worst])
