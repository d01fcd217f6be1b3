"""The A/B switches of the library (environment variables read once per process): every fast path of round 2 against the
path it replaced, on the same inputs, in separate processes.

    CP_FOLD_BN=0       BN-apply passes instead of folding the BatchNorm of linear blocks 1-3 into the next layer's weights
    CP_FUSE_BNBWD=0    un-fused BatchNorm backward instead of the data-gradient epilogue
    CP_TN_PAIR=0       single-CTA weight-gradient kernel instead of the CTA pair
    CP_CLIP_MMA=0      FFMA2 batch x batch sweeps instead of mma.sync
    CP_GLOVE_TC=0      fp32 FFMA glove tower instead of the tensor-core blocks
    CP_SUBSET_WARP=0   thread-per-trial subset evaluator instead of warp-per-trial (integer results: bit-exact)
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.environ["CP_ROOT"]); sys.path.insert(0, os.path.join(os.environ["CP_ROOT"], "tests"))
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.clip import ClipModel, clip_head
from contrastiveprosthetics_b200 import subset as cps
from gpu_util import load_sd, perturbed_state
out = {}
# encoder: forward + backward, dropout with injected masks, AdaBN
n = 1300
g = torch.Generator().manual_seed(1)
x = torch.randn(n, 12, generator=g) + 0.5 * torch.randn(41, 12, generator=g).repeat((n + 40) // 41, 1)[:n]
d_emb = torch.randn(n, 16, generator=g) / n
masks = torch.stack([torch.empty(n, 512, dtype=torch.uint8).bernoulli_(0.5, generator=g) for _ in range(4)])
m = Model({'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.0, 'reg_emg': 0.0, 'reg_glove': 0.0}, adabn=True, device="cuda")
load_sd(m, perturbed_state(3, True)); m.train(True)
m.emg_net.ext_dropout_masks = masks.cuda().contiguous()
emb = m.emg_net.encode_flat(x.cuda()); emb.backward(d_emb.cuda())
out["emb"] = emb.detach().cpu().numpy()
for k, p in m.emg_net.named_parameters():
    out["g|" + k] = p.grad.cpu().numpy()
# batch x batch head (above one tile) and the whole config-5 model
E = torch.randn(700, 16, generator=g); G = 0.6 * E + torch.randn(700, 16, generator=g)
Ed, Gd = E.cuda().requires_grad_(True), G.cuda().requires_grad_(True)
loss, ncor, arg = clip_head(Ed, Gd, 1.0); loss.backward()
out["clip_loss"] = np.array(loss.item()); out["clip_dE"] = Ed.grad.cpu().numpy(); out["clip_dG"] = Gd.grad.cpu().numpy()
out["clip_arg"] = arg.cpu().numpy()
torch.manual_seed(42)
cm = ClipModel({'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 1e-4, 'reg_glove': 1e-4}, glove_dim=22, device="cuda"); cm.train()
EMG = torch.randn(600, 1, 1, 12, generator=g); GL = torch.randn(600, 22, generator=g)
e, gl = cm(EMG.cuda(), GL.cuda()); l2 = cm.loss(e, gl) + cm.l2(); l2.backward()
out["cm_loss"] = np.array(l2.item())
for k, p in cm.glove_net.named_parameters():
    if p.grad is not None:
        out["cg|" + k] = p.grad.cpu().numpy()
# subset evaluator
lg = torch.randn(12, 25, 41, 41, generator=g).cuda()
msk, _ = cps.make_trials(sizes=range(1, 41, 3), trials_per_size=5)
c, t = cps.SubsetEvaluator(lg, 25).evaluate(msk)
out["sub_correct"] = c.cpu().numpy(); out["sub_total"] = t.cpu().numpy()
np.savez(os.environ["CP_OUT"], **out)
'''


def _run(tmp_path, tag, env_extra):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    out = tmp_path / f"{tag}.npz"
    env = dict(os.environ, CP_ROOT=ROOT, CP_OUT=str(out), **env_extra)
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.load(out)


def _rel(a, b):
    a, b = a.astype(np.float64).ravel(), b.astype(np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("switch", ["CP_FOLD_BN", "CP_FUSE_BNBWD", "CP_TN_PAIR", "CP_CLIP_MMA", "CP_GLOVE_TC", "CP_SUBSET_WARP"])
def test_fast_path_equals_the_path_it_replaced(tmp_path, switch):
    fast = _run(tmp_path, "fast", {})
    slow = _run(tmp_path, "slow", {switch: "0"})
    assert set(fast.files) == set(slow.files)
    for k in fast.files:
        if k in ("sub_correct", "sub_total", "clip_arg"):
            assert np.array_equal(fast[k], slow[k]), k                    # integer stages: bit-exact
        elif k.startswith("g|") or k.startswith("cg|"):
            # two fp32 evaluations of the same network may take different ReLU branches within rounding noise of 0
            # (DESIGN.md "ReLU kinks"): 1300-window batch -> a flip moves a tensor by ~1e-3 of its norm at most
            assert _rel(fast[k], slow[k]) < 5e-3, (k, _rel(fast[k], slow[k]))
        else:
            assert _rel(fast[k], slow[k]) < 1e-5, (k, _rel(fast[k], slow[k]))
