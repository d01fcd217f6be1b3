"""GPU parity of the glove-angle tower (cp_glove_forward / cp_glove_backward) and of the whole config-5
model (EMG tower + glove tower + batch x batch CLIP loss) against oracle/clip.py ("parity unpinned":
the reference keeps this tower commented out, models.py:384-429)."""
import numpy as np
import pytest
import torch

from gpu_util import rel_err

pytestmark = pytest.mark.gpu


def _tower(glove_dim, dp=0.0, seed=7):
    from contrastiveprosthetics_b200.clip import GloveTower
    from oracle import clip as OC
    torch.manual_seed(seed)
    tower = GloveTower(glove_dim=glove_dim, dp=dp, device="cuda")
    sd = OC.glove_init_state(seed, glove_dim)
    # identical construction order -> identical initial values
    for k, v in tower.state_dict().items():
        assert torch.equal(v.cpu(), sd["glove_net." + k]), k
    # non-trivial BN affine
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for k, p in tower.named_parameters():
            if ".bn." in k:
                p.copy_((1.0 + 0.2 * torch.randn(p.shape, generator=g) if k.endswith("weight")
                         else 0.1 * torch.randn(p.shape, generator=g)).cuda())
    sd = {"glove_net." + k: v.detach().cpu().clone() for k, v in tower.state_dict().items()}
    return tower, sd


@pytest.mark.parametrize("glove_dim,n", [(20, 300), (22, 1000), (22, 37)])
def test_glove_tower_forward_backward(glove_dim, n):
    from oracle import clip as OC
    tower, sd = _tower(glove_dim)
    tower.train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, glove_dim, generator=g)
    d_emb = torch.randn(n, 16, generator=g)
    emb = tower(x.cuda())
    emb.backward(d_emb.cuda())
    sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    taps = []
    ref = OC.glove_forward(sdr, x.double(), taps=taps)
    ref.backward(d_emb.double())
    assert rel_err(emb, ref) < 1e-5
    # gradients: tight when no ReLU input of the fp64 evaluation is within fp32 noise of its kink (both sides
    # then take the same branches); otherwise only the kink-noise bound holds (DESIGN.md "ReLU kinks")
    margin = min(float(t.abs().min()) for t in taps)
    tol = 1e-5 if margin > 2e-5 else 2e-2
    for k, p in tower.named_parameters():
        assert rel_err(p.grad, sdr["glove_net." + k].grad) < tol, (k, margin)


def test_glove_tower_dropout_masks_and_eval():
    from oracle import clip as OC
    tower, sd = _tower(22, dp=0.5)
    n = 256
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, 22, generator=g)
    masks = (torch.rand(3, n, 256, generator=g) > 0.5).to(torch.uint8)
    tower.train()
    tower.ext_dropout_masks = masks.cuda()
    emb = tower(x.cuda())
    sdr = {k: v.double() for k, v in sd.items()}
    ref = OC.glove_forward(sdr, x.double(), dp=0.5, dropout_masks=[m.double() for m in masks])
    assert rel_err(emb, ref) < 1e-5
    # in-kernel Philox masks: keep rate ~ 0.5, deterministic per (seed, step)
    tower.ext_dropout_masks = None
    a = tower(x.cuda())
    tower._step -= 1
    b = tower(x.cuda())
    assert torch.equal(a, b)
    # eval: no dropout, batch statistics
    tower.eval()
    with torch.no_grad():
        ev = tower(x.cuda())
    assert rel_err(ev, OC.glove_forward(sdr, x.double())) < 1e-5


def test_clip_model_train_step_matches_oracle():
    """EMG tower + glove tower + CLIP loss + l2: loss and every gradient against the CPU oracle."""
    from contrastiveprosthetics_b200.clip import ClipModel
    from oracle import clip as OC, model as OM
    params = {'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 1e-4, 'reg_glove': 1e-4}
    torch.manual_seed(42)
    model = ClipModel(params, glove_dim=22, device="cuda")
    model.train()
    n = 512
    g = torch.Generator().manual_seed(9)
    EMG = torch.randn(n, 1, 1, 12, generator=g)
    GLOVE = torch.randn(n, 22, generator=g)
    e, gl = model(EMG.cuda(), GLOVE.cuda())
    loss = model.loss(e, gl)
    (loss + model.l2()).backward()

    sd = {k: v.detach().cpu().double().requires_grad_(v.dtype.is_floating_point and k != "logit_scale")
          for k, v in model.state_dict().items()}
    er = OM.encoder_forward(sd, EMG.double().reshape(-1, 12), adabn=True, training=True)
    gr = OC.glove_forward(sd, GLOVE.double())
    res = OC.clip_loss(er, gr, 0.0)
    reg = sum(torch.norm(v) for k, v in sd.items()
              if k != "logit_scale" and 'bn' not in k and 'bias' not in k) * 1e-4
    (res["loss"] + reg).backward()
    assert abs(loss.item() - res["loss"].item()) < 1e-5 * abs(res["loss"].item())
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        assert rel_err(p.grad, sd[k].grad) < 2e-2, k                      # un-conditioned ReLU kinks
    assert int(model.n_correct[-1].item()) == res["n_correct"]
