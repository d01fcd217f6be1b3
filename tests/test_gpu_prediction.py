"""--prediction mode on the GPU (cp_encoder_forward/backward with trunk_only + cp_cls_forward_backward, through
Model(prediction=True)) against the fixture of the UNMODIFIED reference (tests/golden/prediction.npz) and the oracle
(oracle/model.py prediction_step, pinned to that fixture by tests/test_oracle_prediction.py)."""
import os

import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200 import _lib
from contrastiveprosthetics_b200.models import Model
from oracle import model as OM
from gpu_util import load_sd, perturbed_state, rel_err
from test_oracle_prediction import check_grads, oracle_step

pytestmark = pytest.mark.gpu
PARAMS = {'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 3e-4, 'reg_glove': 1e-5,
          'lr_emg': 1e-3, 'lr_glove': 2e-3, 'epochs': 1}


@pytest.fixture(scope="module")
def gp(golden_dir):
    return np.load(os.path.join(golden_dir, "prediction.npz"))


@pytest.mark.parametrize("engine", [_lib.ENGINE_SIMT, _lib.ENGINE_TC])
@pytest.mark.parametrize("adabn", [True, False])
def test_prediction_step_matches_reference_fixture(gp, adabn, engine):
    tag = "adabn" if adabn else "stockbn"
    torch.manual_seed(42)
    model = Model(dict(PARAMS), adabn=adabn, prediction=True, device="cuda")       # seed-42 init == the reference's
    assert list(model.state_dict().keys()) == [str(k) for k in gp[f"{tag}|keys"]]
    model.emg_net.engine = engine
    model.set_train()
    model.emg_net.debug_tap = {}
    EMG = torch.from_numpy(gp[f"{tag}|EMG"]).cuda()
    label = torch.from_numpy(gp[f"{tag}|label"]).cuda()
    feats = model.forward(EMG, None, label)
    loss = model.loss(feats, label)
    l2 = model.l2()
    (loss + l2).backward()
    assert feats.shape == (164, 41)
    assert float((feats.detach().cpu() - torch.from_numpy(gp[f"{tag}|features"])).abs().max()) < 1e-5
    assert abs(loss.item() - gp[f"{tag}|loss"]) <= 1e-5 * gp[f"{tag}|loss"]
    assert abs(l2.item() - gp[f"{tag}|l2"]) <= 1e-5 * gp[f"{tag}|l2"]
    assert model.corrects[0] == float(gp[f"{tag}|correct"])
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    check_grads(gp, tag, grads, 3e-2)                 # vs the stored reference gradients: ReLU-flip noise (164 windows)
    # vs the oracle with the kernel's ReLU pattern: rounding-level agreement on every tensor
    pat = [(model.emg_net.read_activation(s, 0).cpu() > 0) for s in range(9)]
    pat.append(model.emg_net.debug_tap["relu_head"].cpu() > 0)          # ... and the head's own ReLU
    _, ofeats, oloss, _, ograds, _ = oracle_step(gp, tag, adabn, relu_masks=pat)
    assert rel_err(feats, ofeats) < 1e-5
    for k, g in ograds.items():
        assert rel_err(grads[k], g) < 1e-5, k
    if not adabn:
        sd = model.state_dict()
        for k in gp.files:
            if k.startswith(f"{tag}|after1|"):
                name = k.split("|")[2]
                if sd[name].is_floating_point():
                    assert rel_err(sd[name], gp[k]) < 1e-5, name
                else:
                    assert int(sd[name]) == int(gp[k]), name


@pytest.mark.parametrize("n,dp", [(41 * 7, 0.5), (1000, 0.0), (41 * 100, 0.5)])
def test_prediction_head_vs_oracle(n, dp):
    """Perturbed BatchNorm affine, dropout with injected masks, ragged sizes; forward values and head gradients."""
    adabn = True
    sd = OM.init_state(5, adabn, prediction=True)
    base = perturbed_state(5, adabn)
    for k in base:
        if k in sd and sd[k].shape == base[k].shape and ".last." not in k:
            sd[k] = base[k]
    g = torch.Generator().manual_seed(n)
    sd["emg_net.last.2.bn.weight"] = 1.0 + 0.2 * torch.randn(128, generator=g)
    sd["emg_net.last.2.bn.bias"] = 0.1 * torch.randn(128, generator=g)
    x = torch.randn(n, 12, generator=g)
    labels = torch.randint(0, 41, (n,), generator=g)
    masks = [torch.empty(n, 512).bernoulli_(0.5, generator=g) for _ in range(4)] if dp > 0 else None
    params = dict(PARAMS)
    params['dp_emg'] = dp
    m = Model(params, adabn=adabn, prediction=True, device="cuda")
    load_sd(m, sd)
    m.set_train()
    m.emg_net.debug_tap = {}
    if masks is not None:
        m.emg_net.ext_dropout_masks = torch.stack(masks).to(torch.uint8).cuda().contiguous()
    feats = m.forward(x.cuda(), None, labels.cuda())
    loss = m.loss(feats, labels.cuda())
    loss.backward()
    pat = [(m.emg_net.read_activation(s, 0).cpu() > 0) for s in range(9)]
    pat.append(m.emg_net.debug_tap["relu_head"].cpu() > 0)
    p = {k: (v.clone().requires_grad_(True) if k in OM.trainable_keys(sd) else v.clone()) for k, v in sd.items()}
    z = OM.encoder_forward(p, x, adabn, True, masks, dp, relu_masks=pat, prediction=True)
    of = z / z.norm(dim=-1, keepdim=True)
    ol = torch.nn.functional.cross_entropy(of, labels)
    ol.backward()
    assert rel_err(feats, of) < 1e-5 and abs(loss.item() - ol.item()) < 1e-5 * ol.item()
    for k in p:
        if getattr(p[k], "grad", None) is not None and k.startswith("emg_net."):
            got = dict(m.named_parameters())[k].grad
            assert rel_err(got, p[k].grad) < 3e-5, (k, rel_err(got, p[k].grad))
    assert m.corrects[0] == float((of.argmax(-1) == labels).double().mean())


def test_prediction_vote_evaluation_raises_like_the_reference():
    m = Model(dict(PARAMS), adabn=True, prediction=True, device="cuda")
    m.set_test()
    EMG = torch.randn(2, 41, 25, 1, 12, device="cuda")
    with pytest.raises(AssertionError, match="wrong logit shape for val time"), torch.no_grad():
        m.forward(EMG, None, torch.arange(41, device="cuda").repeat(2))
