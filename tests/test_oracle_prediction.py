"""--prediction mode: the oracle's restatement (oracle/model.py prediction_step; models.py:113-119, 175-196, 300-309)
pinned against the fixture the UNMODIFIED reference produced (tests/golden/make_golden.py prediction).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import model as OM

REG_EMG = 3e-4          # make_golden.PARAMS


@pytest.fixture(scope="module")
def gp(golden_dir):
    return np.load(os.path.join(golden_dir, "prediction.npz"))


def oracle_step(gp, tag, adabn, dtype=torch.float32, relu_masks=None):
    sd = OM.init_state(42, adabn, prediction=True)
    p = {k: (v.to(dtype).clone().requires_grad_(True) if k in OM.trainable_keys(sd) else
             (v.to(dtype).clone() if v.is_floating_point() else v.clone())) for k, v in sd.items()}
    new_stats = {}
    EMG = torch.from_numpy(gp[f"{tag}|EMG"]).to(dtype)
    label = torch.from_numpy(gp[f"{tag}|label"])
    z = OM.encoder_forward(p, EMG.reshape(-1, 12), adabn, True, new_stats=new_stats, relu_masks=relu_masks, prediction=True)
    feats = z / z.norm(dim=-1, keepdim=True)
    loss = torch.nn.functional.cross_entropy(feats, label)
    l2 = sum(torch.norm(v) for k, v in p.items() if k.startswith("emg_net.") and k in OM.trainable_keys(sd)
             and "bn" not in k and "bias" not in k) * REG_EMG            # Model.l2, prediction branch (models.py:226)
    (loss + l2).backward()
    grads = {k: v.grad for k, v in p.items() if getattr(v, "grad", None) is not None}
    return sd, feats.detach(), loss.detach(), l2.detach(), grads, new_stats


def check_grads(gp, tag, grads, tol):
    n = 0
    for k in gp.files:
        parts = k.split("|")
        if parts[0] != tag or parts[1] not in ("gnorm", "gfull", "gval"):
            continue
        g = grads[parts[2]].detach().cpu().double()
        if parts[1] == "gnorm":
            assert abs(float(g.norm()) - gp[k]) <= tol * gp[k], k
        elif parts[1] == "gfull":
            assert np.linalg.norm(g.numpy() - gp[k]) <= tol * np.linalg.norm(gp[k]), k
        else:
            ref = gp[k]
            got = g.reshape(-1).numpy()[gp[f"{tag}|gidx|{parts[2]}"]]
            assert np.linalg.norm(got - ref) <= tol * np.linalg.norm(ref), k
        n += 1
    assert n >= 20
    # parameters the reference leaves without a gradient (the glove tower, logit_scale) get none here either
    none = {k.split("|")[2] for k in gp.files if k.startswith(tag + "|gnone|")}
    assert none and not (none & {k for k, g in grads.items() if g is not None and float(g.abs().sum()) > 0})


@pytest.mark.parametrize("adabn", [True, False])
def test_prediction_init_and_keys(gp, adabn):
    tag = "adabn" if adabn else "stockbn"
    sd = OM.init_state(42, adabn, prediction=True)
    assert sorted(sd.keys()) == sorted(str(k) for k in gp[f"{tag}|keys"])
    for k in sd:
        v = sd[k].to(torch.float64).reshape(-1)
        dig = np.array([v.sum().item(), v.abs().sum().item()] + v[:4].tolist())
        np.testing.assert_allclose(dig, gp[f"{tag}|init|{k}"], rtol=0, atol=0)


@pytest.mark.parametrize("adabn", [True, False])
def test_prediction_step_matches_reference(gp, adabn):
    tag = "adabn" if adabn else "stockbn"
    sd, feats, loss, l2, grads, new_stats = oracle_step(gp, tag, adabn)
    assert float((feats - torch.from_numpy(gp[f"{tag}|features"])).abs().max()) < 2e-6
    assert abs(loss.item() - gp[f"{tag}|loss"]) <= 2e-6 * gp[f"{tag}|loss"]
    assert abs(l2.item() - gp[f"{tag}|l2"]) <= 2e-6 * gp[f"{tag}|l2"]
    check_grads(gp, tag, grads, 2e-4)
    label = torch.from_numpy(gp[f"{tag}|label"])
    assert float((feats.argmax(-1) == label).double().mean()) == float(gp[f"{tag}|correct"])
    if not adabn:
        for k in gp.files:
            if k.startswith(f"{tag}|after1|"):
                name = k.split("|")[2]
                if name.startswith("glove_net."):
                    continue                         # never run without --glove: stays at its initial value
                np.testing.assert_allclose(new_stats[name].numpy(), gp[k], rtol=1e-5, atol=1e-7)
    assert str(gp[f"{tag}|eval_error"]) == "wrong logit shape for val time"
