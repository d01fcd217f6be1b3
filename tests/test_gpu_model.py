"""End-to-end parity of the drop-in Model (gather -> encoder -> fused head -> loss/l2 -> backward,
eval with voting) against the fixtures produced by the unmodified reference and against the oracle."""
import os

import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.synthetic import fixed_perm, synth_emg
from contrastiveprosthetics_b200.utils import TaskWrapper
from oracle import model as OM
from gpu_util import load_sd, rel_err

pytestmark = pytest.mark.gpu
PARAMS = {'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 3e-4, 'reg_glove': 1e-5,
          'lr_emg': 1e-3, 'lr_glove': 2e-3, 'epochs': 1}


@pytest.fixture(scope="module")
def gm(golden_dir):
    return np.load(os.path.join(golden_dir, "model.npz"))


@pytest.fixture(scope="module")
def tw():
    ds = DB23(db2=False, device="cuda")
    ds.load_tensors(synth_emg())
    return TaskWrapper(ds, with_glove=False)


def _check_grads(gm, tag, grads, tol):
    n = 0
    for k in gm.files:
        parts = k.split("|")
        if parts[0] != tag or parts[1] not in ("gnorm", "gfull", "gval"):
            continue
        g = grads[parts[2]]
        if parts[1] == "gnorm":
            assert abs(float(g.double().norm()) - gm[k]) <= tol * gm[k], k
        elif parts[1] == "gfull":
            assert rel_err(g, gm[k]) < tol, k
        else:
            idx = gm[f"{tag}|gidx|{parts[2]}"]
            assert rel_err(g.reshape(-1)[idx], gm[k]) < tol, k
        n += 1
    assert n >= 40


@pytest.mark.parametrize("adabn", [True, False])
def test_first_train_step_matches_reference_fixture(gm, tw, adabn):
    tag = "adabn" if adabn else "stockbn"
    torch.manual_seed(42)
    model = Model(dict(PARAMS), adabn=adabn, device="cuda")           # same init as the reference (seed 42)
    tw.set_train()
    tw.emg_rand = torch.from_numpy(fixed_perm(41, tw.D, 11)).cuda()
    EMG, GLOVE, label = tw.get_batch(torch.from_numpy(gm[f"{tag}|train_items"][0]))
    assert np.array_equal(EMG.cpu().numpy(), gm[f"{tag}|EMG0"])
    model.set_train()
    model.materialize_logits = True
    model.emg_net.debug_tap = {}
    logits = model.forward(EMG, GLOVE, label.reshape(-1))
    loss = model.loss(logits, label.reshape(-1))
    l2 = model.l2()
    (loss + l2).backward()
    assert abs(loss.item() - gm[f"{tag}|train_losses"][0]) <= 1e-5 * gm[f"{tag}|train_losses"][0]
    assert abs(l2.item() - gm[f"{tag}|l2_0"]) <= 1e-5 * gm[f"{tag}|l2_0"]
    assert float((logits.detach().cpu() - torch.from_numpy(gm[f"{tag}|logits0"])).abs().max()) < 1e-5
    grads = {n: p.grad.detach().cpu() for n, p in model.named_parameters() if p.grad is not None}
    assert "logit_scale" not in grads
    # vs the stored reference gradients: bounded by ReLU-flip noise between two fp32 evaluations
    # (tests/test_gpu_encoder.py explains; 123-window batch -> ~1/sqrt(63k) per flip)
    _check_grads(gm, tag, grads, 3e-2)
    # vs the oracle (pinned to the reference at 2e-5 by tests/test_oracle_golden.py) with the
    # kernel's ReLU pattern injected: rounding-level agreement
    sd = OM.init_state(42, adabn)
    pat = [(model.emg_net.read_activation(s, 0).cpu() > 0) for s in range(9)]
    p = {k: (v.clone().requires_grad_(True) if k in OM.trainable_keys(sd) else v.clone()) for k, v in sd.items()}
    emb = OM.encoder_forward(p, EMG.cpu().reshape(-1, 12), adabn, True, relu_masks=pat)
    emb = emb.reshape(3, 41, 1, 16).transpose(1, 2).reshape(3, 41, 16)
    emb = emb / emb.norm(dim=-1, keepdim=True)
    tab = OM.class_table(p)
    tab = tab / tab.norm(dim=-1, keepdim=True)
    res = OM.contrastive_loss(torch.matmul(emb, tab.t()), True)
    (res["loss"] + OM.l2_penalty(p, PARAMS['reg_emg'], PARAMS['reg_glove'])).backward()
    for k in OM.trainable_keys(sd):
        assert rel_err(grads[k], p[k].grad) < 1e-5, k
    assert model.corrects[0] == gm[f"{tag}|train_corrects"][0]
    if not adabn:
        sd = model.state_dict()
        for k in gm.files:
            if k.startswith(f"{tag}|after1|"):
                name = k.split("|")[2]
                if sd[name].is_floating_point():
                    assert rel_err(sd[name], gm[k]) < 1e-5, name
                else:
                    assert int(sd[name]) == int(gm[k])


@pytest.mark.parametrize("adabn", [True, False])
def test_eval_with_vote_matches_reference_fixture(gm, tw, adabn):
    """state = init weights + running stats after one training forward (as the fixture was taken)."""
    tag = "adabn" if adabn else "stockbn"
    torch.manual_seed(42)
    model = Model(dict(PARAMS), adabn=adabn, device="cuda")
    model.set_train()
    with torch.no_grad():
        model.emg_net.encode_flat(torch.from_numpy(gm[f"{tag}|EMG0"]).cuda())
    tw.set_test()
    tw.emg_rand = torch.from_numpy(fixed_perm(41, tw.D, 13)).cuda()
    model.set_test()
    losses = []
    all_logits = []
    for bi, items in enumerate(gm[f"{tag}|eval_items"]):
        items = torch.tensor([int(i) for i in items if i >= 0])
        EMG, GLOVE, label = tw.get_batch(items)
        assert np.array_equal(EMG.cpu().numpy(), gm[f"{tag}|eval_EMG{bi}"])
        with torch.no_grad():
            logits = model.forward(EMG, GLOVE, label.reshape(-1))
            losses.append(model.loss(logits, label.reshape(-1)).item())
        assert float((logits.cpu() - torch.from_numpy(gm[f"{tag}|eval_logits{bi}"])).abs().max()) < 2e-5
        all_logits.append(logits.cpu())
    np.testing.assert_allclose(losses, gm[f"{tag}|eval_losses"], rtol=1e-5)
    # integer stages vs the oracle run on the kernel's own logits (bit-exact)...
    votes, ypred, cor = [], [], []
    for lg in all_logits:
        r = OM.contrastive_loss(lg, training=False, W=25)
        votes.append(r["voting_counts"]); ypred.append(r["y_pred"]); cor.append(OM.correct_float(r["correct_counts"]))
    assert np.array_equal(model.voting_raw(), np.concatenate(votes) / 41.0)
    assert np.array_equal(model.y_pred_raw(), np.concatenate(ypred))
    assert np.array_equal(model.y_true_raw(), np.tile(np.arange(41), (3, 1)))
    assert np.array_equal(model.correct_raw(), np.array(cor))
    # ...and vs the reference's stored outputs wherever no argmax is within float noise of a tie
    ref_votes = gm[f"{tag}|eval_voting"]
    assert np.abs(model.voting_raw() - ref_votes).max() <= 2 / 41.0
    assert (model.y_pred_raw() != gm[f"{tag}|eval_y_pred"]).mean() < 0.05


def test_c1_shaped_training_run(tw):
    """Config C1 shape (batch_size=8 groups): a short run must learn (loss falls, accuracy above chance)
    and stay finite; every step's loss equals the oracle's on the same weights and batch."""
    torch.manual_seed(0)
    model = Model(dict(PARAMS), adabn=True, device="cuda")
    opt_e = torch.optim.Adam(model.emg_net.parameters(), lr=1e-3)
    opt_g = torch.optim.Adam(model.glove_net.parameters(), lr=1e-3)
    tw.set_train()
    model.set_train()
    g = torch.Generator().manual_seed(1)
    losses = []
    for step, (EMG, GLOVE, label) in enumerate(tw.batches(8, shuffle=True, generator=g)):
        if step % 20 == 0:
            sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
            with torch.no_grad():
                ref = OM.contrastive_loss(OM.forward_logits(sd, EMG.cpu(), True, True), True)["loss"].item()
        logits = model.forward(EMG, GLOVE, label.reshape(-1))
        loss = model.loss(logits, label.reshape(-1))
        if step % 20 == 0:
            assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (step, loss.item(), ref)
        total = loss + model.l2()
        opt_e.zero_grad(set_to_none=True)
        opt_g.zero_grad(set_to_none=True)
        total.backward()
        opt_e.step()
        opt_g.step()
        losses.append(loss.item())
        if step == 100:
            break
    assert np.isfinite(losses).all()
    assert np.mean(losses[-10:]) < np.mean(losses[:10]) - 0.1
    assert model.correct() > 2.0 / 41


def test_l2_kernel_matches_torch_norms():
    """K5 (cp_l2_forward / cp_l2_backward) == sum of torch.norm(p) and its autograd gradient (models.py:344-349),
    including an all-zero tensor (torch's norm backward gives 0 there) and a scaled upstream gradient."""
    from contrastiveprosthetics_b200.models import _L2Fn
    g = torch.Generator().manual_seed(3)
    shapes = [(64, 1, 3, 3), (512, 768), (16, 512), (41,), (7, 5), (1,)]
    ps = [torch.randn(s, generator=g).cuda().requires_grad_() for s in shapes]
    ps[4].data.zero_()
    out = _L2Fn.apply(*ps)
    (out * 0.37).backward()
    qs = [p.detach().double().cpu().requires_grad_() for p in ps]
    ref = sum(torch.norm(q) for q in qs)
    (ref * 0.37).backward()
    assert abs(out.item() - ref.item()) <= 1e-6 * ref.item()
    for p, q in zip(ps, qs):
        assert torch.isfinite(p.grad).all()
        assert (p.grad.cpu().double() - q.grad).norm() <= 1e-6 * max(q.grad.norm().item(), 1e-30) + 1e-12
    assert torch.count_nonzero(ps[4].grad) == 0
    out2 = _L2Fn.apply(*ps)                                   # deterministic: bit-identical on a second call
    assert out2.item() == out.item()
