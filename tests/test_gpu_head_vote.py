"""K3 / K4 / K4' parity through the C-ABI: fused head (loss, grads, argmax), vote, subset."""
import os

import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200 import _lib, subset as cps
from contrastiveprosthetics_b200.models import _HeadFn, _LogitsLossFn
from oracle import cvote
from oracle import model as OM
from oracle import vote_subset as OV
from gpu_util import rel_err

pytestmark = pytest.mark.gpu


def _oracle_head(emb, tw, tb, B, W, dtype=torch.float32):
    e = emb.to(dtype).clone().requires_grad_(True)
    w = tw.to(dtype).clone().requires_grad_(True)
    b = tb.to(dtype).clone().requires_grad_(True)
    x = e.reshape(B, 41, W, 16).transpose(1, 2).reshape(B * W, 41, 16)
    x = x / x.norm(dim=-1, keepdim=True)
    tab = w.t() + b[None, :]
    tab = tab / tab.norm(dim=-1, keepdim=True)
    logits = torch.matmul(x, tab.t())
    res = OM.contrastive_loss(logits, training=(W == 1), W=W, argmax_via_softmax=False)
    res["loss"].backward()
    return logits.detach(), res, e.grad, w.grad, b.grad


@pytest.mark.parametrize("B,W", [(1, 1), (8, 1), (300, 1), (3, 25), (1500, 1)])
def test_fused_head_matches_oracle(B, W):
    g = torch.Generator().manual_seed(B * 31 + W)
    emb = torch.randn(B * 41 * W, 16, generator=g)
    tw, tb = torch.randn(16, 41, generator=g), torch.randn(16, generator=g)
    logits, res, de, dw, db = _oracle_head(emb, tw, tb, B, W)
    e = emb.cuda().requires_grad_(True)
    w = tw.cuda().requires_grad_(True)
    b = tb.cuda().requires_grad_(True)
    loss, pred, ncor, lg = _HeadFn.apply(e, w, b, B, W, True, True)
    loss.backward()
    assert abs(loss.item() - res["loss"].item()) <= 1e-5 * abs(res["loss"].item())
    assert rel_err(lg, logits) < 2e-6
    assert rel_err(e.grad, de) < 1e-5
    assert rel_err(w.grad, dw) < 1e-5
    assert rel_err(b.grad, db) < 1e-5
    # integer outputs: bit-exact on the kernel's own logits (argmax of near-tied floats is not comparable
    # across summation orders)
    own = OM.contrastive_loss(lg.cpu(), training=(W == 1), W=W, argmax_via_softmax=True)
    assert np.array_equal(pred.cpu().numpy(), own["preds"])
    assert np.array_equal(ncor.cpu().numpy(), (own["preds"] == np.arange(41)[None]).sum(1))


def test_head_upstream_gradient_scaling():
    g = torch.Generator().manual_seed(3)
    emb = torch.randn(4 * 41, 16, generator=g).cuda().requires_grad_(True)
    w = torch.randn(16, 41, generator=g).cuda().requires_grad_(True)
    b = torch.randn(16, generator=g).cuda().requires_grad_(True)
    loss = _HeadFn.apply(emb, w, b, 4, 1, True, False)[0]
    (loss * 3.0).backward()
    g3 = emb.grad.clone()
    emb.grad = None
    _HeadFn.apply(emb, w, b, 4, 1, True, False)[0].backward()
    assert torch.allclose(g3, emb.grad * 3.0, rtol=1e-6, atol=0)


def test_loss_from_materialised_logits(golden_dir):
    gm = np.load(os.path.join(golden_dir, "model.npz"))
    ref_logits = torch.from_numpy(gm["adabn|logits0"])                       # the reference's own logits
    lg = ref_logits.cuda().requires_grad_(True)
    loss, pred, ncor = _LogitsLossFn.apply(lg, True)
    loss.backward()
    l0 = ref_logits.clone().requires_grad_(True)
    res = OM.contrastive_loss(l0, training=True)
    res["loss"].backward()
    assert abs(loss.item() - gm["adabn|train_losses"][0]) <= 1e-5 * gm["adabn|train_losses"][0]
    assert rel_err(lg.grad, l0.grad) < 1e-5
    assert np.array_equal(pred.cpu().numpy(), res["preds"])
    assert OM.correct_float(ncor.cpu().numpy()) == gm["adabn|train_corrects"][0]


@pytest.mark.parametrize("labels", [2, 5, 41])
def test_vote_bit_exact(labels):
    rs = np.random.RandomState(labels)
    preds = rs.randint(0, labels, size=(37, 25, 41)).astype(np.int32)
    L = _lib.lib()
    p = torch.from_numpy(preds).cuda()
    votes = torch.empty((37, 249), dtype=torch.int32, device="cuda")
    y_pred = torch.empty((37, 41), dtype=torch.int64, device="cuda")
    _lib.check(L.cp_vote_eval(_lib.ptr(p), 37, 25, 249, _lib.ptr(votes), _lib.ptr(y_pred), _lib.stream()))
    v0, y0 = OV.vote(preds)
    assert np.array_equal(votes.cpu().numpy(), v0)
    assert np.array_equal(y_pred.cpu().numpy(), y0)


def test_vote_on_reference_logits_reproduces_reference_outputs(golden_dir):
    """reference eval logits -> cp_logits_loss argmax -> cp_vote_eval == the reference's voting / y_pred."""
    gm = np.load(os.path.join(golden_dir, "model.npz"))
    L = _lib.lib()
    for tag in ("adabn", "stockbn"):
        votes_all, ypred_all = [], []
        for bi in range(2):
            lg = torch.from_numpy(gm[f"{tag}|eval_logits{bi}"]).cuda()
            loss, pred, ncor = _LogitsLossFn.apply(lg, False)
            B = lg.shape[0] // 25
            votes = torch.empty((B, 249), dtype=torch.int32, device="cuda")
            y_pred = torch.empty((B, 41), dtype=torch.int64, device="cuda")
            _lib.check(L.cp_vote_eval(_lib.ptr(pred), B, 25, 249, _lib.ptr(votes), _lib.ptr(y_pred), _lib.stream()))
            votes_all.append(votes.cpu().numpy())
            ypred_all.append(y_pred.cpu().numpy())
            assert abs(loss.item() - gm[f"{tag}|eval_losses"][bi]) <= 1e-5 * gm[f"{tag}|eval_losses"][bi]
        assert np.array_equal(np.concatenate(ypred_all), gm[f"{tag}|eval_y_pred"])
        assert np.array_equal(np.concatenate(votes_all) / 41.0, gm[f"{tag}|eval_voting"])


def test_subset_eval_bit_exact():
    rs = np.random.RandomState(4)
    logits = rs.randn(6, 25, 41, 41).astype(np.float32)
    logits[1, :, 5, :] = 0.5                   # a fully tied row: first label of the subset wins
    logits[2, 3, 7, 10] = logits[2, 3, 7, 20]  # a two-way tie
    masks, sizes = cps.make_trials(sizes=[1, 2, 3, 10, 25, 40], trials_per_size=7, seed=2)
    ev = cps.SubsetEvaluator(torch.from_numpy(logits).cuda(), 25)
    c, t = ev.evaluate(masks)
    c0, t0 = OV.subset_eval(logits, masks)
    assert np.array_equal(c.cpu().numpy(), c0) and np.array_equal(t.cpu().numpy(), t0)
    # ranked rows really are a descending stable sort
    order = ev.order.cpu().numpy().reshape(6, 25, 41, 41)
    ref_order = np.argsort(-logits, axis=-1, kind="stable")
    assert np.array_equal(order, ref_order)
    # edge cases: no trials; the empty subset
    c, t = ev.evaluate(np.zeros((0, 41), dtype=np.uint8))
    assert c.numel() == 0
    c, t = ev.evaluate(np.zeros((2, 41), dtype=np.uint8))
    assert c.tolist() == [0, 0] and t.tolist() == [0, 0]


def test_subset_full_set_equals_vote(golden_dir):
    gm = np.load(os.path.join(golden_dir, "model.npz"))
    logits = np.concatenate([gm["adabn|eval_logits0"], gm["adabn|eval_logits1"]])
    ev = cps.SubsetEvaluator(torch.from_numpy(logits).cuda(), 25)
    c, t = ev.evaluate(np.ones((1, 41), dtype=np.uint8))
    ref_counts = np.rint(gm["adabn|eval_voting"][:, -1] * 41).astype(np.int64)
    assert int(c[0]) == ref_counts.sum() and int(t[0]) == 41 * 3


def test_subset_full_size_against_c_oracle():
    """C4 shape: 160 test items x 25 x 41 x 41 logits, 144 trials (one subset size) -- bit-exact vs the
    plain-C oracle twin; plus a size-independent property: permuting trials permutes counts."""
    rs = np.random.RandomState(5)
    logits = rs.randn(160, 25, 41, 41).astype(np.float32)
    masks, _ = cps.make_trials(sizes=[10], trials_per_size=144, seed=3)
    ev = cps.SubsetEvaluator(torch.from_numpy(logits).cuda(), 25)
    c, t = ev.evaluate(masks)
    c0, t0 = cvote.subset_eval(logits, masks)
    assert np.array_equal(c.cpu().numpy(), c0) and np.array_equal(t.cpu().numpy(), t0)
    perm = rs.permutation(144)
    c2, _ = ev.evaluate(masks[perm])
    assert np.array_equal(c2.cpu().numpy(), c0[perm])
