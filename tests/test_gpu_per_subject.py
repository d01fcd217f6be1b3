"""Per-subject AdaBN (models.py:245 "momentum = 0 and batch per subject in order to have adaptive normalization";
SURVEY 8f row 3): BatchNorm statistics keyed by a subject-id vector.  No runnable reference exists (the comment
describes it, nothing implements it) -> parity unpinned beyond the op-level restatement in oracle/model.py."""
import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200.load import DB23
from contrastiveprosthetics_b200.models import Model
from contrastiveprosthetics_b200.utils import TaskWrapper
from oracle import model as OM
from gpu_util import load_sd, perturbed_state, rel_err

pytestmark = pytest.mark.gpu
PARAMS = {'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 0.0, 'reg_glove': 0.0}


def _model(sd, dp=0.0, streams=4):
    m = Model(dict(PARAMS, dp_emg=dp), adabn=True, device="cuda")
    load_sd(m, sd)
    m.train(True)
    m.emg_net.per_subject = True
    m.emg_net.segment_streams = streams
    return m


def _case(n, n_subj, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 12, generator=g)
    subj = torch.randint(0, n_subj, (n,), generator=g) * 3 + 5          # non-contiguous ids, unsorted rows
    x = x + 0.5 * subj[:, None].float() / n_subj                         # subjects differ in offset ...
    x = x * (1.0 + subj[:, None].float() / (3 * n_subj))                 # ... and in gain
    d_emb = torch.randn(n, 16, generator=g)
    return x, subj, d_emb


@pytest.mark.parametrize("streams", [1, 4])
def test_forward_matches_oracle_and_differs_from_pooled_statistics(streams):
    sd = perturbed_state(3, True)
    x, subj, _ = _case(1500, 5, 0)
    m = _model(sd, streams=streams)
    with torch.no_grad():
        emb = m.emg_net.encode_flat(x.cuda(), subjects=subj.cuda()).cpu()
        m.emg_net.per_subject = False
        pooled = m.emg_net.encode_flat(x.cuda()).cpu()
    ref = OM.encoder_forward(sd, x, True, True, subjects=subj)
    assert rel_err(emb, ref) < 1e-5
    assert rel_err(pooled, ref) > 1e-2                                   # the mode really changes the statistics


def test_equals_one_encoder_pass_per_subject_bit_for_bit():
    """Forward rows and the SUM of the weight gradients of stand-alone passes over each subject's rows (each of which
    the encoder tests pin to the oracle at 1e-5 / 2e-5) -- the segment loop adds nothing but the row permutation."""
    sd = perturbed_state(4, True)
    x, subj, d_emb = _case(1100, 4, 1)
    m = _model(sd)
    emb = m.emg_net.encode_flat(x.cuda(), subjects=subj.cuda())
    emb.backward(d_emb.cuda())
    got = {k: p.grad.clone() for k, p in m.emg_net.named_parameters()}
    total = None
    for s in torch.unique(subj).tolist():
        rows = torch.nonzero(subj == s).reshape(-1)
        m2 = _model(sd)
        m2.emg_net.per_subject = False
        e = m2.emg_net.encode_flat(x[rows].cuda())
        assert torch.equal(e.detach(), emb.detach()[rows.cuda()])
        e.backward(d_emb[rows].cuda())
        g = {k: p.grad.double() for k, p in m2.emg_net.named_parameters()}
        total = g if total is None else {k: total[k] + g[k] for k in g}
    for k in got:
        assert rel_err(got[k], total[k]) < 2e-6, k                       # fp32 summation order over the segments only


def test_gradients_against_oracle():
    """Un-conditioned bound (ReLU flips between two fp32 evaluations, tests/test_gpu_encoder.py) and the loss path."""
    sd = perturbed_state(5, True)
    x, subj, d_emb = _case(2000, 3, 2)
    m = _model(sd)
    emb = m.emg_net.encode_flat(x.cuda(), subjects=subj.cuda())
    emb.backward(d_emb.cuda())
    p = {k: (v.clone().requires_grad_(True) if k in OM.trainable_keys(sd) else v.clone()) for k, v in sd.items()}
    OM.encoder_forward(p, x, True, True, subjects=subj).backward(d_emb)
    for k, q in m.emg_net.named_parameters():
        assert rel_err(q.grad, p["emg_net." + k].grad) < 2e-2, k


def test_dropout_masks_follow_the_rows():
    sd = perturbed_state(6, True)
    x, subj, _ = _case(900, 3, 3)
    g = torch.Generator().manual_seed(9)
    masks = [(torch.rand(900, 512, generator=g) > 0.5) for _ in range(4)]
    m = _model(sd, dp=0.5)
    m.emg_net.ext_dropout_masks = torch.stack(masks).to(torch.uint8).cuda().contiguous()
    with torch.no_grad():
        emb = m.emg_net.encode_flat(x.cuda(), subjects=subj.cuda()).cpu()
    ref = OM.encoder_forward(sd, x, True, True, dropout_masks=masks, dp=0.5, subjects=subj)
    assert rel_err(emb, ref) < 1e-5
    # generated masks: segments draw from different Philox streams
    m.emg_net.ext_dropout_masks = None
    with torch.no_grad():
        a = m.emg_net.encode_flat(x.cuda(), subjects=subj.cuda())
        b = m.emg_net.encode_flat(x.cuda(), subjects=subj.cuda())
    assert torch.isfinite(a).all() and not torch.equal(a, b)


def test_single_window_subject_raises():
    sd = perturbed_state(7, True)
    x, subj, _ = _case(64, 2, 4)
    subj[0] = 999
    m = _model(sd)
    with pytest.raises(ValueError):
        m.emg_net.encode_flat(x.cuda(), subjects=subj.cuda())


def test_dataset_subject_ids_and_model_step():
    """TaskWrapper(with_subjects) labels every class row with its person; Model.forward/loss/backward run on it in
    training and in the voted evaluation (25 windows per class row share the row's subject)."""
    ds = DB23(device="cuda", mixed=True)
    ds.load_synthetic(with_glove=False)
    tw = TaskWrapper(ds, with_glove=False)
    tw.with_subjects = True
    tw.set_train()
    items = torch.arange(64)
    EMG, GLOVE, label = tw.get_batch(items)
    subj = EMG._cp_subjects
    assert subj.shape == (64, 41)
    # against the table itself: row id -> (class, person, rep, t) of the (41, 46, R, 100, 12) split tensor
    rows = tw.emg_rand[:, items.cuda()].t()
    sub = ds.EMG[ds.tasks_mask][:, ds.people_mask][:, :, ds.rep_mask][:, :, :, :100]
    P, R = sub.shape[1], sub.shape[2]
    k = rows % ds.D
    person = k // (R * 100)
    assert torch.equal(ds.people_mask[person], subj)
    torch.manual_seed(42)
    model = Model(dict(PARAMS, reg_emg=1e-5, reg_glove=1e-5), adabn=True, device="cuda")
    model.emg_net.per_subject = True
    model.set_train()
    loss = model.loss(model.forward(EMG, GLOVE, label.reshape(-1)), label.reshape(-1))
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = OM.contrastive_loss(OM.forward_logits(sd, EMG.cpu(), True, True, subjects=subj.cpu()), True)
    assert abs(loss.item() - ref["loss"].item()) < 1e-5 * ref["loss"].item()
    (loss + model.l2()).backward()
    assert all(torch.isfinite(p.grad).all() for p in model.emg_net.parameters())
    tw.set_test()
    model.set_test()
    EMG, GLOVE, label = tw.get_batch(torch.arange(24))
    with torch.no_grad():
        logits = model.forward(EMG, GLOVE, label.reshape(-1))
        loss = model.loss(logits, label.reshape(-1))
    ref_logits = OM.forward_logits(sd, EMG.cpu(), True, False, subjects=EMG._cp_subjects.cpu())
    assert float((logits.cpu() - ref_logits).abs().max()) < 1e-5
    ref = OM.contrastive_loss(ref_logits, False, W=25)
    assert np.array_equal(model.y_pred_raw(), ref["y_pred"])


def test_train_script_flag(tmp_path):
    from contrastiveprosthetics_b200 import train as cptrain
    argv = ["--final_epochs=1", "--crossval_size=1", "--crossval_epochs=1", "--batch_size=512", "--synthetic",
            "--no_verbose", "--per_subject_adabn", "--test", f"--data_dir={tmp_path}/data/",
            f"--checkpoint_dir={tmp_path}/ckpt/"]
    loss, acc = cptrain.main(cptrain.build_parser().parse_args(argv))
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0
