"""Oracle (TEST INFRASTRUCTURE): encoder, contrastive head, loss, accuracy, vote, l2 on the CPU.

Functional restatement of /root/reference/code/models.py over a plain state dict
with the reference's key names.  torch-CPU arithmetic (the
reference's own library); `dtype=torch.float64` gives the error-bar variant.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

EMG_DIM = 12
N_TASKS = 41
VOTE_LOOP = 250         # PREDICTION_WINDOW: `for win in range(1, PREDICTION_WINDOW)` (models.py:153)
BN_EPS = 1e-5

# (linear index, bn index, dropout-after?) inside emg_net.linear  (models.py:266-298)
LINEAR_BLOCKS = [(0, 2, False), (3, 5, False), (6, 8, False), (9, 11, True),
                 (13, 15, True), (17, 19, True), (21, 23, True)]
CONV_BLOCKS = [(0, 2), (3, 5)]     # (conv index, bn index) inside emg_net.conv_emg (models.py:248-264)


def bn_prefix(adabn, seq, idx):
    """AdaBN wraps the BatchNorm in `.bn` (models.py:17-35); stock BN does not (models.py:241-243)."""
    return f"emg_net.{seq}.{idx}.bn" if adabn else f"emg_net.{seq}.{idx}"


def init_state(seed=42, adabn=True, d_e=16, prediction=False):
    """Parameter init in the reference's construction order (models.py:67-85, 231-317, 353-430):
    EMGNet (conv, conv, 7 linears, projection) -> GLOVENet (easy, last) -> logit_scale, so that
    `torch.manual_seed(seed)` consumes the CPU generator exactly as `Model(...)` does.
    prediction=True: the classifier heads of models.py:300-309 / 413-421 instead of the projections."""
    torch.manual_seed(seed)
    sd = {}

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().clone()

    def bn(seq, idx, feats, two_d):
        cls = nn.BatchNorm2d if two_d else nn.BatchNorm1d
        m = cls(feats, momentum=0, track_running_stats=False) if adabn else cls(feats)
        put(bn_prefix(adabn, seq, idx), m)

    put("emg_net.conv_emg.0", nn.Conv2d(1, 64, (3, 3), padding=(1, 1)))
    bn("conv_emg", 2, 64, True)
    put("emg_net.conv_emg.3", nn.Conv2d(64, 64, (3, 3), padding=(1, 1)))
    bn("conv_emg", 5, 64, True)
    fan_in = EMG_DIM * 64
    for li, bi, _ in LINEAR_BLOCKS:
        put(f"emg_net.linear.{li}", nn.Linear(fan_in, 512))
        bn("linear", bi, 512, False)
        fan_in = 512
    if prediction:
        put("emg_net.last.0", nn.Linear(512, 128))
        m = nn.BatchNorm1d(128, momentum=0, track_running_stats=False) if adabn else nn.BatchNorm1d(128)
        put("emg_net.last.2.bn" if adabn else "emg_net.last.2", m)
        put("emg_net.last.3", nn.Linear(128, N_TASKS, bias=False))
        put("glove_net.easy.0", nn.Linear(N_TASKS, d_e))
        put("glove_net.last.0", nn.Linear(512 // 2, 128))
        m = nn.BatchNorm1d(128, momentum=0, track_running_stats=False) if adabn else nn.BatchNorm1d(128)
        put("glove_net.last.2.bn" if adabn else "glove_net.last.2", m)
        put("glove_net.last.4", nn.Linear(128, N_TASKS, bias=False))
    else:
        put("emg_net.last.0", nn.Linear(512, d_e, bias=False))
        put("glove_net.easy.0", nn.Linear(N_TASKS, d_e))
        put("glove_net.last.0", nn.Linear(512 // 2, d_e, bias=False))
    out = {"logit_scale": torch.ones([]) * np.log(1) / 0.07}    # models.py:81  (== 0.0)
    out.update(sd)
    return out


def trainable_keys(sd):
    return [k for k in sd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))
            and k != "logit_scale"]


def _batch_norm(sd, prefix, x, adabn, training, new_stats, subjects=None):
    """models.py:17-35 (AdaBN: batch statistics in train AND eval) or stock nn.BatchNorm.
    subjects (N,) ints: PER-SUBJECT AdaBN -- "momentum = 0 and batch per subject in order to have adaptive
    normalization" (models.py:245, described there and never implemented: PARITY UNPINNED): the batch statistics of
    every feature are taken over the rows of one subject at a time; gamma / beta are shared."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if adabn and subjects is not None:
        parts, index = [], []
        for s in torch.unique(subjects).tolist():
            rows = torch.nonzero(subjects == s).reshape(-1)
            parts.append(F.batch_norm(x[rows], None, None, w, b, True, 0.0, BN_EPS))
            index.append(rows)
        inv = torch.empty(x.shape[0], dtype=torch.long)
        inv[torch.cat(index)] = torch.arange(x.shape[0])
        return torch.cat(parts)[inv]
    if adabn:
        return F.batch_norm(x, None, None, w, b, True, 0.0, BN_EPS)
    rm = sd[prefix + ".running_mean"].clone()
    rv = sd[prefix + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, w, b, training, 0.1, BN_EPS)
    if training and new_stats is not None:
        new_stats[prefix + ".running_mean"] = rm.detach()
        new_stats[prefix + ".running_var"] = rv.detach()
        new_stats[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
    return y


def _relu(out, relu_masks, i):
    """ReLU, or -- for kink-controlled gradient parity -- multiplication by an injected 0/1 mask (the
    ReLU pattern another implementation took: pre-activations within float noise of 0 otherwise
    flip between any two fp32 evaluations and change the gradient by O(1/sqrt(#elements)))."""
    if relu_masks is None:
        return F.relu(out)
    return out * relu_masks[i].to(out.dtype)


def encoder_forward(sd, x, adabn=True, training=True, dropout_masks=None, dp=0.0, new_stats=None,
                    taps=None, relu_masks=None, prediction=False, subjects=None):
    """EMGNet.forward up to the projection (models.py:319-323): x (N,12) -> emb (N,d_e).

    dropout_masks: optional list of 4 {0,1} tensors (N,512) for the dropout after linear blocks
    4..7 (models.py:282-297); applied as mask/(1-dp) in training.  None and dp==0 -> identity.
    taps: optional dict that receives intermediate activations (for layer-level parity tests).
    relu_masks: optional list of 9 0/1 tensors (2 conv stages (N,64,1,12), 7 linear stages (N,512)).
    subjects: optional (N,) subject id per window -> per-subject AdaBN (see _batch_norm)."""
    out = x.reshape(-1, 1, 1, EMG_DIM)
    stage = 0
    for ci, bi in CONV_BLOCKS:
        out = F.conv2d(out, sd[f"emg_net.conv_emg.{ci}.weight"], sd[f"emg_net.conv_emg.{ci}.bias"],
                       padding=(1, 1))
        if taps is not None:
            taps[f"pre{stage}"] = out
        out = _relu(out, relu_masks, stage)
        if taps is not None:
            taps[f"relu{stage}"] = out
        stage += 1
        out = _batch_norm(sd, bn_prefix(adabn, "conv_emg", bi), out, adabn, training, new_stats, subjects)
    out = out.flatten(1)
    d = 0
    for li, bi, has_dp in LINEAR_BLOCKS:
        out = F.linear(out, sd[f"emg_net.linear.{li}.weight"], sd[f"emg_net.linear.{li}.bias"])
        if taps is not None:
            taps[f"pre{stage}"] = out
        out = _relu(out, relu_masks, stage)
        if taps is not None:
            taps[f"relu{stage}"] = out
        stage += 1
        out = _batch_norm(sd, bn_prefix(adabn, "linear", bi), out, adabn, training, new_stats, subjects)
        if has_dp:
            if training and dropout_masks is not None and dp > 0:
                out = out * dropout_masks[d].to(out.dtype) / (1.0 - dp)
            d += 1
    if prediction:
        # EMGNet.last with prediction=True (models.py:300-309): Linear(512,128) -> ReLU -> BN(128) -> Linear(128,41)
        if taps is not None:
            taps["trunk"] = out
        out = F.linear(out, sd["emg_net.last.0.weight"], sd["emg_net.last.0.bias"])
        out = _relu(out, relu_masks, 9) if relu_masks is not None and len(relu_masks) > 9 else F.relu(out)
        pre = "emg_net.last.2.bn" if adabn else "emg_net.last.2"
        out = _batch_norm(sd, pre, out, adabn, training, new_stats)
        return F.linear(out, sd["emg_net.last.3.weight"])
    return F.linear(out, sd["emg_net.last.0.weight"])


def prediction_step(sd, EMG, labels, adabn=True, training=True, dropout_masks=None, dp=0.0, new_stats=None):
    """Model.forward + Model.loss with prediction=True, glove=False (models.py:113-119, 175-196) outside the (broken)
    vote evaluation: features = z / ||z||, loss = F.cross_entropy(features, labels), accuracy = mean(argmax == labels).
    EMG: anything reshapeable to (-1,12) with one row per label.  Returns dict(features, loss, pred, correct)."""
    z = encoder_forward(sd, EMG.reshape(-1, EMG_DIM), adabn, training, dropout_masks, dp, new_stats, prediction=True)
    feats = z / z.norm(dim=-1, keepdim=True)
    labels = labels.reshape(-1).to(torch.long)
    loss = F.cross_entropy(feats, labels)
    pred = F.softmax(feats, dim=-1).argmax(-1)
    correct = float((pred.detach().numpy() == labels.numpy()).mean())
    return {"features": feats, "loss": loss, "pred": pred, "correct": correct}


def class_table(sd):
    """GLOVENet.forward default branch (models.py:457-458): Linear(41->d_e)(one_hot(label)) is the
    label-th row of W^T + b."""
    return sd["glove_net.easy.0.weight"].t() + sd["glove_net.easy.0.bias"][None, :]


def forward_logits(sd, EMG, adabn=True, training=True, dropout_masks=None, dp=0.0, new_stats=None, subjects=None):
    """Model.forward contrastive branch (models.py:121-130) incl. the EMGNet regrouping
    (models.py:337-341) and the GLOVENet expand in eval (models.py:463-464).

    EMG: (B,41,W,1,12).  Returns logits (B*W,41,41): group (b,w), row = EMG class, col = class table."""
    B, T, W = EMG.shape[0], EMG.shape[1], EMG.shape[2]
    if subjects is not None:                   # (B,41) subject of every class row -> one id per window
        subjects = subjects.reshape(B, T, 1).expand(B, T, W).reshape(-1)
    emb = encoder_forward(sd, EMG.reshape(-1, EMG_DIM), adabn, training, dropout_masks, dp, new_stats,
                          subjects=subjects)
    emb = emb.reshape(B, T, W, -1).transpose(1, 2).reshape(B * W, T, -1)
    emb = emb / emb.norm(dim=-1, keepdim=True)
    tab = class_table(sd).to(emb.dtype)
    tab = tab / tab.norm(dim=-1, keepdim=True)
    return torch.matmul(emb, tab.t())


def prefix_mode(pred):
    """pred (W,41) ints -> modes (W,41): modes[w-1] = pred[:w].mode(0) with the torch-CPU tie rule
    (smallest label among the most frequent; SURVEY.md A.3, models.py:154)."""
    W, T = pred.shape
    out = np.empty((W, T), dtype=np.int64)
    for c in range(T):
        counts = np.zeros(N_TASKS + 1, dtype=np.int64)
        for w in range(W):
            counts[pred[w, c]] += 1
            out[w, c] = int(np.argmax(counts))          # first max == smallest label
    return out


def vote_group(pred):
    """models.py:151-163 for one group.  pred (25,41) -> (votes_correct (249,) int, y_pred (41,))."""
    modes = prefix_mode(pred)
    W = pred.shape[0]
    tgt = np.arange(pred.shape[1])
    correct_per_w = (modes == tgt[None, :]).sum(1)       # (W,)
    idx = np.minimum(np.arange(1, VOTE_LOOP), W) - 1     # pred[:win] clamps at W rows
    return correct_per_w[idx], modes[-1]


def contrastive_loss(logits, training, W=1, with_acc=True, argmax_via_softmax=True):
    """Model.loss -> contrastive_loopy_loss x2 (models.py:198-208, 132-173), vectorised.

    logits (G,41,41) with G = B (train) or B*W (eval, group order (b,w)).
    Returns dict: loss (tensor), correct_counts (B,) int (#correct of 41 per item, after the full
    vote in eval), voting_counts (B,249) int | None, y_pred (B,41) | None, preds (G,41) int.
    The reference accumulates count/41 in float32 (models.py:134,166,171): see `correct_float`."""
    G, T = logits.shape[0], logits.shape[-1]
    tgt = torch.arange(T).repeat(G)
    loss_e = F.cross_entropy(logits.reshape(-1, T), tgt)
    loss_g = F.cross_entropy(logits.transpose(1, 2).reshape(-1, T), tgt)
    res = {"loss": (loss_e + loss_g) / 2, "loss_e": loss_e, "loss_g": loss_g}
    if not with_acc:
        return res
    with torch.no_grad():
        src = F.softmax(logits, dim=-1) if argmax_via_softmax else logits
        preds = src.argmax(-1).numpy()                    # (G,41)  models.py:148
    res["preds"] = preds
    if training:
        res["correct_counts"] = (preds == np.arange(T)[None, :]).sum(1)
        res["voting_counts"] = None
        res["y_pred"] = None
    else:
        B = G // W
        p = preds.reshape(B, W, T)
        votes = np.empty((B, VOTE_LOOP - 1), dtype=np.int64)
        y_pred = np.empty((B, T), dtype=np.int64)
        for b in range(B):
            votes[b], y_pred[b] = vote_group(p[b])
        res["voting_counts"] = votes
        res["y_pred"] = y_pred
        res["correct_counts"] = votes[:, -1]
    return res


def correct_float(correct_counts, T=N_TASKS):
    """models.py:134,166,170-172: float32 running sum of (count/41 as float64) then / bs."""
    acc = np.float32(0.0)
    for c in correct_counts:
        acc = np.float32(acc + np.float32(np.float64(c) / T))
    return float(np.float32(acc / np.float32(len(correct_counts))))


def l2_penalty(sd, reg_emg, reg_glove):
    """Model.l2 / EMGNet.l2 / GLOVENet.l2 (models.py:225-228, 344-349, 467-472): sum of UN-squared
    Frobenius norms over parameters whose name has neither 'bn' nor 'bias' (so stock-BN gammas,
    named `linear.N.weight`, are included when --no_adabn)."""
    tot_e, tot_g = 0, 0
    for k in trainable_keys(sd):
        if "bn" in k or "bias" in k:
            continue
        if k.startswith("emg_net."):
            tot_e = tot_e + torch.norm(sd[k])
        elif k.startswith("glove_net."):
            tot_g = tot_g + torch.norm(sd[k])
    return tot_g * reg_glove + tot_e * reg_emg


def train_step_grads(sd, EMG, adabn=True, dp=0.0, dropout_masks=None, reg_emg=0.0, reg_glove=0.0,
                     dtype=torch.float32, subjects=None):
    """One train_loop iteration up to backward (train.py:95-105): returns (res, grads, new_stats)."""
    p = {}
    for k, v in sd.items():
        if k in trainable_keys(sd):
            p[k] = v.detach().to(dtype).clone().requires_grad_(True)
        elif v.is_floating_point():
            p[k] = v.detach().to(dtype).clone()
        else:
            p[k] = v.clone()
    new_stats = {}
    logits = forward_logits(p, EMG.to(dtype), adabn, True, dropout_masks, dp, new_stats, subjects=subjects)
    res = contrastive_loss(logits, True)
    l2 = l2_penalty(p, reg_emg, reg_glove)
    total = res["loss"] + l2
    total.backward()
    res["l2"] = l2.detach() if torch.is_tensor(l2) else torch.tensor(float(l2))
    grads = {k: (p[k].grad if p[k].grad is not None else torch.zeros_like(p[k]))
             for k in trainable_keys(sd)}
    res["logits"] = logits.detach()
    res["total"] = total.detach()
    return res, grads, new_stats


def adam_update(param, grad, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor rule, weight_decay 0 (train.py:72-73)."""
    m.mul_(b1).add_(grad, alpha=1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    param.addcdiv_(m, denom, value=-lr / bc1)
