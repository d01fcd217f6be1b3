"""Oracle (TEST INFRASTRUCTURE): the batch x batch CLIP variant of the head and the glove-angle tower.

PARITY UNPINNED: the reference never shipped this variant in runnable form (SURVEY.md section 0 and
8c) -- there is no golden vector to pin it to.  It is restated from:
  * models.py:112-130   the contrastive branch (L2-normalise both towers, inner products), generalised
                        from the per-group 41 x 41 `bmm` to one batch x batch matrix;
  * models.py:65        "modeled after https://github.com/openai/CLIP": symmetric cross-entropy with
                        `arange` targets, logits multiplied by `logit_scale.exp()`;
  * models.py:81,129    `logit_scale` (= 0.0 at init, so the multiplier is 1.0 unless the caller trains it);
  * utils.py:56-59      flat sampling (commented): item -> one window, label = row // D;
  * models.py:384-429   the commented glove tower: Linear(GLOVE_DIM -> 256, no bias) -> BN -> ReLU ->
                        3 x [Linear(256 -> 256) -> ReLU -> BN -> Dropout] -> Linear(256 -> d_e, no bias).
Everything is op-level torch-CPU arithmetic (float32, or float64 for error bars).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

BN_EPS = 1e-5
GLOVE_HIDDEN = 256
GLOVE_BLOCKS = 3


def clip_loss(E, G, logit_scale=0.0, row0=0):
    """E (n, d), G (B, d) un-normalised embeddings; rows of E are global samples row0 .. row0+n-1.
    With n == B this is the full symmetric loss; returns dict(loss, logits, pred, n_correct).
    loss = 1/2 [ mean_i CE(S[i, :], i) + mean_j CE(S[:, j], j) ],  S = exp(logit_scale) * Ehat Ghat^T."""
    Eh = E / E.norm(dim=-1, keepdim=True)
    Gh = G / G.norm(dim=-1, keepdim=True)
    s = torch.as_tensor(logit_scale, dtype=E.dtype).exp()
    S = s * (Eh @ Gh.t())
    n = E.shape[0]
    tgt = torch.arange(row0, row0 + n)
    out = {"logits": S, "pred": S.argmax(dim=1), "n_correct": int((S.argmax(dim=1) == tgt).sum())}
    if n == G.shape[0]:
        out["loss"] = 0.5 * (F.cross_entropy(S, tgt) + F.cross_entropy(S.t(), tgt))
    return out


def clip_loss_sharded(E_parts, G_parts, logit_scale=0.0):
    """What `world` ranks compute together: rank r owns rows E_parts[r] / G_parts[r];
    embeddings are all-gathered, every rank evaluates its row strip, column sums are all-reduced and the
    column-tower gradients reduce-scattered.  Returns the global loss and the per-rank gradients of the
    GLOBAL loss w.r.t. the rank's own (un-normalised) embeddings -- computed here simply by autograd
    on the concatenation (the sharded kernel path must reproduce exactly this)."""
    E = torch.cat(E_parts).detach().clone().requires_grad_(True)
    G = torch.cat(G_parts).detach().clone().requires_grad_(True)
    res = clip_loss(E, G, logit_scale)
    res["loss"].backward()
    sizes = [p.shape[0] for p in E_parts]
    return res["loss"].detach(), list(E.grad.split(sizes)), list(G.grad.split(sizes))


# ------------------------------------------------------------------------------- glove tower
def glove_init_state(seed=7, glove_dim=20, d_e=16):
    """Parameters of the commented glove tower (models.py:384-429), keys as `glove_net.<seq>.<idx>...`
    would have been had the block been live: linear.1 (no bias), linear.2 (BN), then blocks at
    4/6, 8/10, 12/14 (Linear / BN), last.0 (projection)."""
    torch.manual_seed(seed)
    sd = {}

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().clone()

    put("glove_net.linear.1", nn.Linear(glove_dim, GLOVE_HIDDEN, bias=False))
    put("glove_net.linear.2.bn", nn.BatchNorm1d(GLOVE_HIDDEN, momentum=0, track_running_stats=False))
    for b in range(GLOVE_BLOCKS):
        put(f"glove_net.linear.{4 + 4 * b}", nn.Linear(GLOVE_HIDDEN, GLOVE_HIDDEN))
        put(f"glove_net.linear.{6 + 4 * b}.bn", nn.BatchNorm1d(GLOVE_HIDDEN, momentum=0, track_running_stats=False))
    put("glove_net.last.0", nn.Linear(GLOVE_HIDDEN, d_e, bias=False))
    return sd


def glove_forward(sd, glove, dp=0.0, dropout_masks=None, relu_masks=None, taps=None):
    """glove (n, glove_dim) -> (n, d_e).  AdaBN semantics (batch statistics, models.py:17-25).
    dropout_masks: 3 x (n, 256) keep masks (scaled by 1/(1-dp)) or None.  relu_masks: optional list of
    4 boolean (n,256) patterns imposed on the ReLUs (kink-controlled gradient tests)."""
    def relu(x, i):
        if taps is not None:
            taps.append(x.detach())              # ReLU inputs (how close the kinks are)
        if relu_masks is not None and relu_masks[i] is not None:
            return x * relu_masks[i].to(x.dtype)
        return F.relu(x)

    def bn(x, prefix):
        return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.0, BN_EPS)

    out = F.linear(glove, sd["glove_net.linear.1.weight"])
    out = relu(bn(out, "glove_net.linear.2.bn"), 0)                       # Linear -> BN -> ReLU
    for b in range(GLOVE_BLOCKS):
        li, bi = 4 + 4 * b, 6 + 4 * b
        out = F.linear(out, sd[f"glove_net.linear.{li}.weight"], sd[f"glove_net.linear.{li}.bias"])
        out = bn(relu(out, 1 + b), f"glove_net.linear.{bi}.bn")           # Linear -> ReLU -> BN -> Dropout
        if dp > 0.0 and dropout_masks is not None:
            out = out * dropout_masks[b].to(out.dtype) / (1.0 - dp)
    return F.linear(out, sd["glove_net.last.0.weight"])
