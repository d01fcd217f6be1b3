import numpy as np
import torch


def rel_err(got, ref):
    got = torch.as_tensor(got).detach().cpu().double().reshape(-1)
    ref = torch.as_tensor(ref).detach().cpu().double().reshape(-1)
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


def load_sd(model, sd):
    model.load_state_dict({k: v.clone() for k, v in sd.items()})


def perturbed_state(seed, adabn):
    """Reference-initialised weights with non-trivial BN affine / running stats."""
    from oracle import model as OM
    sd = OM.init_state(seed, adabn)
    g = torch.Generator().manual_seed(seed + 1)
    for k in list(sd.keys()):
        is_bn = (".bn." in k) if adabn else any(k.startswith(f"emg_net.{s}.{i}.") for s, idx in
                                                  (("conv_emg", (2, 5)), ("linear", (2, 5, 8, 11, 15, 19, 23)))
                                                  for i in idx)
        if not is_bn:
            continue
        if k.endswith(".weight"):
            sd[k] = 1.0 + 0.2 * torch.randn(sd[k].shape, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.1 * torch.randn(sd[k].shape, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.3 * torch.randn(sd[k].shape, generator=g)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
    return sd


# ---------------------------------------------------------------- encoder parity measurement (tests + scripts/parity_report.py)
def scale_state(sd, adabn, gamma=1.0, weight=1.0):
    """Copy of `sd` with every BatchNorm weight multiplied by `gamma` and every conv / linear weight (and bias)
    by `weight` -- the dynamic-range cases of the fp16-plane operand format."""
    out = {}
    for k, v in sd.items():
        is_bn = (".bn." in k) if adabn else any(k.startswith(f"emg_net.{s}.{i}.") for s, idx in
                                                  (("conv_emg", (2, 5)), ("linear", (2, 5, 8, 11, 15, 19, 23)))
                                                  for i in idx)
        v = v.clone()
        if v.is_floating_point() and k.startswith("emg_net."):
            if is_bn and k.endswith(".weight"):
                v = v * gamma
            elif not is_bn and k.endswith((".weight", ".bias")):
                v = v * weight
        out[k] = v
    return out


def encoder_parity_errors(sd, adabn, x, d_emb, engine, dp=0.0, masks=None, fp64=False):
    """Forward + backward of the EMG encoder through the C ABI vs the fp32 oracle with the kernel's ReLU pattern
    injected (and, fp64=True, vs the float64 oracle).  Returns a dict of norm-wise relative errors:
    {"emb":..., "stage<k>":..., "grad|<param>":..., ["emb64", "grad64|<param>", "oracle32_vs_64|<param>"]}."""
    from contrastiveprosthetics_b200.models import Model
    from oracle import model as OM
    params = {'d_e': 16, 'dp_emg': dp, 'dp_glove': 0.0, 'reg_emg': 0.0, 'reg_glove': 0.0}
    m = Model(params, adabn=adabn, device="cuda")
    load_sd(m, sd)
    m.emg_net.engine = engine
    m.train(True)
    m.emg_net.debug_tap = {}
    if masks is not None:
        m.emg_net.ext_dropout_masks = torch.stack(masks).to(torch.uint8).cuda().contiguous()
    emb = m.emg_net.encode_flat(x.cuda())
    emb.backward(d_emb.cuda())
    taps = [m.emg_net.read_activation(s, 0).cpu() for s in range(9)]
    got = {"emg_net." + k: p.grad.detach().cpu() for k, p in m.emg_net.named_parameters()}
    emb = emb.detach().cpu()
    del m
    torch.cuda.empty_cache()
    pattern = [t > 0 for t in taps]

    def oracle(dtype, relu_masks, want_taps):
        p = {}
        for k, v in sd.items():
            p[k] = v.to(dtype).clone().requires_grad_(k in OM.trainable_keys(sd)) if v.is_floating_point() else v.clone()
        ot = {} if want_taps else None
        e = OM.encoder_forward(p, x.to(dtype), adabn, True, masks, dp, taps=ot, relu_masks=relu_masks)
        e.backward(d_emb.to(dtype))
        grads = {k: p[k].grad for k in p if k.startswith("emg_net.") and getattr(p[k], "grad", None) is not None}
        return e.detach(), grads, ot

    out = {}
    e32, g32, ot = oracle(torch.float32, None, True)
    out["emb"] = rel_err(emb, e32)
    for s in range(9):
        out[f"stage{s}"] = rel_err(taps[s], ot[f"relu{s}"].detach())
    flips = sum(int((pattern[s] != (ot[f"relu{s}"] > 0)).sum()) for s in range(9))
    out["relu_flip_fraction"] = flips / sum(t.numel() for t in taps)
    del ot
    for k in g32:
        out[f"grad_unconditioned|{k}"] = rel_err(got[k], g32[k])
    e32p, g32p, _ = oracle(torch.float32, pattern, False)
    for k in g32p:
        out[f"grad|{k}"] = rel_err(got[k], g32p[k])
    if fp64:
        e64, g64, _ = oracle(torch.float64, pattern, False)
        out["emb64"] = rel_err(emb, e64)
        out["emb_oracle32_vs_64"] = rel_err(e32p, e64)
        num = den = num32 = 0.0
        for k in g64:
            out[f"grad64|{k}"] = rel_err(got[k], g64[k])
            out[f"oracle32_vs_64|{k}"] = rel_err(g32p[k], g64[k])
            out[f"norm64|{k}"] = float(g64[k].norm())
            num += float((got[k].double() - g64[k]).norm()) ** 2
            num32 += float((g32p[k].double() - g64[k]).norm()) ** 2
            den += float(g64[k].norm()) ** 2
        # the whole gradient as ONE vector (what an optimizer step sees)
        out["grad64_global"] = (num / den) ** 0.5
        out["oracle32_vs_64_global"] = (num32 / den) ** 0.5
    return out


def worst(errs, prefix):
    items = [(v, k) for k, v in errs.items() if k.startswith(prefix)]
    return max(items) if items else (0.0, None)
