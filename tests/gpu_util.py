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
