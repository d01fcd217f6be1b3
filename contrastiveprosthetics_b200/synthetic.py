"""Seeded synthetic NinaPro-DB2/DB3-shaped tensors (the real dataset is not available offline).

Shapes follow the reference's on-disk tensors: emg.pt is (46 people, 41 stimuli, 6 reps, 100
samples, 12 channels) fp32, already normalised (load.py:118,143-147); glove.pt is
(41, 39*6*25, 20) (utils.py:197-246).  A per-class channel offset makes the classes learnable.
Host-side (CPU generator) so every side of a parity test sees identical values.
"""
import numpy as np
import torch


def synth_emg(seed=0, people=46):
    g = torch.Generator().manual_seed(seed)
    emg = torch.randn(people, 41, 6, 100, 12, generator=g)
    off = 0.5 * torch.randn(41, 12, generator=g)
    return emg + off[None, :, None, None, :]


def synth_glove(seed=1, dim=20, class_offset=False):
    """dim = 20 (GLOVE_DIM, constants.py:96) or 22 (all CyberGlove sensors, config 5).  class_offset adds
    a per-class posture offset so that glove rows are informative about the class (config 5)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(41, 5850, dim, generator=g)
    return x + torch.randn(41, 1, dim, generator=g) if class_offset else x


def fixed_perm(T, D, seed):
    """Deterministic per-class permutation + class offset, same form as TaskWrapper.return_rand
    (utils.py:34-36) but from numpy so CPU and CUDA sides can share it."""
    r = np.random.RandomState(seed).rand(T, D)
    return np.argsort(r, axis=-1, kind="stable") + (np.arange(T) * D)[:, None]
