"""Run the reference's OWN code (staged by oracle/ref_fetch.py in oracle/_ref) on the host CPU.

Test / bench infrastructure only (`bench.py --impl reference` and the cpu_baseline leg): nothing in the product
imports this.  The modules are loaded unmodified apart from the two mechanical shims of SURVEY.md section 8c:
  1. stub modules for imports that are not installed and not on the hot path (line_profiler, ipdb, pyxis, matplotlib);
  2. the hard-coded "cuda" device strings (utils.py:19,24,192; load.py:25; models.py:19,29,231,353) become "cpu".
"""
import os
import sys
import time
import types

import numpy as np
import torch

from .ref_fetch import MODULES, REF_DST, staged

T = 41


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference(ref_dir=REF_DST):
    class _LP:
        def print_stats(self, *a, **k):
            pass

        def __call__(self, f):
            return f

    for name, kw in (("line_profiler", {"LineProfiler": _LP}), ("ipdb", {}), ("pyxis", {})):
        if name not in sys.modules:
            _stub(name, **kw)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
    mods = {}
    for name in MODULES:
        path = os.path.join(ref_dir, name + ".py")
        src = open(path).read().replace('"cuda"', '"cpu"')
        m = types.ModuleType(name)
        m.__file__ = path
        sys.modules[name] = m
        exec(compile(src, path, "exec"), m.__dict__)
        mods[name] = m
    return mods


def synthetic_dataset(R, db2=True, seed=0):
    """The reference's DB23 + TaskWrapper (load.py, utils.py) over seeded NinaPro-shaped tensors (emg.pt is not
    available offline): EMG (46 people, 41 stimuli, 6 reps, 100, 12) ~ N(0,1) + per-class channel offsets."""
    g = torch.Generator().manual_seed(seed)
    emg = torch.randn(46, 41, 6, 100, 12, generator=g)
    emg = emg + (0.5 * torch.randn(41, 12, generator=g))[None, :, None, None, :]
    ds = R["load"].DB23(db2=db2)
    ds.EMG = emg.transpose(0, 1)                     # load.py:70
    ds.glover.GLOVE = torch.randn(41, 600, 20, generator=g)
    ds.GLOVE = ds.glover.GLOVE
    return R["utils"].TaskWrapper(ds)


def train_steps(n_steps, warmup, groups, params, threads=None):
    """The body of the reference's train_loop (train.py:83-108) timed per step on the host: DataLoader over the
    reference's TaskWrapper (per-item __getitem__ + default_collate), Model.forward, Model.loss (the per-group
    loop of models.py:132-173), + Model.l2, zero_grad, backward, two Adam steps.  Returns (windows/s, ms/step,
    threads)."""
    import torch.optim as optim
    import torch.utils.data as data
    if not staged():
        raise RuntimeError("oracle/_ref is not staged (python -m oracle.ref_fetch in the build container)")
    torch.set_num_threads(threads or os.cpu_count() or 1)
    R = load_reference()
    torch.manual_seed(42)
    tw = synthetic_dataset(R)
    model = R["models"].Model(params=dict(params), train_model=True, adabn=True, device="cpu").to(torch.float32)
    opt_e = optim.Adam(model.emg_net.parameters(), lr=params['lr_emg'], weight_decay=0)
    opt_g = optim.Adam(model.glove_net.parameters(), lr=params['lr_glove'], weight_decay=0)
    tw.set_train()
    model.set_train()
    loader = iter(data.DataLoader(tw, batch_size=groups, shuffle=True))
    times = []
    for s in range(warmup + n_steps):
        t0 = time.perf_counter()
        EMG, GLOVE, label = next(loader)
        label = label.reshape(-1)
        logits = model.forward(EMG, GLOVE, label)
        loss = model.loss(logits, label)
        _ = loss.item()                               # train.py:99
        loss = loss + model.l2()
        opt_e.zero_grad(set_to_none=True)
        opt_g.zero_grad(set_to_none=True)
        loss.backward()
        opt_e.step()
        opt_g.step()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    return groups * T / (ms / 1e3), ms, torch.get_num_threads()
