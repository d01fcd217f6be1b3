#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference on the CPU (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/code/{constants,utils,load,models}.py through tests/golden/_refload.py
(stub imports + "cuda"->"cpu"), feeds seeded synthetic NinaPro-shaped tensors, and stores the
reference's own outputs in tests/golden/*.npz.  The fixtures travel to the GPU box; the reference
does not.  tests/test_oracle_golden.py replays them through oracle/, tests/test_gpu_*.py through the
CUDA path.
"""
import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refload import load_reference  # noqa: E402

torch.set_num_threads(8)
R = load_reference()
RM, RL, RU, RC = R["models"], R["load"], R["utils"], R["constants"]


def synth_emg(seed=0):
    """(46 people, 41 stimuli, 6 reps, 100, 12) like emg.pt, N(0,1) + per-class channel offset."""
    g = torch.Generator().manual_seed(seed)
    emg = torch.randn(46, 41, 6, 100, 12, generator=g)
    off = 0.5 * torch.randn(41, 12, generator=g)
    return emg + off[None, :, None, None, :]


def synth_glove(seed=1):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(41, 5850, 20, generator=g)


def make_dataset(db2):
    ds = RL.DB23(db2=db2)
    ds.EMG = synth_emg().transpose(0, 1)            # load.py:70 transposes people<->tasks
    ds.glover.GLOVE = synth_glove()
    ds.GLOVE = ds.glover.GLOVE
    return ds


def fixed_perm(T, D, seed):
    """Deterministic stand-in for TaskWrapper.return_rand (utils.py:34-36), numpy so that the same
    indices can be injected on every side (torch CPU/CUDA RNG streams differ)."""
    r = np.random.RandomState(seed).rand(T, D)
    return np.argsort(r, axis=-1, kind="stable") + (np.arange(T) * D)[:, None]


def sample_idx(n, k, seed):
    return np.sort(np.random.RandomState(seed).choice(n, size=min(k, n), replace=False))


# --------------------------------------------------------------------------- A. dataset
def gen_dataset():
    out = {}
    for db2 in (False, True):
        ds = make_dataset(db2)
        tw = RU.TaskWrapper(ds)
        for split in ("train", "val", "test"):
            getattr(tw, "set_" + split)()
            tag = f"db2{int(db2)}_{split}"
            out[tag + "_D"] = np.int64(ds.D)
            out[tag + "_len"] = np.int64(len(ds))
            out[tag + "_twlen"] = np.int64(len(tw))
            out[tag + "_tasks"] = ds.tasks_mask.numpy()
            out[tag + "_people"] = ds.people_mask.numpy()
            out[tag + "_reps"] = ds.rep_mask.numpy()
            rows = sample_idx(ds.EMG_use.shape[0], 48, 7)
            out[tag + "_rows"] = rows
            out[tag + "_EMG_use"] = ds.EMG_use[rows].numpy()
            trow = sample_idx(ds.tensor.shape[0], 6, 8)
            out[tag + "_trows"] = trow
            out[tag + "_tensor"] = ds.tensor[trow].numpy()
            tw.emg_rand = torch.from_numpy(fixed_perm(41, ds.D, 11))
            tw.glove_rand = torch.from_numpy(fixed_perm(41, ds.glover.D, 12))
            items = np.array([0, 5, ds.D - 1])
            out[tag + "_items"] = items
            e0, g0, l0 = tw[int(items[1])]
            out[tag + "_item_emg"] = e0.numpy()
            out[tag + "_item_glove"] = g0.numpy()
            out[tag + "_item_label"] = l0.numpy()
    # RunningStats.normalize semantics (utils.py:129-130) with per-channel and scalar stats
    rs = RU.RunningStats("/tmp/_cp_golden_", complete=False)
    x = synth_emg()[0, 0, 0].numpy().astype(np.float32)      # (100,12)
    mean = np.linspace(-0.3, 0.4, 12).astype(np.float32)
    std = np.linspace(0.5, 2.0, 12).astype(np.float32)
    rs.np = True
    rs.new_mean = mean
    rs.counter = 2
    rs.new_std = (std ** 2).astype(np.float32)               # variance()*(counter-1)
    out["norm_x"] = x
    out["norm_mean"] = mean
    out["norm_std"] = rs.std()
    out["norm_y"] = rs.normalize(x)
    np.savez_compressed(os.path.join(HERE, "dataset.npz"), **out)
    print("dataset.npz", len(out), "arrays")


# --------------------------------------------------------------------------- B/C/D. model
PARAMS = {'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 3e-4, 'reg_glove': 1e-5,
          'lr_emg': 1e-3, 'lr_glove': 2e-3, 'epochs': 1}


def grad_digest(name, g, out, tag):
    g = g.detach().numpy().astype(np.float32)
    out[f"{tag}|gnorm|{name}"] = np.float64(np.linalg.norm(g.astype(np.float64)))
    if g.size <= 4096:
        out[f"{tag}|gfull|{name}"] = g
    else:
        idx = sample_idx(g.size, 64, 3)
        out[f"{tag}|gidx|{name}"] = idx
        out[f"{tag}|gval|{name}"] = g.reshape(-1)[idx]


def gen_eval(model, tw, ds, out, tag):
    """validate()/test() (train.py:27-63) with the vote loop of models.py:138-166, run on the state
    after the first training forward/backward and BEFORE the optimizer step.  All logits are stored so the integer stages (argmax,
    vote, counts) can be replayed bit-exactly from the reference's own floats."""
    tw.set_test()
    tw.emg_rand = torch.from_numpy(fixed_perm(41, ds.D, 13))
    tw.glove_rand = torch.from_numpy(fixed_perm(41, ds.glover.D, 14))
    model.set_test()
    ev_batches = [[0, 3], [10]]
    out[f"{tag}|eval_items"] = np.array([b + [-1] * (2 - len(b)) for b in ev_batches])
    ev_losses = []
    for bi, items in enumerate(ev_batches):
        samples = [tw[i] for i in items]
        EMG = torch.stack([s[0] for s in samples])
        GLOVE = torch.stack([s[1] for s in samples])
        label = torch.stack([s[2] for s in samples]).reshape(-1)
        with torch.no_grad():
            logits = model.forward(EMG, GLOVE, label)
            loss = model.loss(logits, label)
        ev_losses.append(loss.item())
        out[f"{tag}|eval_EMG{bi}"] = EMG.numpy()
        out[f"{tag}|eval_logits{bi}"] = logits.numpy()
    out[f"{tag}|eval_losses"] = np.array(ev_losses, dtype=np.float64)
    out[f"{tag}|eval_corrects"] = model.correct_raw()
    out[f"{tag}|eval_correct_mean"] = np.float64(model.correct())
    out[f"{tag}|eval_voting"] = model.voting_raw()
    out[f"{tag}|eval_y_pred"] = model.y_pred_raw()
    out[f"{tag}|eval_y_true"] = model.y_true_raw()


def gen_model():
    out = {}
    ds = make_dataset(False)
    tw = RU.TaskWrapper(ds)
    for adabn in (True, False):
        tag = "adabn" if adabn else "stockbn"
        torch.manual_seed(42)
        model = RM.Model(params=dict(PARAMS), adabn=adabn, device="cpu").to(torch.float32)
        # B. init digest
        for k, v in model.state_dict().items():
            v = v.detach().to(torch.float64).reshape(-1)
            out[f"{tag}|init|{k}"] = np.array([v.sum().item(), v.abs().sum().item()] +
                                              v[:4].tolist())
        opt_e = torch.optim.Adam(model.emg_net.parameters(), lr=PARAMS['lr_emg'], weight_decay=0)
        opt_g = torch.optim.Adam(model.glove_net.parameters(), lr=PARAMS['lr_glove'], weight_decay=0)

        # C. training steps on fixed batches (train.py:95-108); D. eval after the first step
        batches = [[0, 7, 19], [3, 1799, 42], [100, 200, 300], [5, 6, 8]]
        out[f"{tag}|train_items"] = np.array(batches)
        losses = []
        train_corrects = []
        for step, items in enumerate(batches):
            tw.set_train()
            tw.emg_rand = torch.from_numpy(fixed_perm(41, ds.D, 11))
            tw.glove_rand = torch.from_numpy(fixed_perm(41, ds.glover.D, 12))
            model.set_train()
            samples = [tw[i] for i in items]
            EMG = torch.stack([s[0] for s in samples])
            GLOVE = torch.stack([s[1] for s in samples])
            label = torch.stack([s[2] for s in samples]).reshape(-1)
            if step == 0:
                out[f"{tag}|EMG0"] = EMG.numpy()
            logits = model.forward(EMG, GLOVE, label)
            loss = model.loss(logits, label)
            losses.append(loss.item())
            train_corrects.append(model.corrects[-1])
            l2 = model.l2()
            total = loss + l2
            opt_e.zero_grad(set_to_none=True)
            opt_g.zero_grad(set_to_none=True)
            total.backward()
            if step == 0:
                out[f"{tag}|logits0"] = logits.detach().numpy()
                out[f"{tag}|l2_0"] = np.float64(l2.item())
                for n, p in model.named_parameters():
                    if p.grad is not None:
                        grad_digest(n, p.grad, out, tag)
                    else:
                        out[f"{tag}|gnone|{n}"] = np.int64(1)
            if step == 0:
                # eval on (init weights, running stats after ONE training forward): reproducible
                # tightly, unlike any post-Adam state (update ~ lr*sign(g) flips on noise-level grads)
                if not adabn:
                    for k, v in model.state_dict().items():
                        if "running" in k or "num_batches" in k:
                            out[f"{tag}|after1|{k}"] = v.numpy().copy()
                gen_eval(model, tw, ds, out, tag)
            opt_e.step()
            opt_g.step()
            if step == 0:
                for n, p in model.named_parameters():
                    v = p.detach().reshape(-1)
                    idx = sample_idx(v.numel(), 32, 5)
                    out[f"{tag}|p1idx|{n}"] = idx
                    out[f"{tag}|p1val|{n}"] = v[idx].numpy()
        out[f"{tag}|train_losses"] = np.array(losses, dtype=np.float64)
        out[f"{tag}|train_corrects"] = np.array(train_corrects, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "model.npz"), **out)
    print("model.npz", len(out), "arrays")


def gen_dropout():
    """Pins dropout placement and scaling (models.py:282-297) with the CPU generator stream."""
    out = {}
    ds = make_dataset(False)
    tw = RU.TaskWrapper(ds)
    tw.set_train()
    tw.emg_rand = torch.from_numpy(fixed_perm(41, ds.D, 11))
    tw.glove_rand = torch.from_numpy(fixed_perm(41, ds.glover.D, 12))
    params = dict(PARAMS)
    params['dp_emg'] = 0.5
    torch.manual_seed(42)
    model = RM.Model(params=params, adabn=True, device="cpu").to(torch.float32)
    model.set_train()
    samples = [tw[i] for i in (0, 7, 19)]
    EMG = torch.stack([s[0] for s in samples])
    GLOVE = torch.stack([s[1] for s in samples])
    label = torch.stack([s[2] for s in samples]).reshape(-1)
    torch.manual_seed(123)
    logits = model.forward(EMG, GLOVE, label)
    loss = model.loss(logits, label)
    loss.backward()
    out["EMG"] = EMG.numpy()
    out["logits"] = logits.detach().numpy()
    out["loss"] = np.float64(loss.item())
    grad_digest("emg_net.linear.9.weight", model.emg_net.linear[9].weight.grad, out, "dp")
    grad_digest("emg_net.last.0.weight", model.emg_net.last[0].weight.grad, out, "dp")
    np.savez_compressed(os.path.join(HERE, "dropout.npz"), **out)
    print("dropout.npz", len(out), "arrays")


def gen_prediction():
    """--prediction mode (models.py:113-119, 175-196, 300-309): the reference's Model(prediction=True) on the CPU -- init
    digests, one training forward / loss / backward on a fixed batch in both BN modes, and the AssertionError of its
    vote evaluation."""
    out = {}
    ds = make_dataset(False)
    tw = RU.TaskWrapper(ds)
    for adabn in (True, False):
        tag = "adabn" if adabn else "stockbn"
        torch.manual_seed(42)
        model = RM.Model(params=dict(PARAMS), adabn=adabn, prediction=True, device="cpu").to(torch.float32)
        out[f"{tag}|keys"] = np.array(list(model.state_dict().keys()))
        for k, v in model.state_dict().items():
            v = v.detach().to(torch.float64).reshape(-1)
            out[f"{tag}|init|{k}"] = np.array([v.sum().item(), v.abs().sum().item()] + v[:4].tolist())
        tw.set_train()
        tw.emg_rand = torch.from_numpy(fixed_perm(41, ds.D, 11))
        tw.glove_rand = torch.from_numpy(fixed_perm(41, ds.glover.D, 12))
        model.set_train()
        samples = [tw[i] for i in (0, 7, 19, 3)]
        EMG = torch.stack([s[0] for s in samples])
        GLOVE = torch.stack([s[1] for s in samples])
        label = torch.stack([s[2] for s in samples]).reshape(-1)
        feats = model.forward(EMG, GLOVE, label)
        loss = model.loss(feats, label)
        l2 = model.l2()
        (loss + l2).backward()
        out[f"{tag}|EMG"] = EMG.numpy()
        out[f"{tag}|label"] = label.numpy()
        out[f"{tag}|features"] = feats.detach().numpy()
        out[f"{tag}|loss"] = np.float64(loss.item())
        out[f"{tag}|l2"] = np.float64(l2.item())
        out[f"{tag}|correct"] = np.float64(model.corrects[-1])
        for n, p in model.named_parameters():
            if p.grad is not None:
                grad_digest(n, p.grad, out, tag)
            else:
                out[f"{tag}|gnone|{n}"] = np.int64(1)
        if not adabn:
            for k, v in model.state_dict().items():
                if "running" in k or "num_batches" in k:
                    out[f"{tag}|after1|{k}"] = v.numpy().copy()
        # the vote evaluation of this mode is broken in the reference itself
        tw.set_test()
        tw.emg_rand = torch.from_numpy(fixed_perm(41, ds.D, 13))
        tw.glove_rand = torch.from_numpy(fixed_perm(41, ds.glover.D, 14))
        model.set_test()
        e, g, l = tw[0]
        try:
            with torch.no_grad():
                model.loss(model.forward(e[None], g[None], l), l)
            out[f"{tag}|eval_error"] = np.array("none")
        except AssertionError as ex:
            out[f"{tag}|eval_error"] = np.array(str(ex))
    np.savez_compressed(os.path.join(HERE, "prediction.npz"), **out)
    print("prediction.npz", len(out), "arrays")


def copy_ref_artifacts():
    """The reference's own result artefacts = its only golden vectors."""
    dst = os.path.join(HERE, "ref_data")
    os.makedirs(dst, exist_ok=True)
    for f in ("y_true.npy", "y_pred.npy", "voting.npy", "confusion_matrix.npy",
              "cross_val_keys.npy", "cross_val_values.npy", "emg_mean.npy", "emg_std.npy"):
        shutil.copy(os.path.join("/root/reference/data", f), os.path.join(dst, f))
    # xlsx = zip of XML; openpyxl is not installed, so read the single numeric column directly
    import re
    import zipfile
    tab = {}
    for f in ("mean_grasp", "std_grasp", "min_grasp", "max_grasp"):
        xml = zipfile.ZipFile(f"/root/reference/data/{f}.xlsx").read("xl/worksheets/sheet1.xml").decode()
        vals = [float(v) for v in re.findall(r"<v>(.*?)</v>", xml)]
        tab[f] = np.array(vals[1:])           # first cell is the pandas header `0`
    np.savez_compressed(os.path.join(dst, "grasp_tables.npz"), **tab)


def gen_preprocess():
    """Offline preprocessing (load.py:85-101, utils.py:79-156): the reference's own filter / rms / time_mask /
    RunningStats on seeded raw segments (float32 like the NinaPro .mat files, plus one float64 segment)."""
    import tempfile
    rs = np.random.RandomState(2024)
    L = RC.TOTAL_WINDOW_SIZE + 2 * RC.WINDOW_EDGE
    t = np.arange(L)[:, None] / RC.Hz
    raws = []
    for s in range(6):
        # broadband noise + a 50 Hz line + slow drift, channel-dependent amplitude (roughly raw-sEMG-in-volts scale)
        amp = (1e-5 * (1 + rs.rand(1, 12) * 4)).astype(np.float64)
        x = amp * rs.randn(L, 12) + 2e-5 * np.sin(2 * np.pi * 50 * t + s) + 1e-4 * (s - 2.5) * t
        raws.append(x.astype(np.float32))
    raw = np.stack(raws)                                                   # (6, 2010, 12) float32
    mask_wrap = np.arange(0, RC.TOTAL_WINDOW_SIZE, RC.FACTOR, dtype=np.uint8)       # load.py:116 verbatim
    mask_full = np.arange(0, RC.TOTAL_WINDOW_SIZE, RC.FACTOR)
    out = {"raw": raw, "time_mask": mask_wrap.astype(np.int64)}

    def ref_segment(x, mask):
        f = RU.filter(x * 2 ** 10, (20, 450), butterworth_order=4, btype="bandpass")   # load.py:96
        return RU.rms(f)[mask]                                                          # load.py:98,100

    out["emg_wrap"] = np.stack([ref_segment(x.copy(), mask_wrap) for x in raw])
    out["emg_full"] = np.stack([ref_segment(x.copy(), mask_full) for x in raw])
    out["emg_wrap_f64"] = ref_segment(raw[0].astype(np.float64), mask_wrap)
    assert out["emg_wrap"].dtype == np.float32 and out["emg_wrap_f64"].dtype == np.float64
    with tempfile.TemporaryDirectory() as d:
        for complete in (False, True):
            st = RU.RunningStats(d + "/emg_", complete=complete)
            for w in out["emg_wrap"][:5]:                                  # the "training subset": 5 of 6 segments
                st.push(RU.torchize(w))
            mean, std = st.mean_std()
            tag = "complete" if complete else "perch"
            out[f"stats_mean_{tag}"] = np.asarray(mean.cpu().numpy())
            out[f"stats_std_{tag}"] = np.asarray(std.cpu().numpy())
            out[f"normalized_{tag}"] = st.normalize(RU.torchize(out["emg_wrap"])).cpu().numpy()
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print("preprocess.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "preprocess":
        gen_preprocess()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "prediction":
        gen_prediction()
        sys.exit(0)
    gen_dataset()
    gen_model()
    gen_dropout()
    gen_preprocess()
    gen_prediction()
    copy_ref_artifacts()
