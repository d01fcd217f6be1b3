"""Extra bench legs of SURVEY.md section 8(d):

  torch_eager_gpu   the reference's network written with stock torch.nn modules (cuDNN conv / BN, cuBLAS
                    sgemm, foreach Adam) on the same B200 -- the practical bar a PyTorch user starts from.
                    Stand-alone restatement of models.py:248-315 / 112-130 (no import from oracle/): the loss is
                    the VECTORISED form of the reference's per-group Python loop (models.py:132-173), i.e. this
                    bar is faster than the reference's own code, which launches 3*B kernels + B host syncs.
  hbm_kernels       the HBM-bound integer / byte kernels at sizes beyond the 126 MB L2: K1 gather + normalise on a
                    x64-replicated row table, K4' rank + subset evaluation on x64 test windows; achieved GB/s of
                    ALGORITHMIC bytes against the measured copy bandwidth.
"""
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

T = 41


def _events(fn, reps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


class _AdaBN1d(nn.BatchNorm1d):
    def __init__(self, f):
        super().__init__(f, momentum=0, track_running_stats=False)


class _AdaBN2d(nn.BatchNorm2d):
    def __init__(self, f):
        super().__init__(f, momentum=0, track_running_stats=False)


def _torch_emg_net(dp):
    layers = [nn.Conv2d(1, 64, 3, padding=1), nn.ReLU(), _AdaBN2d(64),
              nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(), _AdaBN2d(64), nn.Flatten()]
    k = 768
    for l in range(7):
        layers += [nn.Linear(k, 512), nn.ReLU(), _AdaBN1d(512)]
        if l >= 3:
            layers.append(nn.Dropout(dp))
        k = 512
    layers.append(nn.Linear(512, 16, bias=False))
    return nn.Sequential(*layers)


def torch_eager_gpu(dev, B=4096, steps=5, warmup=2):
    """One train step = forward, symmetric CE, l2 (un-squared norms), backward, Adam x2, on B x 41 windows."""
    torch.manual_seed(42)
    emg = _torch_emg_net(0.5).to(dev)
    table = nn.Linear(T, 16).to(dev)
    opt_e = torch.optim.Adam(emg.parameters(), lr=1e-3)
    opt_g = torch.optim.Adam(table.parameters(), lr=1e-3)
    eye = torch.eye(T, device=dev)
    target = torch.arange(T, device=dev).repeat(B)
    x = torch.randn(B * T, 1, 1, 12, device=dev)

    def step():
        e = F.normalize(emg(x).reshape(B, T, 16), dim=-1)
        g = F.normalize(table(eye), dim=-1)
        logits = e @ g.t()                                                   # (B, 41, 41)
        loss = 0.5 * (F.cross_entropy(logits.reshape(-1, T), target) +
                      F.cross_entropy(logits.transpose(1, 2).reshape(-1, T), target))
        reg = sum(p.norm() for n, p in emg.named_parameters() if p.dim() > 1) + table.weight.norm()
        opt_e.zero_grad(set_to_none=True)
        opt_g.zero_grad(set_to_none=True)
        (loss + 1e-5 * reg).backward()
        opt_e.step()
        opt_g.step()

    ms = _events(step, steps, warmup)
    out = {"ms_per_step": ms, "windows_per_s": B * T / (ms / 1e3),
           "what": f"stock torch.nn eager (cuDNN conv/BN, cuBLAS sgemm fp32, foreach Adam), batch_size {B} groups, "
                   "vectorised loss; torch defaults (matmul fp32 'highest', cuDNN conv may use TF32)"}
    del emg, table, opt_e, opt_g, x
    torch.cuda.empty_cache()
    return out


def eval_pipeline(dev, batch_groups=64, reps=3):
    """The --test pass of train.py:27-44 on the DB2-shaped test split (160 items x 41 classes x 25-sample voting
    windows = 164,000 sEMG windows): gather -> encoder (inference) -> head/logits -> loss + argmax -> windowed vote,
    test batch = 8 x batch_size 8 = 64 items like the reference.  Returns windows/s and voted decisions/s."""
    from contrastiveprosthetics_b200.load import DB23
    from contrastiveprosthetics_b200.models import Model
    from contrastiveprosthetics_b200.utils import TaskWrapper
    params = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.5, 'reg_emg': 1e-5, 'reg_glove': 1e-5}
    torch.manual_seed(42)
    model = Model(dict(params), adabn=True, device=str(dev))
    ds = DB23(db2=True, device=dev)
    ds.load_synthetic(with_glove=False)
    tw = TaskWrapper(ds, with_glove=False)
    tw.set_test()
    model.set_test()
    sink = {}

    def one_pass():
        model.reset()
        losses = []
        with torch.no_grad():
            for (EMG, GLOVE, label) in tw.batches(batch_groups, shuffle=True):
                label = label.reshape(-1)
                losses.append(model.loss(model.forward(EMG, GLOVE, label), label))
        sink["loss"] = torch.stack(losses).mean().item()
        sink["acc"] = model.correct()                      # resolves the vote counts on the host

    ms = _events(one_pass, reps, 1)
    windows = tw.D * T * 25
    return {"ms_per_pass": ms, "windows": windows, "windows_per_s": windows / (ms / 1e3),
            "voted_decisions_per_s": tw.D * T / (ms / 1e3),
            "workload": f"--test pass: {tw.D} items x 41 x 25 windows, test batch {batch_groups} items, AdaBN "
                        "(batch statistics in eval, models.py:17-35), logits materialised, vote over 249 window lengths"}


def preprocess_leg(dev):
    """offline preprocessing of the whole dataset shape (46 x 41 x 6 segments of 2010 x 12
    float32 samples = 1.09 GB) on the GPU, next to the reference's scipy calls (utils.py:134-156) on a sample."""
    import time
    import numpy as np
    from scipy import signal
    from scipy.ndimage import uniform_filter1d
    from contrastiveprosthetics_b200 import preprocess as PP
    raw = PP.synthetic_raw(people=46, seed=0, device=dev)
    n_seg = raw.numel() // (PP.SEG_LEN * 12)
    out = {"segments": n_seg, "raw_gb": raw.numel() * 4 / 1e9}
    for name, wrap in (("reference_time_mask_uint8_wrap", True), ("full_2000_samples", False)):
        out[name + "_ms"] = _events(lambda: PP.preprocess_segments(raw, wrap=wrap), 3, 1)
    b, a = PP.butter_bandpass()
    sample = raw[0, :8].reshape(-1, PP.SEG_LEN, 12).cpu().numpy()         # 48 segments
    t0 = time.perf_counter()
    for x in sample:
        x = x * 2 ** 10
        for c in range(12):
            f = signal.lfilter(b, a, x[:, c]).astype(np.float32)
            np.sqrt(uniform_filter1d(np.square(f), size=11, mode="nearest"))[5:-5]
    cpu_ms = (time.perf_counter() - t0) * 1e3 / len(sample)
    out["cpu_scipy_ms_per_segment"] = cpu_ms
    out["cpu_scipy_ms_whole_dataset_extrapolated"] = cpu_ms * n_seg
    out["note"] = ("one thread per (segment, channel), sequential in time, bit-exact with scipy's lfilter / "
                   "uniform_filter1d loops (un-fused fp64); bound by the 2010-step fp64 recurrence, not by HBM")
    del raw
    torch.cuda.empty_cache()
    return out


def hbm_kernels(dev, hbm_gbs, scale=64):
    """K1 and K4' at x`scale` of the native DB2-shaped sizes, so that inputs exceed the L2."""
    from contrastiveprosthetics_b200 import subset as cps
    from contrastiveprosthetics_b200.utils import gather_rows
    out = {}
    # ---- K1: table of 41 * 20,000 * scale rows x 12 fp32 (2.5 GB at x64); one batch = 4096 * 41 * 16 rows
    rows = 41 * 20000 * scale
    table = torch.randn(rows, 12, device=dev)
    mean = torch.zeros(12, device=dev)
    std = torch.ones(12, device=dev)
    n = 4096 * T * 16
    idx = torch.randint(0, rows, (n,), device=dev)
    ms = _events(lambda: gather_rows(table, idx, mean, std, 12), 10)
    gbs = n * 104 / (ms * 1e-3) / 1e9
    out["gather_norm"] = {"ms": ms, "rows": n, "algorithmic_bytes_per_row": 104, "achieved_gbs": gbs,
                          "frac_of_hbm_peak": gbs / hbm_gbs,
                          "note": f"random 48-byte rows of a {rows * 48 / 1e9:.1f} GB table: every row costs a 64-byte "
                                  "read (2 sectors) + 8-byte index + 48-byte write, so 104 algorithmic bytes move 120"}
    del table, idx
    # ---- K4': 160 * scale items x 25 x 41 rows of 41 logits (1.7 GB at x64), 144 trials (one subset size)
    items, W = 160 * scale, 25
    lg = torch.randn(items * W, T, T, device=dev)
    masks, _ = cps.make_trials(sizes=[20], trials_per_size=144, seed=0)
    mdev = torch.from_numpy(masks).to(dev)
    ev = {}

    def rank():
        ev["e"] = cps.SubsetEvaluator(lg, W)

    ms_rank = _events(rank, 5)
    ms_eval = _events(lambda: ev["e"].evaluate(mdev), 5)
    n_rows = items * W * T
    gbs_rank = n_rows * (164 + 41) / (ms_rank * 1e-3) / 1e9
    out["subset_rank_rows"] = {"ms": ms_rank, "rows": n_rows, "algorithmic_bytes_per_row": 205,
                               "achieved_gbs": gbs_rank, "frac_of_hbm_peak": gbs_rank / hbm_gbs,
                               "note": "164 B of logits in, 41 B of rank order out per (item, sample, class) row"}
    out["subset_eval"] = {"ms": ms_eval, "trials": int(masks.shape[0]),
                          "preds_per_s": n_rows * masks.shape[0] / (ms_eval * 1e-3),
                          "order_table_gbs": n_rows * 41 / (ms_eval * 1e-3) / 1e9,
                          "note": "the rank-order table (41 B per row) is read from HBM once per launch and staged in "
                                  "shared memory per item; one warp per (item, trial) then probes it there, one lane per window (~42 byte "
                                  "probes per (window, trial); vote = match.any + redux.max), so the kernel is bound by "
                                  "instruction issue, not HBM"}
    del lg
    torch.cuda.empty_cache()
    return out
