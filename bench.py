#!/usr/bin/env python
"""bench.py -- train sEMG windows/s (+ subset-eval preds/s) of the B200 hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own code on the host CPU

A "step" is one full train_loop iteration (train.py:95-108) on one batch of synthetic DB2-shaped
data: gather -> encoder forward -> fused head/loss (+ l2) -> backward -> [gradient all-reduce]
-> Adam x2.  Workload at every N: config C2 of BASELINE.json per GPU -- batch_size 4096 groups =
167,936 windows per step per GPU, AdaBN on, fp32 (weak scaling; BatchNorm statistics stay local
to the rank).  One JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T = 41
PARAMS = {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.5, 'reg_emg': 1e-5, 'reg_glove': 1e-5,
          'lr_emg': 1e-3, 'lr_glove': 1e-3, 'epochs': 1}
FLOP_PER_WINDOW_TRAIN = 2 * 6_369_792          # fwd 2.124 M MAC, train 3x minus conv1 dX


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class stdout_to_stderr:
    """fd-level redirect: NCCL prints its version banner to stdout when the first communicator is created; the
    driver wants exactly ONE JSON line there."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


def init_distributed():
    import torch.distributed as dist
    from contrastiveprosthetics_b200 import dist as cpdist
    with stdout_to_stderr():
        rank, world, dev = cpdist.init_from_env()
        if world > 1:
            dist.barrier()                 # creates the NCCL communicator (and its banner) now
            torch.cuda.synchronize()
    return rank, world, dev


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (NVML, 20 ms period: the default
    timed region is ~0.1 s)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- CPU oracle arm
def cpu_oracle_steps(n_steps, warmup, groups=256, seed=0):
    """The reference's CPU path as restated by oracle/ (torch CPU, all host threads): forward, loss,
    l2, backward, two Adams on `groups` x 41 windows per step.  Returns (windows/s, ms/step, cores)."""
    from oracle import model as OM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = OM.init_state(42, True)
    g = torch.Generator().manual_seed(seed)
    tk = OM.trainable_keys(sd)
    m = {k: torch.zeros_like(sd[k]) for k in tk}
    v = {k: torch.zeros_like(sd[k]) for k in tk}
    N = groups * T
    times = []
    for s in range(warmup + n_steps):
        EMG = torch.randn(groups, T, 1, 1, 12, generator=g)
        masks = [torch.empty(N, 512).bernoulli_(0.5, generator=g) for _ in range(4)]
        t0 = time.perf_counter()
        res, grads, _ = OM.train_step_grads(sd, EMG, True, dp=0.5, dropout_masks=masks, reg_emg=1e-5, reg_glove=1e-5)
        for k in tk:
            OM.adam_update(sd[k], grads[k], m[k], v[k], s + 1, 1e-3)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    return N / (ms / 1e3), ms, torch.get_num_threads()


CPU_SAMPLE_GROUPS = 256          # ONE bounded sample of the workload for every CPU number of both arms


def cpu_reference_baseline(n_steps, warmup):
    """cpu_baseline object: the reference's OWN code (oracle/_ref, staged by oracle/ref_fetch.py) timed on this
    box's host cores on CPU_SAMPLE_GROUPS groups per step, with the oracle port (oracle/model.py) beside it on
    the same sample; falls back to the port alone when the staged reference is absent."""
    from oracle import ref_fetch
    groups = CPU_SAMPLE_GROUPS
    port_wps, port_ms, cores = cpu_oracle_steps(n_steps, warmup, groups)
    sample = f"{n_steps} steps x {groups * T} windows ({groups} groups) of the C2 train step, same per-window work"
    if ref_fetch.staged():
        from oracle import refrun
        wps, ms, cores = refrun.train_steps(n_steps, warmup, groups, PARAMS)
        return {"value": wps, "unit": "windows/s", "cores": cores, "kind": "reference", "ms_per_step": ms,
                "sample": sample + "; the UNMODIFIED reference (models.py / utils.py / load.py via oracle/_ref, "
                          "'cuda' -> 'cpu', stub imports) running train.py:83-108's loop body: DataLoader over its "
                          "TaskWrapper, Model.forward / loss / l2, backward, Adam x2",
                "port": {"value": port_wps, "unit": "windows/s", "ms_per_step": port_ms,
                         "what": "oracle/model.py (vectorised torch-CPU restatement) on the same sample"}}
    return {"value": port_wps, "unit": "windows/s", "cores": cores, "kind": "port", "ms_per_step": port_ms,
            "sample": sample + "; oracle/ torch-CPU port of train.py:95-108 (oracle/_ref not staged)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = cpu_reference_baseline(args.steps, args.warmup)
    wps, ms = cpu["value"], cpu["ms_per_step"]
    line = {
        "impl": "reference", "metric": "train sEMG windows/s", "value": wps, "unit": "windows/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: train step, batch_size 4096 groups x 41 windows, AdaBN on, fp32",
                   "note": f"CPU arm times a bounded sample of the workload: {CPU_SAMPLE_GROUPS} groups "
                           f"({CPU_SAMPLE_GROUPS * T} windows) per step, same per-window work"},
        "cpu_baseline": cpu,
        "e2e": {"value": wps, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    import torch.distributed as dist
    from contrastiveprosthetics_b200 import _lib, dist as cpdist, subset as cps
    from contrastiveprosthetics_b200.load import DB23
    from contrastiveprosthetics_b200.models import Model
    from contrastiveprosthetics_b200.utils import TaskWrapper

    rank, world, dev = init_distributed()
    assert dev.type == "cuda", "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    L = _lib.lib()
    # weak scaling (default): --batch_size groups per GPU; strong scaling: --global_batch groups split over the ranks
    B = args.global_batch // world if args.global_batch else args.batch_size
    strong = bool(args.global_batch)
    N = B * T
    torch.manual_seed(42)
    model = Model(dict(PARAMS), adabn=True, device=str(dev))
    model.emg_net.engine = {"tc": _lib.ENGINE_TC, "simt": _lib.ENGINE_SIMT, "tc_fp16": _lib.ENGINE_TC_FP16}[args.engine]
    model.emg_net.sync_bn = args.sync_bn
    # train.py:72-73's two Adams, in torch's single-kernel implementation (fused=True) at every N -- same update rule,
    # 2 launches instead of ~14 multi-tensor launches per step
    opt_e = torch.optim.Adam(model.emg_net.parameters(), lr=PARAMS['lr_emg'], weight_decay=0, fused=True)
    opt_g = torch.optim.Adam(model.glove_net.parameters(), lr=PARAMS['lr_glove'], weight_decay=0, fused=True)
    sync_grads = cpdist.FlatGradAllReduce(list(model.emg_net.parameters()) + list(model.glove_net.parameters()))
    # N = 1: DB2-shaped (config C2).  N > 1: DB2 + DB3 subjects mixed, the 6 DB3 subjects 11-channel (config C3)
    mixed = world > 1 or args.mixed
    ds = DB23(db2=True, device=dev, mixed=mixed)
    ds.load_synthetic(with_glove=False)
    tw = TaskWrapper(ds, with_glove=False)
    tw.set_train()
    model.set_train()
    gen = torch.Generator().manual_seed(1234)
    gen_rank = torch.Generator().manual_seed(4321 + rank)

    def step_resident(items):
        EMG, GLOVE, label = tw.get_batch(items)
        label = label.reshape(-1)
        logits = model.forward(EMG, GLOVE, label)
        loss = model.loss(logits, label) + model.l2()
        opt_e.zero_grad(set_to_none=True)
        opt_g.zero_grad(set_to_none=True)
        loss.backward()
        sync_grads()
        opt_e.step()
        opt_g.step()
        return loss

    def draw_items():
        # sample sharding: every rank draws the same global batch and takes its slice.  When the global batch
        # exceeds the D items of one epoch (mixed subjects: D = 13,800 < 8 x 4096) every rank draws its own
        # B items instead, so the per-GPU batch stays B (weak scaling)
        if B * world <= tw.D:
            order = torch.randperm(tw.D, generator=gen)[:B * world]
            return order[rank * B:(rank + 1) * B].to(dev)
        return torch.randperm(tw.D, generator=gen_rank)[:B].to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.cp_launch_count()
        with ClockSampler(dev.index or 0) as cs:
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1) / steps
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), cs.summary(), L.cp_launch_count() - l0

    # ---- device-resident metric
    items_ring = [draw_items() for _ in range(8)]
    it = {"i": 0}

    def resident_step():
        it["i"] += 1
        step_resident(items_ring[it["i"] % len(items_ring)])

    ms, clocks, launches = timed(resident_step, args.steps, args.warmup)
    value = N * world / (ms / 1e3)
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_only": True, "value": value, "ms_per_step": ms, "gpu_launches": int(launches)}))
        return

    # ---- e2e: host batches (pinned) -> H2D every step, loss + correct counts -> host every step
    host_batches = []
    for i in range(4):
        EMG, _, label = tw.get_batch(items_ring[i])
        host_batches.append((EMG.cpu().pin_memory(), label.cpu().pin_memory()))
    h2d = host_batches[0][0].numel() * 4 + host_batches[0][1].numel() * 8
    sink = {"loss": 0.0}

    class HostFeed:
        """Double-buffered input feed: the pinned-host -> device copy of the NEXT step's batch is issued on a copy stream
        right after this step's kernels are enqueued, so it overlaps the step's compute; the step that consumes a batch
        waits on the copy's event.  One H2D copy of a full batch per step, all of them inside the timed region (the
        copy a step consumes was issued by the step before it; the last timed step issues the one after)."""

        def __init__(self):
            self.stream = torch.cuda.Stream(device=dev)
            self.pending = None
            self.i = 0

        def _issue(self):
            hE, hl = host_batches[self.i % len(host_batches)]
            self.i += 1
            with torch.cuda.stream(self.stream):
                E = hE.to(dev, non_blocking=True)
                lab = hl.to(dev, non_blocking=True)
                ev = self.stream.record_event()
            return E, lab, ev

        def get(self):
            if self.pending is None:
                self.pending = self._issue()
            E, lab, ev = self.pending
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            E.record_stream(cur)
            lab.record_stream(cur)
            return E, lab

        def prefetch(self):
            self.pending = self._issue()

    feed = HostFeed()

    def e2e_step():
        it["i"] += 1
        EMG, label = feed.get()
        label = label.reshape(-1)
        logits = model.forward(EMG, None, label)
        loss = model.loss(logits, label)
        total = loss + model.l2()
        opt_e.zero_grad(set_to_none=True)
        opt_g.zero_grad(set_to_none=True)
        total.backward()
        sync_grads()
        opt_e.step()
        opt_g.step()
        feed.prefetch()                                  # next step's H2D copy, behind this step's kernels on its own stream
        sink["loss"] = loss.item()                       # device -> host read of the step's result
        sink["acc"] = model.corrects[-1]                 # per-group correct counts (B int32) -> host

    ms_e2e, _, _ = timed(e2e_step, max(3, args.steps // 2), 2)
    d2h = 4 + B * 4
    modes = {"eager": {"ms_per_step": ms, "e2e_ms_per_step": ms_e2e}}
    step_mode = e2e_mode = "eager"

    # ---- same step captured once as a CUDA graph (bit-identical to the eager step, tests/test_gpu_graph.py): one
    #      graph launch per step instead of ~400 kernel launches; at N > 1 the NCCL gradient all-reduce (and the SyncBN
    #      all-reduces) are captured inside the graph
    if not args.no_graph and not strong and (world == 1 or dist.get_backend() == "nccl"):
        from contrastiveprosthetics_b200.graph import GraphedTrainStep
        torch.manual_seed(42)
        model_g = Model(dict(PARAMS), adabn=True, device=str(dev))
        model_g.emg_net.engine = model.emg_net.engine
        model_g.emg_net.sync_bn = args.sync_bn
        model_g.set_train()
        # same optimizer, torch's single-kernel implementation (fused=True): 2 graph nodes instead of ~14
        opts_g = [torch.optim.Adam(model_g.emg_net.parameters(), lr=PARAMS['lr_emg'], weight_decay=0, capturable=True,
                                   fused=True),
                  torch.optim.Adam(model_g.glove_net.parameters(), lr=PARAMS['lr_glove'], weight_decay=0, capturable=True,
                                   fused=True)]
        sync_g = cpdist.FlatGradAllReduce(list(model_g.emg_net.parameters()) + list(model_g.glove_net.parameters())) \
            if world > 1 else None
        gstep = GraphedTrainStep(model_g, opts_g, tw.get_batch(items_ring[0])[0], sync_grads=sync_g)

        def graph_resident():
            it["i"] += 1
            gstep(tw.get_batch(items_ring[it["i"] % len(items_ring)])[0])

        def graph_e2e():
            it["i"] += 1
            EMG, _ = feed.get()
            loss, ncor = gstep(EMG)
            feed.prefetch()
            sink["loss"] = loss.item()
            sink["acc"] = ncor.cpu()

        ms_g, clocks_g, _ = timed(graph_resident, args.steps, args.warmup)
        ms_g_e2e, _, _ = timed(graph_e2e, max(3, args.steps // 2), 2)
        modes["cuda_graph"] = {"ms_per_step": ms_g, "e2e_ms_per_step": ms_g_e2e}
        graph_best = ("cuda_graph", ms_g, clocks_g, ms_g_e2e)
        del gstep, model_g, opts_g, sync_g
        # ---- the same step WITHOUT autograd / torch.optim inside the graph (step.LeanTrainStep: one prologue launch,
        #      one Adam launch for both optimizers, the gradient bucket all-reduced in place): ~35 fewer nodes per step
        torch.manual_seed(42)
        model_g = Model(dict(PARAMS), adabn=True, device=str(dev))
        model_g.emg_net.engine = model.emg_net.engine
        model_g.emg_net.sync_bn = args.sync_bn
        model_g.set_train()
        opts_g = [torch.optim.Adam(model_g.emg_net.parameters(), lr=PARAMS['lr_emg'], weight_decay=0),
                  torch.optim.Adam(model_g.glove_net.parameters(), lr=PARAMS['lr_glove'], weight_decay=0)]
        sync_g = True if world > 1 else None
        gstep = GraphedTrainStep(model_g, opts_g, tw.get_batch(items_ring[0])[0], sync_grads=sync_g, lean=True)
        ms_l, clocks_l, _ = timed(graph_resident, args.steps, args.warmup)
        ms_l_e2e, _, _ = timed(graph_e2e, max(3, args.steps // 2), 2)
        modes["cuda_graph_lean"] = {"ms_per_step": ms_l, "e2e_ms_per_step": ms_l_e2e}
        if ms_l < ms_g:
            graph_best = ("cuda_graph_lean", ms_l, clocks_l, graph_best[3])
        if ms_l_e2e < ms_g_e2e:
            graph_best = graph_best[:3] + (ms_l_e2e,)
            e2e_graph_name = "cuda_graph_lean"
        else:
            e2e_graph_name = "cuda_graph"
        gname, ms_g, clocks_g, ms_g_e2e = graph_best
        # every rank must take the same branch: the timings are already the max over ranks
        if ms_g < ms:
            step_mode, ms, clocks = gname, ms_g, clocks_g
            value = N * world / (ms / 1e3)
        if ms_g_e2e < ms_e2e:
            e2e_mode, ms_e2e = e2e_graph_name, ms_g_e2e
        del gstep, model_g, opts_g, sync_g
    e2e_value = N * world / (ms_e2e / 1e3)

    # ---- dominant kernel alone: fc forward GEMM at the step's shape (CUDA events on its stream)
    roof = None
    sub = None
    cpu = None
    if rank == 0:
        peaks = measured_peaks()
        M, Nn, K = N, 512, 512
        A = torch.randn(M, K, device=dev)
        Wt = torch.randn(Nn, K, device=dev)
        bias = torch.randn(Nn, device=dev)
        Y = torch.empty(M, Nn, device=dev)
        nb = L.cp_linear_workspace_bytes(M, Nn, K)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        P = _lib.ptr

        use_tc = model.emg_net.engine != _lib.ENGINE_SIMT
        one_product = model.emg_net.engine == _lib.ENGINE_TC_FP16
        if use_tc:      # operands pre-split, as they are inside the encoder (the producing kernels write planes)
            Ah, Al, Wh, Wl = (torch.empty_like(A, dtype=torch.float16), torch.empty_like(A, dtype=torch.float16),
                              torch.empty_like(Wt, dtype=torch.float16), torch.empty_like(Wt, dtype=torch.float16))
            _lib.check(L.cp_split_planes(P(A), P(Ah), P(Al), A.numel(), _lib.stream()))
            _lib.check(L.cp_split_planes(P(Wt), P(Wh), P(Wl), Wt.numel(), _lib.stream()))

        def gemm():
            if use_tc:
                _lib.check(L.cp_linear_forward_planes(P(Ah), None if one_product else P(Al), P(Wh), None if one_product else P(Wl),
                                                      P(bias), P(Y), M, Nn, K, 1, None, None, P(ws), nb, _lib.stream()))
            else:
                _lib.check(L.cp_linear_forward(P(A), P(Wt), P(bias), P(Y), M, Nn, K, 1, None, None, P(ws), nb,
                                               _lib.ENGINE_SIMT, _lib.stream()))
        for _ in range(3):
            gemm()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            gemm()
        e1.record()
        torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / reps
        ach = 2.0 * M * Nn * K / (gms * 1e-3) / 1e12
        kname = ("gemm_tc_nt_pair_kernel (tcgen05 kind::f16 cta_group::2, ONE fp16 product, TMA, TMEM)" if one_product else
                 "gemm_tc_nt_pair_kernel (tcgen05 kind::f16 cta_group::2, 3-product fp16 split, TMA, TMEM)" if use_tc
                 else "gemm_nt_kernel<128,128> (fp32 FFMA)")
        roof = {"kernel": kname + ": Linear 512->512 + bias + ReLU + BN-stat partials at the step's shape "
                          "(M = 167,936); 7 fwd + 7 dgrad + 7 wgrad launches of this family per step",
                "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"],
                # the fp32-parity split issues 3 tensor-core products per algorithmic product: its own ceiling
                "frac_of_parity_ceiling": (ach / (peaks["bf16_tflops"] / 3.0)) if (use_tc and not one_product) else None,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel at this shape, from the
                # ncu --set full capture profiles/r1_ncu_gemm_nt_pair_f16_summary.txt (algorithmic: 344 MB of fp16
                # operand planes in + 344 MB of fp32 out = 688 MB; the weights stay in L2)
                "traffic": 654.8e6 if use_tc and not one_product else None, "traffic_unit": "bytes/launch",
                "peak_source": peaks["source"],
                "ms_per_launch": gms,
                "note": "achieved = algorithmic fp32 FLOPs (2MNK); the fp32-parity path issues 3 fp16 tensor-core "
                        "products per algorithmic product (x = hi + lo/2048), so its ceiling is 1/3 of the "
                        "bf16/fp16 peak; step-level: "
                        f"{FLOP_PER_WINDOW_TRAIN * N / (ms * 1e-3) / 1e12:.2f} TFLOP/s algorithmic"}
        del A, Wt, Y, ws

    # ---- subset evaluator (C4): 160 test items x 25 x 41 windows, 144 trials x 40 sizes, trials sharded
    items_eval, Wv = 160, 25
    lg = torch.randn(items_eval * Wv, T, T, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    masks, sizes = cps.make_trials(sizes=range(1, 41), trials_per_size=144, seed=0)
    lo, hi = cps.shard_trials(len(masks), rank, world)
    mdev = torch.from_numpy(masks[lo:hi]).to(dev)

    def subset_run():
        ev = cps.SubsetEvaluator(lg, Wv)
        return ev.evaluate(mdev)

    for _ in range(2):
        subset_run()
    ms_sub, _, _ = timed(subset_run, 5, 1)
    preds_per_s = items_eval * Wv * T * len(masks) / (ms_sub / 1e3)

    c1 = eager = hbm = evalp = prep = None
    if rank == 0 and world == 1 and not strong:
        # the reference's own launch-bound configuration (go.sh:6): eager vs CUDA-graph step
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        from bench_c1 import c1_concurrent_folds, c1_small_batch
        from bench_extra import eval_pipeline, hbm_kernels, preprocess_leg, torch_eager_gpu
        c1 = c1_small_batch(dev)
        c1["concurrent_folds"] = c1_concurrent_folds(dev)
        eager = torch_eager_gpu(dev, B)              # stock PyTorch on the same GPU, the practical bar
        hbm = hbm_kernels(dev, measured_peaks()["hbm_gbs"])
        evalp = eval_pipeline(dev)
        prep = preprocess_leg(dev)
    c5 = None
    if not strong:
        torch.cuda.empty_cache()
        c5 = c5_leg(rank, world, dev)            # config 5 at this N (every rank takes part in its collectives)
    if rank == 0:
        if world == 1:       # N = 1 only: at N > 1 the other ranks spin in the barrier on the same host cores
            cpu = cpu_reference_baseline(n_steps=10, warmup=2)
        line = {
            "metric": "train sEMG windows/s", "value": value, "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("C3: sample-sharded train step, " if world > 1 else "C2: train step, ") +
                                   (f"global batch {args.global_batch} groups, " if strong else "") +
                                   f"batch_size {B} groups x 41 windows = {N} windows per GPU per step, AdaBN on, "
                                   "dropout 0.5, fp32, " +
                                   ("DB2+DB3 mixed-subject synthetic sEMG (46 subjects, DB3 subjects 11-channel)"
                                    if mixed else "DB2-shaped synthetic sEMG"),
                       "batch_size_groups_per_gpu": B, "windows_per_step": N * world,
                       "engine": {0: "simt-fp32", 1: "tcgen05-3xfp16-split", 2: "tcgen05-1xfp16 (reduced precision, 1e-2 path)"}[model.emg_net.engine],
                       "step_mode": step_mode + (" (one graph launch per step, torch.optim.Adam(fused=True); gpu_launches "
                                                 "counts the kernels of the eager step, the graph replays the same ones "
                                                 "of this library)" if step_mode == "cuda_graph" else
                                                 " (one graph launch per step; the step runs without autograd / torch.optim: "
                                                 "cp_step_prologue + the fused kernels + cp_adam_step, gradients in one flat "
                                                 "bucket; gpu_launches counts the kernels of the eager autograd step)"
                                                 if step_mode == "cuda_graph_lean" else ""),
                       "parallelism": f"dp{world} (sample-sharded, " + ("SyncBN" if args.sync_bn and world > 1 else "local BatchNorm")
                                      + ", one flat grad all-reduce)",
                       "optimizer": ("both Adams of train.py:72-73 in ONE cp_adam_step launch (torch.optim.Adam's update rule)"
                                     if step_mode == "cuda_graph_lean" else
                                     "2 x torch.optim.Adam(fused=True) (train.py:72-73 wiring)"),
                       "l2_policy": "per-step working set ~8 GB of activations >> 126 MB L2; no explicit flush"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e, "step_mode": e2e_mode,
                    "input_feed": "pinned host batch -> device on a copy stream, double-buffered (the copy of step i+1 "
                                  "overlaps the compute of step i; one full-batch copy per step inside the timed region); "
                                  "loss and per-group counts read back to the host every step"},
            "roofline": roof, "cpu_baseline": cpu, "step_modes": modes,
            "c1_small_batch": c1, "torch_eager_gpu": eager, "hbm_kernels": hbm, "eval_pipeline": evalp, "offline_preprocess": prep,
            "c5_clip": c5,
            "subset_eval": {"value": preds_per_s, "unit": "preds/s", "ms": ms_sub,
                            "workload": "C4: 160 items x 25 x 41 test windows x 5760 trials (144 x 40 sizes), "
                                        "rank + vote + count; trials sharded over ranks"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- config 5 (CLIP) arm
def c5_leg(rank, world, dev, B=65536, steps=5, warmup=2):
    """Config 5 inside the default line (`c5_clip` key, every N): the batch x batch CLIP train step at GLOBAL batch B
    sharded over the ranks (strong scaling), and the head alone on this rank's (B/world) x B strip.  Device-resident
    inputs, CUDA events, max over ranks."""
    import torch.distributed as dist
    from contrastiveprosthetics_b200 import clip as C, dist as cpdist
    from contrastiveprosthetics_b200.clip import ClipModel
    from contrastiveprosthetics_b200.load import DB23
    from contrastiveprosthetics_b200.utils import TaskWrapper
    n = B // world
    torch.manual_seed(42)
    model = ClipModel(dict(PARAMS), glove_dim=22, device=str(dev))
    model.train()
    opt_e = torch.optim.Adam(model.emg_net.parameters(), lr=PARAMS['lr_emg'], weight_decay=0)
    opt_g = torch.optim.Adam(model.glove_net.parameters(), lr=PARAMS['lr_glove'], weight_decay=0)
    sync_grads = cpdist.FlatGradAllReduce(list(model.emg_net.parameters()) + list(model.glove_net.parameters()),
                                          average=False)
    ds = DB23(db2=True, device=dev)
    ds.load_synthetic(with_glove=True, glove_dim=22)
    tw = TaskWrapper(ds, with_glove=True)
    tw.set_train()
    tw.idx = torch.randperm(tw.TASKS * tw.D, generator=torch.Generator().manual_seed(5)).to(dev)   # same on every rank
    n_batches = (tw.TASKS * tw.D) // B
    it = {"i": 0}

    def step():
        it["i"] += 1
        EMG, GLOVE, _ = tw.get_flat_batch((it["i"] % n_batches) * B + rank * n, n)
        e, g = model(EMG, GLOVE)
        total = model.loss(e, g) + model.l2()
        opt_e.zero_grad(set_to_none=True)
        opt_g.zero_grad(set_to_none=True)
        total.backward()
        sync_grads()
        opt_e.step()
        opt_g.step()

    E = torch.randn(n, 16, device=dev, requires_grad=True)
    G = torch.randn(n, 16, device=dev, requires_grad=True)

    def head_only():
        E.grad = G.grad = None
        C.clip_head(E, G, 0.0)[0].backward()

    def timed(fn, k, w):
        for _ in range(w):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / k], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = timed(step, steps, warmup)
    ms_head = timed(head_only, steps, warmup)
    model.n_correct.clear()
    return {"ms_per_step": ms, "samples_per_s": B / (ms / 1e3), "global_batch": B, "samples_per_gpu": n, "scaling": "strong",
            "head_ms": ms_head, "head_pairs_per_s": float(n) * float(B) * world / (ms_head / 1e3),
            "workload": "C5: glove-angle (22-dim) tower + EMG tower, CLIP batch x batch loss (B x B never materialised; "
                        "mma.sync m16n8k16 fp16 3-product split + MUFU sweeps), all-gather Ghat / all-reduce column sums / reduce-scatter dGhat at N > 1; "
                        "`bench.py --workload c5` prints the full line (e2e, clocks)"}


def run_c5(args):
    """BASELINE.json config 5: glove-angle (22-dim) tower + EMG tower, batch x batch CLIP loss, GLOBAL batch
    65,536 sharded over the ranks (strong scaling): all-gather of the glove embeddings, all-reduce of the
    column sums, reduce-scatter of the glove-embedding gradients, one flat parameter-gradient all-reduce."""
    import torch.distributed as dist
    from contrastiveprosthetics_b200 import _lib, dist as cpdist
    from contrastiveprosthetics_b200.clip import ClipModel
    from contrastiveprosthetics_b200.load import DB23
    from contrastiveprosthetics_b200.utils import TaskWrapper

    rank, world, dev = init_distributed()
    assert dev.type == "cuda", "bench.py needs a GPU (no CPU fallback)"
    L = _lib.lib()
    B = args.clip_batch
    n = B // world
    torch.manual_seed(42)
    model = ClipModel(dict(PARAMS), glove_dim=22, device=str(dev))
    model.train()
    opt_e = torch.optim.Adam(model.emg_net.parameters(), lr=PARAMS['lr_emg'], weight_decay=0)
    opt_g = torch.optim.Adam(model.glove_net.parameters(), lr=PARAMS['lr_glove'], weight_decay=0)
    sync_grads = cpdist.FlatGradAllReduce(list(model.emg_net.parameters()) + list(model.glove_net.parameters()),
                                          average=False)
    ds = DB23(db2=True, device=dev)
    ds.load_synthetic(with_glove=True, glove_dim=22)
    tw = TaskWrapper(ds, with_glove=True)
    tw.set_train()
    tw.idx = torch.randperm(tw.TASKS * tw.D, generator=torch.Generator().manual_seed(5)).to(dev)   # same on every rank
    n_batches = (tw.TASKS * tw.D) // B
    it = {"i": 0}

    def train_on(EMG, GLOVE):
        e, g = model(EMG, GLOVE)
        loss = model.loss(e, g)
        total = loss + model.l2()
        opt_e.zero_grad(set_to_none=True)
        opt_g.zero_grad(set_to_none=True)
        total.backward()
        sync_grads()
        opt_e.step()
        opt_g.step()
        return loss

    def resident_step():
        it["i"] += 1
        start = (it["i"] % n_batches) * B + rank * n
        EMG, GLOVE, _ = tw.get_flat_batch(start, n)
        train_on(EMG, GLOVE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        model.n_correct.clear()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.cp_launch_count()
        with ClockSampler(dev.index or 0) as cs:
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), cs.summary(), L.cp_launch_count() - l0

    ms, clocks, launches = timed(resident_step, args.steps, args.warmup)

    host = []
    for i in range(3):
        EMG, GLOVE, _ = tw.get_flat_batch(i * B + rank * n, n)
        host.append((EMG.cpu().pin_memory(), GLOVE.cpu().pin_memory()))
    sink = {}

    def e2e_step():
        it["i"] += 1
        hE, hG = host[it["i"] % len(host)]
        loss = train_on(hE.to(dev, non_blocking=True), hG.to(dev, non_blocking=True))
        sink["loss"] = loss.item()
        sink["ncor"] = int(model.n_correct[-1].item())

    ms_e2e, _, _ = timed(e2e_step, max(3, args.steps // 2), 2)

    # the head alone at this rank's strip (n x B): 2 sums + 2 gradient sweeps; algorithmic fp32 FLOPs =
    # 2*16 per pair for the similarity in each of the 4 sweeps + 2*16 per pair in each gradient sweep
    from contrastiveprosthetics_b200 import clip as C
    E = torch.randn(n, 16, device=dev, requires_grad=True)
    G = torch.randn(n, 16, device=dev, requires_grad=True)

    def head_only():
        E.grad = G.grad = None
        C.clip_head(E, G, 0.0)[0].backward()

    ms_head, _, _ = timed(head_only, 5, 2)
    pairs = float(n) * float(B)
    head_flops = pairs * 32.0 * 6
    if rank == 0:
        line = {
            "metric": "train sEMG windows/s", "value": B / (ms / 1e3), "unit": "windows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C5: glove-angle (22-dim) tower + EMG tower, CLIP batch x batch loss, global batch {B} "
                                   f"({n} samples per GPU), AdaBN (rank-local statistics), dropout 0.5, fp32",
                       "global_batch": B, "parallelism": f"dp{world}: all-gather Ghat, all-reduce column sums, "
                       "reduce-scatter d Ghat, one flat grad all-reduce (sum)",
                       "l2_policy": "activation stream of the towers >> 126 MB L2; no explicit flush"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": B / (ms_e2e / 1e3), "unit": "windows/s",
                    "h2d_bytes_per_step": int(host[0][0].numel() * 4 + host[0][1].numel() * 4),
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e},
            "clip_head": {"ms": ms_head, "pairs_per_s": pairs / (ms_head / 1e3),
                          "fp32_tflops": head_flops / (ms_head * 1e-3) / 1e12,
                          "kernel": "clip_sweep_mma_kernel (mma.sync m16n8k16 fp16, 3-product hi/lo split, fp32 accumulate + MUFU.EX2; B x B never materialised)",
                          "note": "bound by instruction issue / MUFU (one ex2 per pair), not HBM (inputs 8 MB)"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch_size", type=int, default=4096, help="groups of 41 windows per GPU per step")
    ap.add_argument("--global_batch", type=int, default=0,
                    help="strong scaling: this many groups per step split over the ranks (C3: 32768); default: weak "
                         "scaling at --batch_size groups per GPU")
    ap.add_argument("--profile", action="store_true",
                    help="only the device-resident train steps (for ncu launch lists); prints a reduced line")
    ap.add_argument("--engine", default="tc", choices=["tc", "simt", "tc_fp16"],
                    help="tc: tcgen05 GEMMs on the 3-product fp16 split, fp32 parity (default); simt: fp32 FFMA GEMMs; "
                         "tc_fp16: ONE fp16 tensor-core product (TF32-class accuracy; the 1e-2 path, not the headline)")
    ap.add_argument("--no_graph", action="store_true", help="N = 1: skip the CUDA-graph variant of the step")
    ap.add_argument("--sync_bn", action="store_true", help="N > 1: BatchNorm statistics over every rank's rows")
    ap.add_argument("--mixed", action="store_true", help="mixed DB2+DB3 subjects also at N = 1 (default at N > 1)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2: the headline train step (default); c5: glove CLIP batch x batch variant, global batch sharded")
    ap.add_argument("--clip_batch", type=int, default=65536, help="global batch of --workload c5")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
