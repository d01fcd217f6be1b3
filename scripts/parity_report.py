#!/usr/bin/env python
"""Measured worst-case parity errors of the EMG encoder per tensor -> profiles/parity_r2.json (run on the GPU box).

    python scripts/parity_report.py [--out gpurun_out/parity_r2.json] [--full]

Every number is a norm-wise relative error |got - ref| / |ref| of one tensor (embeddings, each stage's post-ReLU
activation, each parameter gradient) between the CUDA path (through the C ABI) and the CPU oracle -- fp32 with the
kernel's ReLU pattern injected, un-conditioned fp32, and (small sizes) float64 beside the fp32 oracle's own distance
to float64."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from contrastiveprosthetics_b200 import _lib  # noqa: E402
from gpu_util import encoder_parity_errors, perturbed_state, scale_state, worst  # noqa: E402


def inputs(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 12, generator=g) + 0.5 * torch.randn(41, 12, generator=g).repeat((n + 40) // 41, 1)[:n]
    return x, torch.randn(n, 16, generator=g) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_r2.json"))
    ap.add_argument("--full", action="store_true", help="include the C2-size case (167,936 windows)")
    ap.add_argument("--only_full", action="store_true", help="only the C2-size case")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 8)
    cases = [("n328_adabn_dp0", 328, True, 0.0, 1.0, 1.0, True), ("n328_stockbn_dp0", 328, False, 0.0, 1.0, 1.0, True),
             ("n4100_adabn_dp0.5", 4100, True, 0.5, 1.0, 1.0, True), ("n20992_adabn_dp0.5", 20992, True, 0.5, 1.0, 1.0, True),
             ("n1640_gamma1e-3", 1640, True, 0.5, 1e-3, 1.0, True), ("n1640_gamma3e2", 1640, True, 0.5, 3e2, 1.0, True),
             ("n1640_weight1e-4", 1640, True, 0.5, 1.0, 1e-4, True), ("n1640_weight1e3", 1640, True, 0.5, 1.0, 1e3, True)]
    if args.only_full:
        cases = []
    if args.full or args.only_full:
        cases.append(("n167936_adabn_dp0.5_C2", 4096 * 41, True, 0.5, 1.0, 1.0, True))
    report = {"what": "norm-wise relative error per tensor, CUDA path vs CPU oracle", "tolerance": 1e-5, "cases": {}}
    for name, n, adabn, dp, gamma, weight, fp64 in cases:
        sd = scale_state(perturbed_state(3, adabn), adabn, gamma, weight)
        x, d_emb = inputs(n, n)
        g = torch.Generator().manual_seed(7)
        masks = [torch.empty(n, 512, dtype=torch.uint8).bernoulli_(0.5, generator=g) for _ in range(4)] if dp > 0 else None
        for eng_name, eng in (("simt", _lib.ENGINE_SIMT), ("tc", _lib.ENGINE_TC)):
            if n > 50000 and eng == _lib.ENGINE_SIMT:
                continue
            t0 = time.time()
            e = encoder_parity_errors(sd, adabn, x, d_emb, eng, dp, masks, fp64=fp64)
            summary = {"emb": e["emb"], "worst_stage": worst(e, "stage"), "worst_grad_relu_fixed": worst(e, "grad|"),
                       "worst_grad_unconditioned": worst(e, "grad_unconditioned|"),
                       "relu_flip_fraction": e["relu_flip_fraction"]}
            if fp64:
                summary["emb_vs_fp64"] = e["emb64"]
                summary["emb_fp32_oracle_vs_fp64"] = e["emb_oracle32_vs_64"]
                summary["worst_grad_vs_fp64"] = worst(e, "grad64|")
                summary["worst_fp32_oracle_vs_fp64"] = worst(e, "oracle32_vs_64|")
            report["cases"][f"{name}|{eng_name}"] = {"summary": summary, "per_tensor": e, "seconds": time.time() - t0}
            print(name, eng_name, json.dumps(summary), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
