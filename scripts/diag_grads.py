"""Diagnostic: per-tensor gradient error of the CUDA encoder vs the fp32 and fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_encoder import _grads_cuda, _grads_oracle
from gpu_util import perturbed_state, rel_err

for adabn, n in [(True, 328), (True, 4100), (False, 777)]:
    sd = perturbed_state(11, adabn)
    g = torch.Generator().manual_seed(n + 1)
    x, d_emb = torch.randn(n, 12, generator=g), torch.randn(n, 16, generator=g)
    _, got = _grads_cuda(sd, adabn, x, d_emb)
    _, r32 = _grads_oracle(sd, adabn, x, d_emb, torch.float32)
    _, r64 = _grads_oracle(sd, adabn, x, d_emb, torch.float64)
    print(f"--- adabn={adabn} n={n}:  tensor | cuda-vs-fp64 | oracle32-vs-fp64 | cuda-vs-oracle32 | |g|")
    for k in r64:
        print(f"{k:36s} {rel_err(got[k], r64[k]):.2e} {rel_err(r32[k], r64[k]):.2e} {rel_err(got[k], r32[k]):.2e} {float(r64[k].norm()):.3e}")
