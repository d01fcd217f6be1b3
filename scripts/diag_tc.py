"""Diagnostic: forward/backward error of both engines vs the fp64 oracle (ReLU pattern fixed)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_encoder import _grads_cuda, _grads_oracle, _relu_pattern
from gpu_util import perturbed_state, rel_err

for adabn, n in [(True, 41 * 16), (True, 41 * 200), (False, 777)]:
    sd = perturbed_state(11, adabn)
    g = torch.Generator().manual_seed(n + 1)
    x, d_emb = torch.randn(n, 12, generator=g), torch.randn(n, 16, generator=g)
    for engine in (0, 1):
        taps = {}
        emb, got = _grads_cuda(sd, adabn, x, d_emb, taps=taps, engine=engine)
        pat = _relu_pattern(taps)
        e32, r32 = _grads_oracle(sd, adabn, x, d_emb, torch.float32, relu_masks=pat)
        e64, r64 = _grads_oracle(sd, adabn, x, d_emb, torch.float64, relu_masks=pat)
        worst = max((rel_err(got[k], r64[k]), k) for k in r64)
        worst32 = max((rel_err(r32[k], r64[k]), k) for k in r64)
        print(f"adabn={adabn} n={n} engine={engine}: emb err vs fp64 {rel_err(emb, e64):.2e} (oracle32 {rel_err(e32, e64):.2e}); "
              f"worst grad {worst[0]:.2e} {worst[1]} (oracle32 worst {worst32[0]:.2e})")
