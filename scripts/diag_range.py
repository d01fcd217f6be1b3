"""Per-tensor gradient errors (vs float64) of one dynamic-range case, both engines.  python scripts/diag_range.py gamma weight"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from contrastiveprosthetics_b200 import _lib
from gpu_util import encoder_parity_errors, perturbed_state, scale_state
gamma, weight = float(sys.argv[1]), float(sys.argv[2])
n = 41 * 40
sd = scale_state(perturbed_state(29, True), True, gamma, weight)
g = torch.Generator().manual_seed(200)
x = torch.randn(n, 12, generator=g) + 0.5 * torch.randn(41, 12, generator=g).repeat(40, 1)
d_emb = torch.randn(n, 16, generator=g) / n
g = torch.Generator().manual_seed(9)
masks = [torch.empty(n, 512, dtype=torch.uint8).bernoulli_(0.5, generator=g) for _ in range(4)]
for name, eng in (("simt", _lib.ENGINE_SIMT), ("tc", _lib.ENGINE_TC)):
    e = encoder_parity_errors(sd, True, x, d_emb, eng, 0.5, masks, fp64=True)
    print("==", name, "emb", e["emb"], "emb64", e["emb64"])
    for k in sorted(k for k in e if k.startswith("grad64|")):
        p = k.split("|")[1]
        print(f"  {p:36s} cuda-vs-64 {e[k]:.2e}  oracle32-vs-64 {e['oracle32_vs_64|' + p]:.2e}  cuda-vs-32 {e['grad|' + p]:.2e}")
