"""Encoder parity at the size that is benchmarked, and over the dynamic range of the fp16-plane operand format.

* C2 size (BASELINE.json configs[1]: 4096 groups x 41 = 167,936 windows, AdaBN, dropout 0.5 with injected masks)
  and a C3-shaped mixed-subject batch (DB3 subjects' channel 10 zeroed, load.py:269-272): the size-dependent
  mechanisms of the tensor-core engine -- weight-gradient chains cut every 512 rows and re-summed over 328 chunks,
  per-layer power-of-two plane scales, two-level partial sums over 1,312 tiles, 32-bit gather indexing -- against
  the fp32 oracle with the kernel's ReLU pattern (DESIGN.md "ReLU kinks").
* range: BatchNorm gammas x {1e-3, 3e2}, weights x {1e-4, 1e3}, inputs x 1e3.

Tolerance: north_star's 1e-5 relative (norm-wise per tensor) for the loss-side values (embeddings, every stage)
and for the gradients.  Measured worst cases: profiles/parity_r2.json (scripts/parity_report.py).

What the 1e-5 is measured against at C2 size.  The reference's own fp32 arithmetic is NOT 1e-5-accurate at 167,936 rows:
torch's CPU BatchNorm1d accumulates its statistics in float32, and the fp32 oracle (= the reference's CPU path) sits
1.2e-5 (20,992 rows) ... 1e-4 (167,936 rows) from a float64 evaluation of the same network, growing by ~1.2e-5 per
BatchNorm1d layer (profiles/parity_r2.json: `worst_fp32_oracle_vs_fp64`).  Two fp32 implementations cannot agree better
than either agrees with the exact result, so at this size the kernels are held to 1e-5 against the FLOAT64 oracle (their
statistics are finalised in double: 3e-6 ... 6e-6 measured), and against the fp32 oracle to that oracle's own distance
from float64."""
import os

import pytest
import torch

from contrastiveprosthetics_b200 import _lib
from gpu_util import encoder_parity_errors, perturbed_state, scale_state, worst

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _inputs(n, seed, mixed=False, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 12, generator=g) + 0.5 * torch.randn(41, 12, generator=g).repeat((n + 40) // 41, 1)[:n]
    if mixed:                       # 6 of 46 subjects are 11-channel DB3 recordings: channel 10 is zero
        x[torch.rand(n, generator=g) < 6.0 / 46.0, 10] = 0.0
    d_emb = torch.randn(n, 16, generator=g) / n
    return x * scale, d_emb


@pytest.mark.parametrize("mixed", [False, True])
def test_c2_size_forward_and_gradients(mixed):
    n, dp = 4096 * 41, 0.5
    torch.set_num_threads(os.cpu_count() or 8)
    sd = perturbed_state(3, True)
    x, d_emb = _inputs(n, 100 + mixed, mixed)
    g = torch.Generator().manual_seed(7)
    masks = [torch.empty(n, 512, dtype=torch.uint8).bernoulli_(0.5, generator=g) for _ in range(4)]
    e = encoder_parity_errors(sd, True, x, d_emb, _lib.ENGINE_TC, dp, masks, fp64=True)
    # against the exact (float64) evaluation, same ReLU pattern: north_star's 1e-5 for the embeddings (the loss side)
    # and for the gradient as a whole; per tensor 1e-5, or -- for the cancellation-dominated bias-type sums over
    # 167,936 rows (measured 1.1e-5 ... 1.7e-5) -- at least as close to float64 as the reference's own fp32
    # arithmetic is (3.5e-5 ... 1.1e-4 on EVERY tensor at this size)
    assert e["emb64"] < TOL, e["emb64"]
    assert e["grad64_global"] < TOL, e["grad64_global"]
    for k in [k for k in e if k.startswith("grad64|")]:
        name = k.split("|", 1)[1]
        assert e[k] < max(TOL, e["oracle32_vs_64|" + name]), (name, e[k], e["oracle32_vs_64|" + name])
    assert worst(e, "grad64|")[0] < 2.5 * TOL, worst(e, "grad64|")
    # against the fp32 oracle: no further from it than it is from float64 itself (+ the kernel's own 1e-5)
    o32 = worst(e, "oracle32_vs_64|")[0]
    assert e["emb"] < TOL + 2 * max(o32, e["emb_oracle32_vs_64"]), (e["emb"], o32)
    assert worst(e, "grad|")[0] < TOL + 2 * o32, (worst(e, "grad|"), o32)
    assert e["relu_flip_fraction"] < 1e-4


@pytest.mark.parametrize("engine", [_lib.ENGINE_SIMT, _lib.ENGINE_TC])
@pytest.mark.parametrize("gamma,weight,xscale", [(1e-3, 1.0, 1.0), (3e2, 1.0, 1.0), (1.0, 1e-4, 1.0), (1.0, 1e3, 1.0),
                                                 (1.0, 1.0, 1e3), (3e2, 1e3, 1e3), (1e-3, 1e-4, 1.0)])
def test_dynamic_range(gamma, weight, xscale, engine):
    """Trained BatchNorm gammas far from 1, tiny / huge weights, un-normalised inputs: the plane format (fp16 hi/lo with
    per-tensor power-of-two scales) must keep fp32-level accuracy, not silently fall into fp16 subnormals / overflow."""
    n, dp = 41 * 40, 0.5
    sd = scale_state(perturbed_state(29, True), True, gamma, weight)
    x, d_emb = _inputs(n, 200, scale=xscale)
    g = torch.Generator().manual_seed(9)
    masks = [torch.empty(n, 512, dtype=torch.uint8).bernoulli_(0.5, generator=g) for _ in range(4)]
    e = encoder_parity_errors(sd, True, x, d_emb, engine, dp, masks, fp64=True)
    assert e["emb"] < TOL, e["emb"]
    assert worst(e, "stage")[0] < TOL, worst(e, "stage")
    # gradients.  The whole gradient as one vector: 1e-5 against float64.  Per tensor: within the budget against the
    # fp32 oracle, or as close to the float64 truth as the fp32 oracle itself is (tiny weights / gammas make the
    # BatchNorm of a nearly-constant pre-activation ill-conditioned in ANY fp32 evaluation) -- except tensors whose
    # gradient has VANISHED (norm below 1e-6 of the largest tensor's: nine layers of gamma = 1e-3 leave rounding noise
    # in the first layers for every implementation, the fp32 oracle included), which only count through the global norm.
    assert e["grad64_global"] < TOL, (e["grad64_global"], e["oracle32_vs_64_global"])
    top = max(v for k, v in e.items() if k.startswith("norm64|"))
    for k in [k for k in e if k.startswith("grad|")]:
        name = k.split("|", 1)[1]
        if e["norm64|" + name] < 1e-6 * top:
            continue
        ok = e[k] < TOL or e["grad64|" + name] < max(3 * e["oracle32_vs_64|" + name], TOL)
        assert ok, (name, e[k], e["grad64|" + name], e["oracle32_vs_64|" + name], e["norm64|" + name] / top)
