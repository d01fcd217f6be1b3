"""Dropout mask generator (csrc/common.cuh Philox4x32-10) on the GPU: published known-answer vectors, bit-exact
agreement with the specification-level restatement (oracle/philox.py), keep rate within 4 sigma, independence
across layers / steps / seeds, and that the encoder's in-kernel masks ARE cp_dropout_mask's (models.py:282-297)."""
import ctypes

import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200 import _lib
from contrastiveprosthetics_b200.models import Model
from oracle import philox as OP
from gpu_util import load_sd, perturbed_state

pytestmark = pytest.mark.gpu


def _mask(n, p, seed, layer, step=None):
    L = _lib.lib()
    keep = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = None if step is None else torch.tensor([step], dtype=torch.int64, device="cuda")
    _lib.check(L.cp_dropout_mask(_lib.ptr(keep), n, p, seed, layer, _lib.ptr(st), _lib.stream()), "cp_dropout_mask")
    return keep.cpu().numpy()


def test_known_answer_vectors_on_the_gpu():
    L = _lib.lib()
    g = np.random.RandomState(0)
    ctr = np.concatenate([np.array([k[0] for k in OP.KAT], dtype=np.uint32), g.randint(0, 2 ** 32, (1000, 4), dtype=np.uint64).astype(np.uint32)])
    key = np.concatenate([np.array([k[1] for k in OP.KAT], dtype=np.uint32), g.randint(0, 2 ** 32, (1000, 2), dtype=np.uint64).astype(np.uint32)])
    c = torch.from_numpy(ctr.view(np.int32)).cuda()
    k = torch.from_numpy(key.view(np.int32)).cuda()
    out = torch.empty_like(c)
    _lib.check(L.cp_philox4x32_10(_lib.ptr(c), _lib.ptr(k), ctr.shape[0], _lib.ptr(out), _lib.stream()))
    got = out.cpu().numpy().view(np.uint32)
    for i, (_, _, want) in enumerate(OP.KAT):
        assert [int(x) for x in got[i]] == list(want)
    assert np.array_equal(got, OP.philox4x32_10(ctr, key))


@pytest.mark.parametrize("p", [0.1, 0.5, 0.9])
def test_mask_is_the_specified_stream_and_keeps_at_rate(p):
    n = 1 << 20
    m = _mask(n, p, 0x5EED, 2, step=3)
    assert np.array_equal(m, OP.dropout_mask(n, p, 0x5EED, 2, step=3))          # bit-exact with the restatement
    sigma = np.sqrt(p * (1 - p) / n)
    assert abs(m.mean() - (1 - p)) < 4 * sigma, (m.mean(), 1 - p, sigma)
    # no structure along the row (512 columns) or the 4 words of a Philox block
    cols = m.reshape(-1, 512).mean(0)
    assert np.abs(cols - (1 - p)).max() < 5 * np.sqrt(p * (1 - p) / (n / 512))
    words = m.reshape(-1, 4).mean(0)
    assert np.abs(words - (1 - p)).max() < 4.5 * np.sqrt(p * (1 - p) / (n / 4))


def test_layers_steps_and_seeds_are_independent():
    n, p = 1 << 20, 0.5
    base = _mask(n, p, 0x5EED, 0, step=1).astype(np.float64) - 0.5
    others = {"layer": _mask(n, p, 0x5EED, 1, step=1), "step": _mask(n, p, 0x5EED, 0, step=2),
              "seed": _mask(n, p, 0x5EED + 1, 0, step=1), "rank": _mask(n, p, 0x5EED ^ (1 << 48), 0, step=1)}
    for what, m in others.items():
        corr = float((base * (m.astype(np.float64) - 0.5)).mean() / 0.25)          # ~ N(0, 1/n) when independent
        assert abs(corr) < 4 / np.sqrt(n), (what, corr)
    # shifted copies of the stream do not correlate either
    a = _mask(n, p, 0x5EED, 0, step=1).astype(np.float64) - 0.5
    for lag in (1, 4, 512):
        corr = float((a[:-lag] * a[lag:]).mean() / 0.25)
        assert abs(corr) < 4 / np.sqrt(n - lag), (lag, corr)


@pytest.mark.parametrize("engine", [_lib.ENGINE_SIMT, _lib.ENGINE_TC])
def test_encoder_draws_exactly_these_masks(engine):
    """cp_encoder_forward with in-kernel dropout == cp_encoder_forward fed the same masks from cp_dropout_mask."""
    n, p = 41 * 12, 0.5
    sd = perturbed_state(19, True)
    params = {'d_e': 16, 'dp_emg': p, 'dp_glove': 0.0, 'reg_emg': 0.0, 'reg_glove': 0.0}
    m = Model(params, adabn=True, device="cuda")
    load_sd(m, sd)
    m.emg_net.engine = engine
    m.train(True)
    x = torch.randn(n, 12, generator=torch.Generator().manual_seed(2)).cuda()
    step = torch.tensor([5], dtype=torch.int64, device="cuda")
    m.emg_net.dropout_step = step
    with torch.no_grad():
        a = m.emg_net.encode_flat(x)
    seed = (m.emg_net.dropout_seed * 1000003 + m.emg_net._step) & 0xFFFFFFFFFFFFFFFF
    masks = np.stack([_mask(n * 512, p, seed, layer, step=5).reshape(n, 512) for layer in range(4)])
    m.emg_net.ext_dropout_masks = torch.from_numpy(masks).cuda().contiguous()
    m.emg_net._step -= 1
    with torch.no_grad():
        b = m.emg_net.encode_flat(x)
    assert torch.equal(a, b)
    assert 0.45 < masks.mean() < 0.55
