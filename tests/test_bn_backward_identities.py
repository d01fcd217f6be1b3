"""The identities behind the reduce-free BatchNorm backward (csrc/encoder_kernels.cuh, bn_bwd_stats_from_wgrad_kernel),
checked in float64 on the CPU against the direct sums -- the derivation itself, independent of any kernel.

A BN stage with output A = (gamma * xh + beta) * M (M = dropout keep mask / (1-p), or all ones) feeds a Linear layer
Z = A W^T + b.  With G1 = dL/dZ:   g = dL/dA = G1 W,   g' = g * M (gradient w.r.t. the un-masked BN output),
db = colsum(G1),   dW = G1^T A.   The BN backward needs  S1 = sum_r g'  and  S2 = sum_r g' * xh  per feature.
"""
import numpy as np
import pytest


def _setup(R, K, F, p, seed):
    rs = np.random.RandomState(seed)
    xh = rs.randn(R, F)
    gamma, beta = 1 + 0.3 * rs.randn(F), 0.2 * rs.randn(F)
    M = (rs.rand(R, F) > p) / (1 - p) if p > 0 else np.ones((R, F))
    A = (gamma * xh + beta) * M
    W = rs.randn(K, F) / np.sqrt(F)
    G1 = rs.randn(R, K)
    g = G1 @ W
    return xh, gamma, beta, M, A, W, G1, g


@pytest.mark.parametrize("p", [0.0, 0.5])
def test_linear_stage(p):
    xh, gamma, beta, M, A, W, G1, g = _setup(R=300, K=48, F=40, p=p, seed=1)
    gp = g * M
    S1, S2 = gp.sum(0), (gp * xh).sum(0)
    db, dW = G1.sum(0), G1.T @ A
    if p == 0:
        np.testing.assert_allclose(db @ W, S1, rtol=1e-10, atol=1e-10)          # first identity (no mask only)
    # second identity holds with or without the mask, given S1
    np.testing.assert_allclose(((W * dW).sum(0) - beta * S1) / gamma, S2, rtol=1e-9, atol=1e-9)


def test_conv_stage_group_of_12_positions():
    """conv2 BN (64 channels, 12 positions per window) feeding fc1 through the flatten column ch*12 + p
    (models.py:263): the per-channel sums fold the 12 columns of a channel."""
    R, C, P, K = 50, 6, 12, 20
    rs = np.random.RandomState(2)
    xh = rs.randn(R, C, P)
    gamma, beta = 1 + 0.3 * rs.randn(C), 0.2 * rs.randn(C)
    A = (gamma[:, None] * xh + beta[:, None]).reshape(R, C * P)                  # flatten: column ch*12 + p
    W = rs.randn(K, C * P)
    G1 = rs.randn(R, K)
    g = (G1 @ W).reshape(R, C, P)
    S1, S2 = g.sum((0, 2)), (g * xh).sum((0, 2))
    db, dW = G1.sum(0), G1.T @ A
    a = (db @ W).reshape(C, P).sum(1)
    t = (W * dW).sum(0).reshape(C, P).sum(1)
    np.testing.assert_allclose(a, S1, rtol=1e-10)
    np.testing.assert_allclose((t - beta * a) / gamma, S2, rtol=1e-9)


def test_gamma_zero_is_unobservable():
    """gamma_c = 0: A[:, c] no longer depends on xh, so dW carries no information about S2[c] (the kernel raises a flag
    and the reduce pass runs instead) -- while the data gradient, which is multiplied by gamma_c, is 0 anyway."""
    xh, gamma, beta, M, A, W, G1, g = _setup(R=200, K=16, F=8, p=0.0, seed=3)
    gamma[3] = 0.0
    A = gamma * xh + beta
    dW = G1.T @ A
    S1 = g.sum(0)
    num = (W * dW).sum(0) - beta * S1
    assert abs(num[3]) < 1e-9                       # 0 / 0: nothing to divide
    assert abs((g * xh).sum(0)[3]) > 1e-3           # the true sum is far from 0
