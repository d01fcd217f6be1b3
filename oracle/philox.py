"""Philox4x32-10 restated from its specification (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
1, 2, 3", SC'11; Random123) in numpy uint64 arithmetic, plus the keep-mask mapping of the library's dropout
(include/cpros.h cp_dropout_mask).  Test infrastructure only.

Pinned by the published known-answer vectors (Random123 kat_vectors, philox4x32 10 rounds): KAT below;
tests/test_oracle_philox.py checks this restatement against them on the CPU, tests/test_gpu_philox.py checks the CUDA
generator against both."""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF

# (counter[4], key[2]) -> output[4]
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(ctr, key):
    """ctr (..., 4), key (..., 2) uint32 arrays -> (..., 4) uint32."""
    c = [np.asarray(ctr)[..., i].astype(np.uint64) for i in range(4)]
    k = [np.asarray(key)[..., i].astype(np.uint64) for i in range(2)]
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + np.uint64(W0)) & np.uint64(MASK), (k[1] + np.uint64(W1)) & np.uint64(MASK)]
    return np.stack(c, axis=-1).astype(np.uint32)


def dropout_mask(n, p, seed, layer, step=None):
    """uint8 keep mask of n elements (n % 4 == 0) as cp_dropout_mask draws it."""
    assert n % 4 == 0
    if step is not None:
        seed = (seed + step * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    v = np.arange(n // 4, dtype=np.uint64)
    ctr = np.stack([v & np.uint64(MASK), v >> np.uint64(32), np.full_like(v, layer), np.full_like(v, 0x43505253)], -1)
    key = np.broadcast_to(np.array([seed & MASK, seed >> 32], dtype=np.uint64), (n // 4, 2))
    r = philox4x32_10(ctr, key)
    thr = np.uint32(min(np.float32(p) * np.float32(4294967296.0), np.float32(4294967040.0)))
    return (r >= thr).astype(np.uint8).reshape(-1)
