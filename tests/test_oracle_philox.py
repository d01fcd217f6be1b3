"""The Philox4x32-10 restatement (oracle/philox.py) against the published known-answer vectors (Random123
kat_vectors; Salmon et al., SC'11) -- the pin of the dropout generator's oracle."""
import numpy as np

from oracle import philox as OP


def test_known_answer_vectors():
    for ctr, key, out in OP.KAT:
        got = OP.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(x) for x in got] == list(out), (ctr, key, [hex(int(x)) for x in got])


def test_mask_mapping_and_rate():
    m = OP.dropout_mask(1 << 16, 0.5, seed=0x5EED, layer=1)
    assert m.dtype == np.uint8 and set(np.unique(m)) <= {0, 1}
    n = m.size
    assert abs(m.mean() - 0.5) < 4 * 0.5 / np.sqrt(n)
    assert not np.array_equal(m, OP.dropout_mask(1 << 16, 0.5, seed=0x5EED, layer=2))
    assert not np.array_equal(m, OP.dropout_mask(1 << 16, 0.5, seed=0x5EED, layer=1, step=1))
    assert np.array_equal(m, OP.dropout_mask(1 << 16, 0.5, seed=0x5EED, layer=1, step=0))
