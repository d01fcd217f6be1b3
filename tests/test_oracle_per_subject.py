"""Per-subject AdaBN restatement (models.py:245; parity unpinned -- no implementation exists in the reference):
consistency of oracle.model's `subjects=` path with the reference-pinned pooled path, on the CPU."""
import torch

from oracle import model as OM


def test_one_subject_equals_pooled_and_segments_equal_separate_passes():
    sd = OM.init_state(42, True)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(300, 12, generator=g)
    pooled = OM.encoder_forward(sd, x, True, True)
    same = OM.encoder_forward(sd, x, True, True, subjects=torch.full((300,), 7))
    assert torch.equal(pooled, same)
    subj = torch.randint(0, 3, (300,), generator=g)
    out = OM.encoder_forward(sd, x, True, True, subjects=subj)
    for s in range(3):
        rows = torch.nonzero(subj == s).reshape(-1)
        alone = OM.encoder_forward(sd, x[rows], True, True)
        assert torch.allclose(out[rows], alone, rtol=0, atol=1e-6)
    assert not torch.allclose(out, pooled, atol=1e-3)
