"""The experiment driver with the reference's flags (code/go.sh:6 shape) end to end on the GPU:
hyper-parameter search folds -> final training with cosine annealing + checkpoint -> --test with voting."""
import os

import numpy as np
import pytest
import torch

from contrastiveprosthetics_b200 import train as cptrain

pytestmark = pytest.mark.gpu


def _args(tmp_path, extra=()):
    argv = ["--final_epochs=1", "--crossval_size=2", "--crossval_epochs=1", "--batch_size=64", "--test",
            "--synthetic", "--no_verbose", f"--data_dir={tmp_path}/data/", f"--checkpoint_dir={tmp_path}/ckpt/"]
    return cptrain.build_parser().parse_args(argv + list(extra))


@pytest.mark.parametrize("extra", [(), ("--no_adabn",), ("--lean_step",)])
def test_main_runs_with_reference_flags(tmp_path, extra):
    args = _args(tmp_path, extra)
    assert args.no_adabn is ("--no_adabn" not in extra)    # inverted store_false flag, as in the reference
    loss, acc = cptrain.main(args)
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0
    keys = np.load(f"{tmp_path}/data/cross_val_keys.npy")
    vals = np.load(f"{tmp_path}/data/cross_val_values.npy")
    assert keys.shape == (2, 7) and vals.shape == (2, 2)   # layout of data/cross_val_*.npy
    sd = torch.load(f"{tmp_path}/ckpt/contrastive.pt")
    assert len(sd) == (68 if "--no_adabn" in extra else 41)            # reference state-dict key counts


def test_item_loader_path_matches_batched_path(tmp_path):
    """DataLoader + per-item __getitem__ (the reference's loader) and the one-launch-per-batch path see the
    same samples for the same item ids."""
    from contrastiveprosthetics_b200.load import DB23
    from contrastiveprosthetics_b200.utils import TaskWrapper
    import torch.utils.data as data
    ds = DB23(db2=False, device="cuda")
    ds.load_synthetic()
    tw = TaskWrapper(ds)
    tw.set_val()
    loader = data.DataLoader(tw, batch_size=5, shuffle=False)
    EMG, GLOVE, label = next(iter(loader))
    EMG2, GLOVE2, label2 = tw.get_batch(torch.arange(5))
    assert EMG.shape == (5, 41, 25, 1, 12) and torch.equal(EMG, EMG2)
    assert torch.equal(GLOVE, GLOVE2) and torch.equal(label, label2)
