"""ConcurrentFolds (train.py:140-166): K folds advanced in lockstep on K streams / CUDA graphs end
bit-identical to the same folds trained one after the other on the same batches; train.py --concurrent_folds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PLIST = [{'d_e': 16, 'dp_emg': 0.0, 'dp_glove': 0.0, 'reg_emg': 1e-5, 'reg_glove': 1e-5, 'lr_emg': 1e-3, 'lr_glove': 1e-3},
         {'d_e': 16, 'dp_emg': 0.5, 'dp_glove': 0.2, 'reg_emg': 1e-4, 'reg_glove': 1e-6, 'lr_emg': 3e-4, 'lr_glove': 1e-2},
         {'d_e': 16, 'dp_emg': 0.4, 'dp_glove': 0.0, 'reg_emg': 1e-7, 'reg_glove': 1e-3, 'lr_emg': 1e-2, 'lr_glove': 1e-4}]


def _dataset():
    from contrastiveprosthetics_b200.load import DB23
    from contrastiveprosthetics_b200.utils import TaskWrapper
    ds = DB23(db2=False, device="cuda")
    ds.load_synthetic(with_glove=False)
    tw = TaskWrapper(ds, with_glove=False)
    tw.set_train()
    return tw


def test_concurrent_folds_equal_sequential_folds():
    from contrastiveprosthetics_b200.folds import ConcurrentFolds
    tw = _dataset()
    B, n_steps = 8, 12
    items = [torch.randperm(tw.D, generator=torch.Generator().manual_seed(s))[:B] for s in range(n_steps)]
    batches = [tw.get_batch(i)[0] for i in items]
    together = ConcurrentFolds(tw, PLIST, B)
    for EMG in batches:
        together.step(EMG)
    together.join()
    torch.cuda.synchronize()
    for k, p in enumerate(PLIST):
        alone = ConcurrentFolds(tw, [p], B, seeds=[42], dropout_seeds=[1000 + k])
        for EMG in batches:
            alone.step(EMG)
        alone.join()
        torch.cuda.synchronize()
        sd_a, sd_t = alone.models[0].state_dict(), together.models[k].state_dict()
        for name in sd_a:
            assert torch.equal(sd_a[name], sd_t[name]), (k, name)
        assert [l.item() for l in alone.losses[0]] == [l.item() for l in together.losses[k]]
    # the folds really are different models
    assert not torch.equal(together.models[0].state_dict()["emg_net.last.0.weight"],
                           together.models[1].state_dict()["emg_net.last.0.weight"])


def test_run_epoch_with_ragged_last_batch():
    from contrastiveprosthetics_b200.folds import ConcurrentFolds
    tw = _dataset()
    B = 512                                   # D = 1800 train items (DB3-shaped): 3 full batches + one of 264
    folds = ConcurrentFolds(tw, PLIST[:2], B)
    losses = folds.run_epoch(generator=torch.Generator().manual_seed(0))
    assert len(losses) == 2 and all(np.isfinite(l) and 0 < l < 10 for l in losses)
    assert all(len(l) == 4 for l in folds.losses)


def test_train_script_concurrent_folds(tmp_path):
    from contrastiveprosthetics_b200 import train as cptrain
    argv = ["--final_epochs=1", "--crossval_size=3", "--crossval_epochs=1", "--batch_size=256", "--synthetic",
            "--no_verbose", "--concurrent_folds=3", f"--data_dir={tmp_path}/data/", f"--checkpoint_dir={tmp_path}/ckpt/"]
    loss, acc = cptrain.main(cptrain.build_parser().parse_args(argv))
    assert np.isfinite(loss) and 0.0 <= acc <= 1.0
    vals = np.load(f"{tmp_path}/data/cross_val_values.npy")
    assert vals.shape == (3, 2) and np.isfinite(vals).all()
