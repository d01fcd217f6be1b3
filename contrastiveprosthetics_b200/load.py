"""Drop-in `DB23` dataset object (reference: code/load.py, runtime half: lines 23-73, 157-273).

The whole sEMG tensor stays resident in HBM as (41 stimuli, 46 people, 6 reps, 100, 12) fp32
(54 MB; load.py:66-73).  `set_train/set_val/set_test` re-slice it by the split masks into the
class-major row table `EMG_use` / `tensor`; indexing gathers rows with cp_gather_norm, optionally
fusing the emg_mean/emg_std normalisation the reference applies offline (load.py:143-147)."""
import numpy as np
import torch
import torch.utils.data as data

from .constants import (AMT_PREDICTION_WINDOWS, EMG_DIM, PATH_DIR, PEOPLE_IDXS, PREDICTION_WINDOW_SIZE,
                        REPS, REPS_TEST, REPS_TRAIN, TASKS, TEST_PEOPLE_IDXS, TEST_TASKS,
                        TRAIN_PEOPLE_IDXS, TRAIN_TASKS, VOTE, WINDOW_OUTPUT_DIM, d2_idxs, d3_idxs)
from .utils import Glover, default_device, gather_rows


class DB23(data.Dataset):
    def __init__(self, db2=False, train=True, val=False, device=None, emg_stats=None, mixed=False):
        self.device = torch.device(device) if device is not None else default_device()
        self.train = train
        self.val = val
        self.raw = False
        self.db2 = db2 and not mixed
        # mixed (BASELINE.json config 3; the reference picks DB2 OR DB3, load.py:179-183): all 46 subjects with
        # the DB3 repetition split; the 6 DB3 amputee subjects are 11-channel, their channel 10 is zeroed
        self.mixed = mixed
        self.emg_stats = emg_stats          # RunningStats -> normalise inside the gather kernel

        dev = lambda x: torch.from_numpy(np.array(x)).to(self.device)      # noqa: E731  (torchize, utils.py:18)
        self.tasks_train, self.tasks_test, self.tasks = dev(TRAIN_TASKS), dev(TEST_TASKS), dev(TASKS)
        self.people_train, self.people_test = dev(TRAIN_PEOPLE_IDXS), dev(TEST_PEOPLE_IDXS)
        self.people = dev(PEOPLE_IDXS)
        train_reps, test_reps = dev(REPS_TRAIN), dev(REPS_TEST)
        self.rep_train = train_reps[:-1] - 1        # {0,2,3}
        self.rep_val = train_reps[-1:] - 1          # {5}
        self.rep_test = test_reps - 1               # {1,4}
        self.reps = dev(REPS) - 1
        self._d2 = dev(d2_idxs)
        self._d3 = dev(d3_idxs + len(d2_idxs))
        self.glover = Glover(self.device)
        self.EMG = None
        self.EMG_use = None
        self.tensor = None

    # ---- split selection (load.py:51-64)
    def set_train(self):
        self.train, self.val = True, False
        self.load_valid()

    def set_val(self):
        self.train, self.val = False, True
        self.load_valid()

    def set_test(self):
        self.train, self.val = False, False
        self.load_valid()

    # ---- data sources
    def load_stored(self, emg_path=None, glove_path=None):
        """load.py:66-73: emg.pt is (people, tasks, reps, 100, 12); transposed to tasks-major."""
        emg = torch.load(emg_path or (PATH_DIR + 'data/emg.pt'), map_location=self.device)
        self.EMG = emg.transpose(0, 1)
        self.GLOVE = self.glover.load_stored(glove_path)

    def load_tensors(self, emg_people_major, glove=None):
        """Install already-loaded tensors: emg (46,41,6,100,12) like emg.pt; glove (41,Dg,20) or None."""
        self.EMG = emg_people_major.to(self.device, torch.float32).transpose(0, 1)
        self.glover.GLOVE = None if glove is None else glove.to(self.device, torch.float32)
        self.GLOVE = self.glover.GLOVE

    def load_synthetic(self, seed=0, with_glove=True, glove_dim=20):
        """Seeded NinaPro-shaped stand-in (the dataset download is unavailable offline)."""
        from .synthetic import synth_emg, synth_glove
        self.load_tensors(synth_emg(seed), synth_glove(seed + 1, glove_dim, class_offset=glove_dim != 20) if with_glove else None)

    # ---- masks (load.py:157-203)
    @property
    def tasks_mask(self):
        return torch.cat((self.tasks.to(torch.long), torch.zeros(1, dtype=torch.long, device=self.device)))

    @property
    def people_mask(self):
        if self.mixed:
            return torch.cat((self._d2, self._d3))
        return self._d2 if self.db2 else self._d3

    @property
    def rep_mask(self):
        if self.train:
            return torch.cat((self.rep_train, self.rep_test)) if self.db2 else self.rep_train
        if self.val:
            return self.rep_val
        return self.rep_val if self.db2 else self.rep_test

    @property
    def PEOPLE(self):
        return len(self.people_mask)

    @property
    def TASKS(self):
        return len(self.tasks_mask)

    @property
    def REPS(self):
        return len(self.rep_mask)

    @property
    def OUTPUT_DIM(self):
        if self.train:
            return WINDOW_OUTPUT_DIM
        return PREDICTION_WINDOW_SIZE if VOTE else WINDOW_OUTPUT_DIM

    @property
    def D(self):
        per_rep = WINDOW_OUTPUT_DIM if (self.train or not VOTE) else AMT_PREDICTION_WINDOWS
        return self.PEOPLE * self.REPS * per_rep

    # ---- re-slice (load.py:233-251): class-major row table, row id = class*D + k
    def load_valid(self):
        if self.EMG is None:
            raise RuntimeError("no data: call load_stored(), load_tensors() or load_synthetic() first")
        sub = self.EMG[self.tasks_mask][:, self.people_mask][:, :, self.rep_mask][:, :, :, :WINDOW_OUTPUT_DIM]
        sub = sub.contiguous()
        if self.mixed:
            sub[:, self.people_mask >= len(self._d2), :, :, EMG_DIM - 2] = 0       # load.py:269-272 (commented there)
        self.EMG_use = sub.reshape(-1, EMG_DIM)
        self.tensor = sub.reshape(-1, self.OUTPUT_DIM, EMG_DIM)
        self._rows2d = self.tensor.reshape(self.tensor.shape[0], -1)       # (41*D, 25*12) in eval
        probe = self.D * 2 + 1                                               # load.py:242-249
        if self.train or not VOTE:
            ok = torch.equal(self.EMG_use[probe], sub[2].reshape(-1, EMG_DIM)[1])
        else:
            ok = torch.equal(self.tensor[probe], sub[2].reshape(-1, self.OUTPUT_DIM, EMG_DIM)[1])
        assert ok, "indexing is not correct"
        self.glover.load_valid(self.tasks_mask)

    def __len__(self):
        return self.TASKS * self.D

    def subjects_of(self, idx):
        """Subject (index into constants' 46-person axis) of every row id in `idx` -- row id = class*D + k with k
        running over (person, repetition, window) of the current split (load.py:233-251).  Feeds the per-subject
        AdaBN of models.EMGNet (models.py:245)."""
        idx = torch.as_tensor(idx, device=self.device)
        per_person = self.D // self.PEOPLE
        return self.people_mask.to(torch.long)[(idx % self.D) // per_person]

    def _stats(self):
        if self.emg_stats is None:
            return None, None
        return self.emg_stats.mean_std()

    def slice_batch(self, idx):
        """load.py:256-259: training rows -> (-1,1,1,12)."""
        mean, std = self._stats()
        return gather_rows(self.EMG_use, idx, mean, std, EMG_DIM).reshape(-1, 1, 1, EMG_DIM)

    def __getitem__(self, idx):
        """load.py:261-273.  idx: int64 tensor of row ids, shape (41,) per item or (B,41) per batch."""
        if self.raw:
            return self.EMG
        idx = torch.as_tensor(idx, device=self.device)
        lead = tuple(idx.shape)
        mean, std = self._stats()
        if not self.train and VOTE:
            out = gather_rows(self._rows2d, idx, mean, std, EMG_DIM)
            return out.reshape(lead + (self.OUTPUT_DIM, 1, EMG_DIM))        # (...,25,1,12)
        out = gather_rows(self.EMG_use, idx, mean, std, EMG_DIM)
        if len(lead) <= 1:
            return out.reshape(-1, 1, 1, EMG_DIM)                            # (41,1,1,12) like slice_batch
        return out.reshape(lead + (1, 1, EMG_DIM))                           # (B,41,1,1,12)
