"""Oracle (TEST INFRASTRUCTURE): DB23 split/flatten/gather + TaskWrapper indexing on the CPU.

Follows /root/reference/code/load.py:157-273, utils.py:34-64, utils.py:129-130,
constants.py (values re-derived here so the oracle does not depend on the product).
numpy only.
"""
import numpy as np

EMG_DIM = 12
GLOVE_DIM = 20
VOTE_WIN = 25          # PREDICTION_WINDOW_SIZE  (constants.py:77)
OUT_DIM = 100          # WINDOW_OUTPUT_DIM       (constants.py:90)
N_VOTE_WINDOWS = 4     # AMT_PREDICTION_WINDOWS  (constants.py:78)
N_TASKS = 41


def ref_constants():
    """constants.py:3-54 -- the seed-0 permutations that define label order and subject masks."""
    rs = np.random.RandomState(0)
    d2 = rs.permutation(40)
    d3 = rs.permutation(6)
    ta = np.arange(1, 18, dtype=np.uint8)
    tb = np.arange(18, 41, dtype=np.uint8)
    rs.shuffle(ta)
    rs.shuffle(tb)
    reps = np.array([1, 3, 4, 6, 2, 5])
    return {
        "d2_idxs": d2,
        "d3_idxs": d3,
        "TASKS": np.concatenate((ta, tb)),
        "rep_train": reps[:4][:-1] - 1,     # load.py:43  -> {0,2,3}
        "rep_val": reps[:4][-1:] - 1,       # load.py:44  -> {5}
        "rep_test": reps[4:] - 1,           # load.py:45  -> {1,4}
    }


def masks(db2, split, mixed=False):
    """load.py:157-203: (tasks_mask, people_mask, rep_mask) for split in train/val/test.
    mixed (BASELINE.json config 3; not in the reference, which selects DB2 OR DB3 at load.py:179-183):
    all 46 people with the DB3 repetition split."""
    c = ref_constants()
    tasks = np.concatenate((c["TASKS"].astype(np.int64), [0]))       # label 40 = rest (stimulus 0)
    people = c["d2_idxs"] if db2 else c["d3_idxs"] + 40
    if mixed:
        people = np.concatenate((c["d2_idxs"], c["d3_idxs"] + 40))
        db2 = False
    if split == "train":
        rep = np.concatenate((c["rep_train"], c["rep_test"])) if db2 else c["rep_train"]
    elif split == "val":
        rep = c["rep_val"]
    else:
        rep = c["rep_val"] if db2 else c["rep_test"]
    return tasks, people.astype(np.int64), rep.astype(np.int64)


def load_valid(EMG, db2, split, mixed=False):
    """load.py:233-251.  EMG: (41 stimuli, 46 people, 6 reps, 100, 12) float32.

    Returns (EMG_use (41*D*W? see below), tensor, D):
      train: EMG_use (41*D, 12) with D = P*R*100, tensor = (41*D, 100? ...) unused
      val/test: tensor (41*D, 25, 12) with D = P*R*4.
    Row id = label*D + k (class-major, load.py:242-249).
    """
    t, p, r = masks(db2, split, mixed)
    sub = EMG[t][:, p][:, :, r][:, :, :, :OUT_DIM]
    if mixed:
        # the 6 DB3 (amputee) subjects are 11-channel recordings: channel index 10 carries no signal
        # (the reference's only trace of this is the commented `EMG[:, :, :, -2] = 0`, load.py:269-272)
        sub = sub.copy() if isinstance(sub, np.ndarray) else sub.clone()
        sub[:, p >= 40, :, :, EMG_DIM - 2] = 0
    EMG_use = sub.reshape(-1, EMG_DIM)
    P, R = len(p), len(r)
    if split == "train":
        D = P * R * OUT_DIM
        tensor = sub.reshape(-1, OUT_DIM, EMG_DIM)
    else:
        D = P * R * N_VOTE_WINDOWS
        tensor = sub.reshape(-1, VOTE_WIN, EMG_DIM)
    return EMG_use, tensor, D


def normalize(X, mean, std):
    """utils.py:129-130: (X - mean) / std with a true divide."""
    return (X - mean) / std


def class_permutation(rand):
    """utils.py:34-36: per-class argsort of a (41, D) random matrix plus class offset."""
    T, D = rand.shape
    return np.argsort(rand, axis=-1, kind="stable") + (np.arange(T, dtype=np.int64) * D)[:, None]


def get_items(EMG_use, tensor, emg_rand, item_idx, train):
    """utils.py:51-64 + load.py:256-273 + default_collate, for a list of item ids.

    Returns EMG batch (B,41,1,1,12) in train or (B,41,25,1,12) in val/test and labels (B,41)."""
    item_idx = np.asarray(item_idx, dtype=np.int64)
    rows = emg_rand[:, item_idx].T                        # (B, 41)
    if train:
        emg = EMG_use[rows].reshape(len(item_idx), N_TASKS, 1, 1, EMG_DIM)
    else:
        emg = tensor[rows][:, :, :, None, :]              # (B,41,25,1,12)
    labels = np.tile(np.arange(N_TASKS, dtype=np.int64), (len(item_idx), 1))
    return emg, labels
