"""Shape / split constants of the NinaPro DB2+DB3 contrastive task.

Restates the values of the reference's constants module
(/root/reference/code/constants.py) without touching numpy's global RNG: the
reference seeds the global legacy generator with 0 and draws, in this order,
permutation(40), permutation(6), shuffle(tasks 1..17), shuffle(tasks 18..40)
(constants.py:3,18-19,39-40).  `RandomState(0)` is the same MT19937 stream.

Names are kept identical to the reference so `from constants import *`
callers (train.py, load.py, utils.py, models.py) keep working.
"""
import numpy as np

_rs = np.random.RandomState(0)

# ---- subjects (constants.py:5-35) -------------------------------------------
MAX_PEOPLE_D2 = 40
_D3_SUBJECTS = (2, 3, 4, 5, 8, 9)
MAX_PEOPLE_D3 = len(_D3_SUBJECTS)
MAX_PEOPLE = MAX_PEOPLE_D2 + MAX_PEOPLE_D3

d2_idxs = _rs.permutation(MAX_PEOPLE_D2)
d3_idxs = _rs.permutation(MAX_PEOPLE_D3)

ORIGINAL_D3 = np.array([s + MAX_PEOPLE_D2 - 1 for s in _D3_SUBJECTS])
PEOPLE_D2 = np.arange(MAX_PEOPLE_D2)[d2_idxs]
PEOPLE_D3 = ORIGINAL_D3[d3_idxs]
PEOPLE = np.concatenate((PEOPLE_D2, PEOPLE_D3))
PEOPLE_IDXS = np.concatenate((d2_idxs, d3_idxs + MAX_PEOPLE_D2))

TRAIN_PEOPLE_IDXS = PEOPLE_IDXS
TEST_PEOPLE_IDXS = PEOPLE_IDXS
TRAIN_PEOPLE = PEOPLE[TRAIN_PEOPLE_IDXS]
TEST_PEOPLE = PEOPLE[TEST_PEOPLE_IDXS]
MAX_PEOPLE_TRAIN = MAX_PEOPLE
MAX_PEOPLE_TEST = MAX_PEOPLE

# ---- grasp classes (constants.py:37-48) -------------------------------------
TASK_DIST = np.array([17, 23])
TASKS_A = np.arange(1, 18, dtype=np.uint8)
TASKS_B = np.arange(18, 41, dtype=np.uint8)
_rs.shuffle(TASKS_A)
_rs.shuffle(TASKS_B)
TASKS = np.concatenate((TASKS_A, TASKS_B))
TEST_TASKS = TASKS[:]
TRAIN_TASKS = TASKS[:]
MAX_TASKS = int(TASK_DIST.sum()) + 1          # 40 grasps + rest
MAX_TASKS_TRAIN = MAX_TASKS

# ---- repetitions (constants.py:50-54) ---------------------------------------
REPS = [1, 3, 4, 6, 2, 5]
MAX_REPS = len(REPS)
REPS_TRAIN = REPS[:-2]
REPS_TEST = REPS[-2:]

PATH_DIR = "/home/breezy/hci/prosthetics/db23/"   # constants.py:56 (absent offline)
BLOCK_SIZE = 1

# ---- sampling / windows (constants.py:60-93) --------------------------------
Hz = 2000
DOWNSAMPLE = 100
FACTOR = Hz // DOWNSAMPLE
RMS_WINDOW = 11
WINDOW_EDGE = (RMS_WINDOW - 1) // 2
TOTAL_WINDOW_SIZE = Hz * 1
FINAL_WINDOW_SIZE = TOTAL_WINDOW_SIZE // FACTOR            # 100 samples / repetition

VOTE = True
PREDICTION_WINDOW = 250                                     # ms; also the vote-loop bound (models.py:153)
PREDICTION_WINDOW_SIZE = PREDICTION_WINDOW * DOWNSAMPLE // 1000   # 25 samples voted together
AMT_PREDICTION_WINDOWS = FINAL_WINDOW_SIZE // PREDICTION_WINDOW_SIZE   # 4 vote windows / repetition
assert FINAL_WINDOW_SIZE % AMT_PREDICTION_WINDOWS == 0

Hz_glove = 25
GLOVE_FACTOR = int(1 / Hz_glove * Hz)
GLOVE_WINDOW_SIZE = TOTAL_WINDOW_SIZE // GLOVE_FACTOR

WINDOW_MS = 1
WINDOW_STRIDE = 1
WINDOW_OUTPUT_DIM = FINAL_WINDOW_SIZE
assert FINAL_WINDOW_SIZE % WINDOW_OUTPUT_DIM == 0
assert FINAL_WINDOW_SIZE % WINDOW_MS == 0
AMT_WINDOWS = FINAL_WINDOW_SIZE // WINDOW_MS

GLOVE_DIM = 22 - 2       # sensors 6 and 11 dropped (constants.py:96)
EMG_DIM = 12
