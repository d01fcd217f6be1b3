import os as _os

# folds.ConcurrentFolds replays ONE CUDA graph with K parallel branches (one per cross-validation fold, train.py:140-166):
# branches only run side by side when they land on different hardware work queues, and the driver creates 8 by default
# (measured on B200, K = 16 folds of batch_size 8: 6.9 k fold-steps/s with 8 queues, 8.3 k with 32).  The variable is read
# when the CUDA context is created, so it has to be in the environment before the first CUDA call; a value the user set wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
