"""Oracle (TEST INFRASTRUCTURE): ctypes binding of the plain-C vote / subset twin (oracle/vote_subset.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_vote.so")


def build(force=False):
    src = os.path.join(_HERE, "vote_subset.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def vote(preds, n_votes=249):
    preds = np.ascontiguousarray(preds, dtype=np.int32)
    B, W, T = preds.shape
    votes = np.empty((B, n_votes), dtype=np.int32)
    y_pred = np.empty((B, T), dtype=np.int64)
    lib().oracle_vote(_p(preds), ctypes.c_int64(B), ctypes.c_int(W), ctypes.c_int(T),
                      ctypes.c_int(n_votes), _p(votes), _p(y_pred))
    return votes.astype(np.int64), y_pred


def subset_eval(logits, masks):
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    masks = np.ascontiguousarray(masks, dtype=np.uint8)
    B, W, T, _ = logits.shape
    n = masks.shape[0]
    correct = np.zeros(n, dtype=np.int64)
    total = np.zeros(n, dtype=np.int64)
    lib().oracle_subset(_p(logits), ctypes.c_int64(B), ctypes.c_int(W), ctypes.c_int(T), _p(masks),
                        ctypes.c_int64(n), _p(correct), _p(total))
    return correct, total
