"""ctypes binding of libcpros.so (include/cpros.h) -- the only way the package computes anything.

There is NO CPU fallback: if the shared library is missing or a tensor is not a CUDA tensor the
call raises.  The library is built in-tree by `build()` (nvcc, sm_100a) so that it travels with
the repository snapshot to the GPU box.
"""
import ctypes
import os
import subprocess

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
# CP_LIBCPROS: another build of the same sources (A/B measurements of kernel variants); default: the in-tree library
SO_PATH = os.environ.get("CP_LIBCPROS") or os.path.join(_PKG, "libcpros.so")
SOURCES = ["version.cu", "gather.cu", "encoder.cu", "head.cu", "clip.cu", "vote.cu", "l2.cu", "step.cu", "preprocess.cu", "philox.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

N_BN, N_FC = 9, 7
BN_BATCH, BN_BATCH_UPDATE, BN_RUNNING = 0, 1, 2
ENGINE_SIMT, ENGINE_TC, ENGINE_TC_FP16 = 0, 1, 2


def _stale():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + \
           [os.path.join(os.path.dirname(_PKG), "include", "cpros.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into contrastiveprosthetics_b200/libcpros.so."""
    if not force and not _stale():
        return SO_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-o", SO_PATH] + [os.path.join(_CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO_PATH


class EncoderTensors(ctypes.Structure):
    _fields_ = [("conv1_w", ctypes.c_void_p), ("conv1_b", ctypes.c_void_p),
                ("conv2_w", ctypes.c_void_p), ("conv2_b", ctypes.c_void_p),
                ("fc_w", ctypes.c_void_p * N_FC), ("fc_b", ctypes.c_void_p * N_FC),
                ("proj_w", ctypes.c_void_p),
                ("bn_w", ctypes.c_void_p * N_BN), ("bn_b", ctypes.c_void_p * N_BN),
                ("bn_rm", ctypes.c_void_p * N_BN), ("bn_rv", ctypes.c_void_p * N_BN)]


class EncoderOpts(ctypes.Structure):
    _fields_ = [("bn_mode", ctypes.c_int32), ("engine", ctypes.c_int32),
                ("bn_momentum", ctypes.c_float), ("bn_eps", ctypes.c_float),
                ("dropout_p", ctypes.c_float), ("save_for_backward", ctypes.c_int32),
                ("dropout_seed", ctypes.c_uint64), ("ext_masks", ctypes.c_void_p), ("dropout_step", ctypes.c_void_p),
                ("allreduce", ctypes.c_void_p), ("allreduce_user", ctypes.c_void_p),
                ("trunk_only", ctypes.c_int32), ("reserved", ctypes.c_int32)]


# int (*cp_allreduce_fn)(void *user, void *buf, size_t count, void *stream)   (SyncBN hook, include/cpros.h)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)


class ClsTensors(ctypes.Structure):
    _fields_ = [("w1", ctypes.c_void_p), ("b1", ctypes.c_void_p), ("bn_w", ctypes.c_void_p), ("bn_b", ctypes.c_void_p),
                ("bn_rm", ctypes.c_void_p), ("bn_rv", ctypes.c_void_p), ("w2", ctypes.c_void_p)]


class GloveTensors(ctypes.Structure):
    _fields_ = [("w0", ctypes.c_void_p), ("bn0_w", ctypes.c_void_p), ("bn0_b", ctypes.c_void_p),
                ("w", ctypes.c_void_p * 3), ("b", ctypes.c_void_p * 3),
                ("bn_w", ctypes.c_void_p * 3), ("bn_b", ctypes.c_void_p * 3), ("proj_w", ctypes.c_void_p)]


class GloveOpts(ctypes.Structure):
    _fields_ = [("glove_dim", ctypes.c_int32), ("save_for_backward", ctypes.c_int32),
                ("bn_eps", ctypes.c_float), ("dropout_p", ctypes.c_float),
                ("dropout_seed", ctypes.c_uint64), ("ext_masks", ctypes.c_void_p), ("dropout_step", ctypes.c_void_p)]


_lib = None
_vp, _i64, _i32, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t


def lib():
    """Load libcpros.so; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). contrastiveprosthetics_b200 has no CPU or PyTorch fallback.")
    L = ctypes.CDLL(SO_PATH)
    L.cp_version.restype = ctypes.c_int
    L.cp_launch_count.restype = ctypes.c_ulonglong
    L.cp_status_string.restype = ctypes.c_char_p
    L.cp_status_string.argtypes = [ctypes.c_int]
    L.cp_gather_norm.argtypes = [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp]
    L.cp_encoder_workspace_bytes.restype = _sz
    L.cp_encoder_workspace_bytes.argtypes = [_i64, ctypes.POINTER(EncoderOpts)]
    L.cp_encoder_forward.argtypes = [ctypes.POINTER(EncoderTensors), _vp, _i64, _vp, _vp, _sz,
                                     ctypes.POINTER(EncoderOpts), _vp]
    L.cp_encoder_backward.argtypes = [ctypes.POINTER(EncoderTensors), _vp, _i64,
                                      ctypes.POINTER(EncoderTensors), _vp, _sz,
                                      ctypes.POINTER(EncoderOpts), _vp]
    L.cp_encoder_read_activation.argtypes = [_vp, _sz, _i64, ctypes.POINTER(EncoderOpts), _i32, _i32, _vp, _vp]
    L.cp_encoder_read_activation.restype = ctypes.c_int
    L.cp_philox4x32_10.argtypes = [_vp, _vp, _i64, _vp, _vp]
    L.cp_philox4x32_10.restype = ctypes.c_int
    L.cp_dropout_mask.argtypes = [_vp, _i64, ctypes.c_float, ctypes.c_uint64, _i32, _vp, _vp]
    L.cp_dropout_mask.restype = ctypes.c_int
    L.cp_linear_workspace_bytes.restype = _sz
    L.cp_linear_workspace_bytes.argtypes = [_i64, _i32, _i32]
    L.cp_linear_forward.argtypes = [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _i32, _vp]
    L.cp_split_planes.argtypes = [_vp, _vp, _vp, _i64, _vp]
    L.cp_split_planes.restype = ctypes.c_int
    L.cp_linear_forward_planes.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]
    L.cp_linear_forward_planes.restype = ctypes.c_int
    L.cp_linear_backward.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _sz, _i32, _vp]
    L.cp_head_workspace_bytes.restype = _sz
    L.cp_head_workspace_bytes.argtypes = [_i64]
    L.cp_head_forward_backward.argtypes = [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                           _vp, _sz, _vp]
    L.cp_logits_loss.argtypes = [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]
    _f = ctypes.c_float
    L.cp_clip_normalize.argtypes = [_vp, _i64, _vp, _vp, _vp]
    L.cp_clip_transpose.argtypes = [_vp, _i64, _i64, _vp, _vp]
    L.cp_clip_sums.argtypes = [_vp, _i64, _vp, _i64, _i64, _f, _vp, _vp, _vp]
    L.cp_clip_loss.argtypes = [_vp, _vp, _vp, _vp, _i64, _i64, _f, _vp, _i64, _vp, _vp, _vp]
    L.cp_clip_grad.argtypes = [_vp, _i64, _vp, _i64, _i64, _f, _vp, _vp, _f, _vp, _vp]
    L.cp_clip_embed_backward.argtypes = [_vp, _vp, _vp, _vp, _i64, _f, _vp, _vp]
    L.cp_glove_workspace_bytes.restype = _sz
    L.cp_glove_workspace_bytes.argtypes = [_i64, ctypes.POINTER(GloveOpts)]
    L.cp_glove_forward.argtypes = [ctypes.POINTER(GloveTensors), _vp, _i64, _vp, _vp, _sz, ctypes.POINTER(GloveOpts), _vp]
    L.cp_glove_backward.argtypes = [ctypes.POINTER(GloveTensors), _vp, _i64, ctypes.POINTER(GloveTensors), _vp, _sz,
                                    ctypes.POINTER(GloveOpts), _vp]
    L.cp_cls_workspace_bytes.restype = _sz
    L.cp_cls_workspace_bytes.argtypes = [_i64]
    L.cp_cls_forward_backward.argtypes = [ctypes.POINTER(ClsTensors), _vp, _vp, _i64, _i32, ctypes.c_float, ctypes.c_float,
                                          _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(ClsTensors), _vp, _sz, _vp]
    L.cp_cls_forward_backward.restype = ctypes.c_int
    L.cp_vote_eval.argtypes = [_vp, _i64, _i32, _i32, _vp, _vp, _vp]
    L.cp_rank_rows.argtypes = [_vp, _i64, _vp, _vp]
    L.cp_subset_eval.argtypes = [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp]
    L.cp_confusion_matrix.argtypes = [_vp, _vp, _i64, _i32, _vp, _vp, _vp]
    L.cp_l2_workspace_bytes.restype = _sz
    L.cp_l2_workspace_bytes.argtypes = [_i32]
    L.cp_l2_forward.argtypes = [_vp, _vp, _i32, _vp, _vp, _vp, _sz, _vp]
    L.cp_l2_backward.argtypes = [_vp, _vp, _i32, _vp, _vp, ctypes.c_float, _vp, _vp]
    _d = ctypes.c_double
    L.cp_step_workspace_bytes.restype = _sz
    L.cp_step_workspace_bytes.argtypes = [_i32]
    L.cp_step_prologue.argtypes = [_vp, _vp, _i32, _vp, _vp, _i32, _vp, _sz, _vp]
    L.cp_adam_step.argtypes = [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _vp]
    L.cp_emg_preprocess_scratch_elems.restype = _sz
    L.cp_emg_preprocess_scratch_elems.argtypes = [_i64, _i32, _i32]
    L.cp_emg_preprocess.argtypes = [_vp, _i64, _i32, _i32, _vp, _vp, _i32, ctypes.c_float, _i32, _i32, _vp, _i32, _vp,
                                    _vp, _sz, _vp]
    for name in ("cp_gather_norm", "cp_encoder_forward", "cp_encoder_backward", "cp_linear_forward",
                 "cp_linear_backward", "cp_head_forward_backward", "cp_logits_loss", "cp_vote_eval",
                 "cp_rank_rows", "cp_subset_eval", "cp_clip_normalize", "cp_clip_transpose", "cp_clip_sums",
                 "cp_clip_loss", "cp_clip_grad", "cp_clip_embed_backward", "cp_glove_forward", "cp_glove_backward",
                 "cp_confusion_matrix", "cp_l2_forward", "cp_l2_backward", "cp_emg_preprocess", "cp_step_prologue",
                 "cp_adam_step"):
        getattr(L, name).restype = ctypes.c_int
    _lib = L
    return L


EXPORTS = ["cp_version", "cp_launch_count", "cp_status_string", "cp_gather_norm", "cp_encoder_workspace_bytes",
           "cp_encoder_forward", "cp_encoder_backward", "cp_encoder_read_activation", "cp_linear_workspace_bytes",
           "cp_linear_forward", "cp_linear_backward", "cp_split_planes", "cp_linear_forward_planes", "cp_head_workspace_bytes",
           "cp_head_forward_backward", "cp_logits_loss", "cp_vote_eval", "cp_rank_rows",
           "cp_subset_eval", "cp_clip_normalize", "cp_clip_transpose", "cp_clip_sums", "cp_clip_loss",
           "cp_clip_grad", "cp_clip_embed_backward", "cp_glove_workspace_bytes", "cp_glove_forward",
           "cp_glove_backward", "cp_confusion_matrix", "cp_l2_workspace_bytes", "cp_l2_forward", "cp_l2_backward",
           "cp_emg_preprocess_scratch_elems", "cp_emg_preprocess", "cp_philox4x32_10", "cp_dropout_mask", "cp_cls_workspace_bytes",
           "cp_cls_forward_backward", "cp_step_workspace_bytes", "cp_step_prologue", "cp_adam_step"]


def check(status, what=""):
    if status != 0:
        msg = lib().cp_status_string(status).decode()
        raise RuntimeError(f"libcpros {what}: status {status}: {msg}")


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL).  Refuses CPU tensors."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libcpros takes CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("libcpros needs contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
