"""ctypes binding of libnrhead.so — the C ABI declared in include/nrhead.h.

There is no CPU or PyTorch-eager fallback: if the shared library is missing (and cannot be built with
nvcc) importing any op raises, and every op refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes
import os
import threading

from . import build as _build

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I = ctypes.c_int
_F = ctypes.c_float
_SZ = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/nrhead.h one to one
SIGNATURES = {
    "nr_version": (_I, []),
    "nr_last_error": (ctypes.c_char_p, []),
    "nr_device_supported": (_I, []),
    "nr_prep_partials": (_I64, [_I64]),
    "nr_prep_tokens": (_I, [_P, _I64, _I64, _P, _P, _P, _P, _P, _P]),
    "nr_prep_tokens_split": (_I, [_P, _I64, _I64, _P, _P, _I, _P, _P, _P, _P]),
    "nr_prep_tokens_bwd": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _P, _I, _P]),
    "nr_bank_advance": (_I, [_P, _I64, _I64, _P, _P, _P]),
    "nr_bank_insert": (_I, [_P, _P, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P, _I, _P, _I64, _P]),
    "nr_bank_insert_pair": (_I, [_P, _I, _I64, _I64, _I64, _P, _P]),
    "nr_matmul_f32": (_I, [_P, _I64, _I, _P, _I64, _I64, _I64, _I64, _P, _I64, _I, _P]),
    "nr_matvec_small": (_I, [_P, _I64, _I64, _I, _P, _P, _P, _P]),
    "nr_cast_bf16": (_I, [_P, _P, _I64, _P]),
    "nr_cast_bf16_multi": (_I, [_P, _P, _P, _I, _P]),
    "nr_mlp_fwd": (_I, [_P, _I64, _I64, _P, _I64, _P, _P, _P, _P, _P]),
    "nr_token_softmax": (_I, [_P, _P, _P, _P, _I64, _I64, _I64, _P, _P]),
    "nr_token_softmax_pair": (_I, [_P, _I, _P]),
    "nr_mlp_bwd_dx": (_I, [_P, _I64, _I64, _P, _I64, _P, _I, _P]),
    "nr_mlp_bwd_dw1": (_I, [_P, _I64, _I64, _P, _I64, _P, _P]),
    "nr_mlp_fwd_pair": (_I, [_P, _I, _I64, _I64, _I, _P]),
    "nr_mlp_bwd_pair": (_I, [_P, _I, _I64, _I64, _P]),
    "nr_mlp_chunks": (_I64, [_I64]),
    "nr_token_weights_fwd": (_I, [_P, _I, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _P, _P]),
    "nr_token_weights_bwd": (_I, [_P, _I, _P, _P, _P, _I64, _I64, _I64, _P, _I64, _P, _P, _P]),
    "nr_maxsim_fwd": (_I, [_I, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _F, _P, _I64, _I64, _P, _I64,
                           _I64, _I, _P, _P, _P]),
    "nr_maxsim2_supported": (_I, [_I64, _I64, _I64]),
    "nr_maxsim2_fwd": (_I, [_P, _I, _I64, _I64, _I64, _P, _P]),
    "nr_maxsim2_fwd_ex": (_I, [_P, _I, _I64, _I64, _I64, _P, _I, _P]),
    "nr_maxsim2_rank": (_I, [_P, _I, _I64, _I64, _I64, _P, _I, _P]),
    "nr_maxsim2_bwd": (_I, [_P, _I, _I64, _I64, _I64, _P]),
    "nr_maxsim2_bwd_w": (_I, [_P, _P, _P, _I64, _I64, _F, _I64, _I64, _I64, _I64, _P, _P, _P]),
    "nr_maxsim2_bwd_w_multi": (_I, [_P, _I, _I64, _I64, _P]),
    "nr_transpose_tokens_bf16": (_I, [_P, _I64, _I64, _P, _I64, _P]),
    "nr_maxsim_bwd_x": (_I, [_I, _P, _I64, _P, _P, _P, _P, _P, _I64, _I64, _F, _I64, _I64, _I64, _I64, _I64, _P, _P]),
    "nr_maxsim_bwd_y": (_I, [_I, _P, _I64, _P, _P, _P, _P, _P, _I64, _I64, _F, _I64, _I64, _I64, _I64, _I64, _P, _P]),
    "nr_maxsim_bwd_w": (_I, [_P, _P, _I64, _I64, _F, _I64, _I64, _I64, _P, _P]),
    "nr_centrality_fwd": (_I, [_P, _I64, _I64, _P, _I64, _I64, _F, _P, _P, _P, _P, _P]),
    "nr_centrality_bwd": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _F, _I64, _P, _I, _P, _P]),
    "nr_gram_f32": (_I, [_P, _P, _I64, _I64, _I64, _P, _P, _P]),
    "nr_row_losses_fwd": (_I, [_P, _I64, _P, _I64, _P, _P, _P, _P, _I64, _I64, _I64, _P, _I, _F, _F, _F, _I, _P,
                               _P, _P, _P]),
    "nr_row_losses_bwd": (_I, [_P, _I64, _P, _I64, _P, _P, _P, _P, _I64, _I64, _I64, _P, _I, _F, _F, _F, _I, _P,
                               _P, _P, _P, _I64, _P, _I64, _P, _P, _P, _P]),
    "nr_row_mean": (_I, [_P, _I64, _I64, _I64, _P, _P]),
    "nr_vec_sums": (_I, [_P, _I64, _I64, _P, _P, _P]),
    "nr_transpose_add": (_I, [_P, _I64, _P, _I64, _P, _I64, _I64, _I64, _F, _F, _P]),
    "nr_sinkhorn_workspace_bytes": (_SZ, [_I64]),
    "nr_sinkhorn": (_I, [_P, _P, _I64, _I, _P, _P, _P, _P, _P, _SZ, _P]),
    "nr_sinkhorn_ex": (_I, [_P, _P, _I64, _I, _P, _P, _P, _P, _P, _SZ, _I, _P]),
    "nr_fifo_update": (_I, [_P, _I64, _P, _I64, _P, _I64, _I64, _P]),
    "nr_rank_count": (_I, [_P, _I64, _I64, _I64, _P, _I64, _P, _P, _P]),
    "nr_topk_rows": (_I, [_P, _I64, _I64, _I64, _I, ctypes.c_int32, _P, _P, _P]),
    "nr_topk_merge": (_I, [_P, _P, _I64, _I64, _I, _P, _P, _P]),
    "nr_rank_count_target": (_I, [_P, _I64, _I64, _I64, _P, _P, _I64, _P, _P, _P, _P]),
    "nr_group_max_t": (_I, [_P, _I64, _I64, _I64, _P, _I64, _P, _I64, _P]),
}



class MaxSim2Problem(ctypes.Structure):
    """nr_maxsim2_problem of include/nrhead.h (field order and types must match)."""
    _fields_ = [("x_bf16", _P), ("y_bf16", _P), ("wx", _P), ("wy", _P), ("Rx", _I64), ("Ry", _I64), ("alpha", _F),
                ("out", _P), ("out_sr", _I64), ("out_sc", _I64), ("out2", _P), ("out2_sr", _I64), ("out2_sc", _I64),
                ("pmax_x", _P), ("ystar", _P), ("pmax_y", _P), ("xstar", _P)]


class MaxSim2RankProblem(ctypes.Structure):
    """nr_maxsim2_rank_problem of include/nrhead.h (field order and types must match)."""
    _fields_ = [("x_bf16", _P), ("y_bf16", _P), ("wx", _P), ("wy", _P), ("Rx", _I64), ("Ry", _I64), ("alpha", _F),
                ("gx0", _I64), ("gy0", _I64), ("diag", _P), ("gt_x", _P), ("eq_x", _P), ("gt_y", _P), ("eq_y", _P)]


class SoftmaxSide(ctypes.Structure):
    """nr_softmax_side of include/nrhead.h (field order and types must match)."""
    _fields_ = [("logits", _P), ("b2", _P), ("mask_a", _P), ("mask_b", _P), ("Ra", _I64), ("R", _I64), ("N", _I64), ("w", _P)]


class MaxSim2BwdWJob(ctypes.Structure):
    """nr_maxsim2_bwd_w_job of include/nrhead.h (field order and types must match)."""
    _fields_ = [("pmax_x", _P), ("pmax_y", _P), ("dH", _P), ("dh_sr", _I64), ("dh_sc", _I64), ("dh_scale", _F),
                ("Rx", _I64), ("Ry", _I64), ("dwx", _P), ("dwy", _P)]


class BankSide(ctypes.Structure):
    """nr_bank_side of include/nrhead.h (field order and types must match)."""
    _fields_ = [("new_feat", _P), ("new_mask", _P), ("N", _I64), ("ring_feat", _P), ("ring_mask", _P), ("ring_raw_bf16", _P),
                ("ring_xn_bf16", _P), ("split_role", _I), ("ring_xnT_bf16", _P), ("ld", _I64)]


class MlpSide(ctypes.Structure):
    """nr_mlp_side of include/nrhead.h (field order and types must match)."""
    _fields_ = [("x_bf16", _P), ("w1_bf16", _P), ("T", _I64), ("b1", _P), ("w2", _P), ("h_bf16", _P), ("logits", _P),
                ("dh_bf16", _P), ("dw1", _P), ("dx", _P), ("T_dx", _I64)]


class MaxSim2BwdJob(ctypes.Structure):
    """nr_maxsim2_bwd_job of include/nrhead.h (field order and types must match)."""
    _fields_ = [("side", _I), ("srcT", _P), ("src_ld", _I64), ("wx", _P), ("wy", _P), ("ystar", _P), ("xstar", _P),
                ("dH", _P), ("dh_sr", _I64), ("dh_sc", _I64), ("dh_scale", _F), ("Rx", _I64), ("Ry", _I64), ("dst", _P),
                ("part", _I)]


NR_LOSS_CENTRALITY, NR_LOSS_NEIGHBOR, NR_LOSS_KL, NR_LOSS_UNIFORM = 1, 2, 4, 8
NR_NSAVE = 16
NR_PREC_FP32, NR_PREC_BF16, NR_PREC_BF16X3 = 0, 1, 2

_lock = threading.Lock()
_lib = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load (building first if needed) libnrhead.so; raises if the CUDA extension is unavailable."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            if not build_if_missing:
                raise RuntimeError(f"{path} is missing: run `python -m neighborretr_b200.build`")
            _build.build()
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc, name):
    if rc != 0:
        msg = load().nr_last_error().decode(errors="replace")
        raise RuntimeError(f"{name} failed (rc={rc}): {msg}")
