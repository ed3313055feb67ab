"""ctypes binding of libmm_b200.so — the C-ABI declared in include/mm_b200.h.

This is the only place the product touches native code.  There is NO fallback: if the
library is missing or a symbol cannot be resolved, importing/using the ops raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "_lib" / "libmm_b200.so"
HEADER_PATH = _PKG.parent / "include" / "mm_b200.h"

MM_MAX_TABLES = 64
MM_I32, MM_I64, MM_F32, MM_F64 = 0, 1, 2, 3
ACTIVATIONS = {
    None: 0, "linear": 0, "relu": 1, "sigmoid": 2, "tanh": 3, "selu": 4, "elu": 5, "gelu": 6,
}
COMBINERS = {"mean": 0, "sum": 1, "sqrtn": 2, "max": 3}


class GatherTable(C.Structure):
    """mm_gather_table (include/mm_b200.h)."""

    _fields_ = [
        ("weights", C.c_void_p),
        ("indices", C.c_void_p),
        ("rows", C.c_int64),
        ("dim", C.c_int32),
        ("out_col", C.c_int32),
    ]


class LookupTable(C.Structure):
    """mm_lookup_table (include/mm_b200.h)."""

    _fields_ = [
        ("weights", C.c_void_p),
        ("indices", C.c_void_p),
        ("rows", C.c_int64),
        ("slot", C.c_int32),
        ("idx_bytes", C.c_int32),
        ("peer_weights_host", C.POINTER(C.c_void_p)),
    ]


class SparseTable(C.Structure):
    """mm_sparse_table (include/mm_b200.h)."""

    _fields_ = [
        ("weights", C.c_void_p),
        ("rows", C.c_int64),
        ("indices", C.c_void_p),
        ("idx_bytes", C.c_int32),
        ("reserved", C.c_int32),
        ("grad_rows", C.c_void_p),
        ("rep_map", C.c_void_p),
        ("state1", C.c_void_p),
        ("state2", C.c_void_p),
        ("mirror", C.c_void_p),
        ("dense_grad", C.c_void_p),
    ]


OPTIMIZERS = {"sgd": 0, "adagrad": 1, "adam": 2}
HYPER_LR, HYPER_BETA1, HYPER_BETA2, HYPER_EPS, HYPER_STEP, HYPER_LR_T, HYPER_COUNT = 0, 1, 2, 3, 4, 5, 8


class ConcatPiece(C.Structure):
    """mm_concat_piece (include/mm_b200.h)."""

    _fields_ = [
        ("src", C.c_void_p),
        ("src_stride", C.c_int64),
        ("width", C.c_int32),
        ("dtype", C.c_int32),
        ("out_col", C.c_int32),
        ("reserved", C.c_int32),
    ]


_vp, _i, _i64, _f, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64
_tables = C.POINTER(GatherTable)

# name -> (restype, argtypes); must list every function declared in include/mm_b200.h
SIGNATURES = {
    "mm_version": (_i, []),
    "mm_last_error": (C.c_char_p, []),
    "mm_launch_count": (_i64, []),
    "mm_init_uniform_hash": (_i, [_vp, _i64, _u64, _f, _f, _vp]),
    "mm_gather_multi": (_i, [_tables, _i, _i, _i64, _vp, _i64, _vp, _vp]),
    "mm_gather_bag": (_i, [_vp, _i64, _i, _vp, _i, _vp, _i, _i64, _i, _vp, _i64, _i, _vp, _vp]),
    "mm_gather_seq": (_i, [_vp, _i64, _i, _vp, _i, _i64, _i, _i, _vp, _i64, _i, _vp, _vp]),
    "mm_concat_columns": (_i, [C.POINTER(ConcatPiece), _i, _i64, _vp, _i64, _vp]),
    "mm_concat_split": (_i, [C.POINTER(ConcatPiece), _i, _i64, _vp, _i, _vp]),
    "mm_l2_normalize": (_i, [_vp, _i64, _i, _i64, _vp, _i64, _vp]),
    "mm_scale_shift": (_i, [_vp, _i64, _i, _i64, _vp, _vp, _vp, _i64, _vp]),
    "mm_cross_combine": (_i, [_vp, _vp, _vp, _i64, _i, _i64, _i64, _i64, _vp, _i64, _vp]),
    "mm_dot_interaction": (_i, [_vp, _i64, _i, _i, _i64, _vp, _i, _i64, _i, _vp, _i64, _vp, _i, _vp]),
    "mm_dlrm_gather_interact": (_i, [_tables, _i, _i, _i64, _i, _vp, _i64, _i, _vp, _i64, _vp, _i, _vp, _vp]),
    "mm_dlrm_lookup_interact": (_i, [C.POINTER(LookupTable), _i, _i64, _i, _i, _i, _vp, _i64, _i, _vp, _i64, _vp, _i, _vp, _i, _vp]),
    "mm_dense_fp32": (_i, [_vp, _i64, _i, _i64, _vp, _vp, _i, _i, _vp, _i64, _vp, _i64, _vp]),
    "mm_tc_padded_k": (_i, [_i]),
    "mm_tc_padded_n": (_i, [_i]),
    "mm_split_rows": (_i, [_vp, _i64, _i, _i64, _vp, _i, _vp]),
    "mm_split_weights": (_i, [_vp, _i, _i, _vp, _i, _i, _vp]),
    "mm_dense_tc": (_i, [_vp, _i64, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _i64, _vp, _i64,
                         _vp, _i, _vp]),
    "mm_dense_tc_head": (_i, [_vp, _i64, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _f, _i, _vp, _vp]),
    "mm_mlp_workspace_bytes": (_i64, [_i64, _i, _i, C.POINTER(C.c_int)]),
    "mm_mlp_forward": (_i, [_vp, _i64, _i, _i64, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                            C.POINTER(C.c_int), _vp, _i64, _vp, _i64, _vp]),
    "mm_cross_workspace_bytes": (_i64, [_i64, _i, _i]),
    "mm_cross_forward": (_i, [_vp, _i64, _i, _i64, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _vp, _i64, _vp, _i64, _vp]),
    "mm_mlp_tc_supported": (_i, [_i, _i, C.POINTER(C.c_int), _i]),
    "mm_mlp_tc": (_i, [_vp, _i64, _i, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                       _vp, _i64, _vp, _f, _i, _vp, _vp]),
    "mm_mlp_tc_operand_out": (_i, [_vp, _i64, _i, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_int), _vp, _i64, _vp, _vp]),
    "mm_tower2_small_supported": (_i, [_i, _i, _i]),
    "mm_tower2_small": (_i, [C.POINTER(ConcatPiece), _i, _i64, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i64, _vp, _vp]),
    "mm_rowwise_dot": (_i, [_vp, _vp, _i64, _i, _i64, _i64, _vp, _vp]),
    "mm_catalog_workspace_bytes": (_i64, [_i64, _i64, _i]),
    "mm_catalog_score": (_i, [_vp, _i64, _i, _vp, _i64, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _i64, _vp]),
    "mm_inbatch_softmax_ce": (_i, [_vp, _vp, _i64, _i64, _i, _vp, _vp, _i, _i, _f, _vp, _vp, _f, _vp, _vp, _i64, _vp]),
    "mm_shard_gather_push": (_i, [_tables, _i, _i, _i64, _i64, _i, _i, _i, C.POINTER(C.c_void_p), _i64, _vp, _vp]),
    "mm_init_uniform_hash_rows": (_i, [_vp, _i64, _i, _u64, _f, _f, _i64, _i64, _vp]),
    "mm_positive_scores": (_i, [_vp, _vp, _i64, _i, _vp, _f, _vp, _i64, _vp]),
    "mm_inbatch_scores_tc": (_i, [_vp, _vp, _i64, _i64, _i, _vp, _vp, _i, _i, _f, _vp, _f, _vp, _i64, _vp]),
    "mm_inbatch_scores": (_i, [_vp, _vp, _vp, _i64, _i64, _i, _vp, _vp, _i, _i, _f, _vp, _vp, _f,
                               _vp, _i64, _vp]),
    "mm_fm_pairwise": (_i, [_vp, _i64, _i, _i, _vp, _vp]),
    "mm_deepfm_head": (_i, [C.POINTER(LookupTable), C.POINTER(C.c_int64), _i, _i64, _i, C.POINTER(ConcatPiece), C.POINTER(C.c_int64), _i,
                            _vp, _vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp]),
    "mm_bce_head_fwd_bwd": (_i, [_vp, _i64, _i, _i64, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "mm_dense_wgrad": (_i, [_vp, _i64, _i, _i64, _vp, _i, _i64, _vp, _vp, _vp]),
    "mm_dense_wgrad_split": (_i, [_vp, _i64, _i, _i, _vp, _i, _i64, _vp, _vp, _vp]),
    "mm_dense_dgrad": (_i, [_vp, _i64, _i, _i64, _vp, _i, _vp, _i64, _vp, _i64, _vp]),
    "mm_relu_mask": (_i, [_vp, _i64, _i, _i64, _vp, _i64, _vp]),
    "mm_dlrm_interact_backward": (_i, [C.POINTER(LookupTable), _i, _i64, _i, _vp, _i64, _i, _i, _vp, _i64,
                                       C.POINTER(C.c_void_p), _i64, _vp, _i64, _i, _i, _vp]),
    "mm_sparse_rows_apply": (_i, [C.POINTER(SparseTable), _i, _i64, _i, _i, _vp, _vp]),
    "mm_dense_apply": (_i, [_i, _vp, _vp, _vp, _vp, _i64, _vp, _f, _vp]),
    "mm_opt_tick": (_i, [_vp, _vp]),
    "mm_fill_i32": (_i, [_vp, _i64, C.c_int32, _vp]),
}


def declared_symbols() -> list[str]:
    """Function names declared in include/mm_b200.h (parsed from the header text)."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mm_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every entry point; raise loudly if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m models_b200.csrc.build` "
            "(or __graft_entry__.build()). There is no CPU/PyTorch fallback for this path."
        )
    lib = C.CDLL(os.fspath(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover - only on a broken build
            raise RuntimeError(f"libmm_b200.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().mm_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = last_error()
    if rc < 0:
        raise ValueError(f"{what}: {msg} (code {rc})")
    raise RuntimeError(f"{what}: {msg} (cudaError {rc})")
