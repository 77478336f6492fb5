"""ctypes binding of libaau.so (include/aau.h).  No compute happens in Python: this file only marshals pointers.

The shared library is built in-tree by ``__graft_entry__.build()`` (``nvcc -gencode arch=compute_100a,code=sm_100a``)
and must sit next to this file.  There is no fallback: a missing library raises at import of the engine.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("AAU_LIB", _HERE / "libaau.so"))

AAU_VARIANT_PIPELINE, AAU_VARIANT_ABLATION = 0, 1
AAU_ACT_BF16, AAU_ACT_FP16 = 0, 1
AAU_X_F32, AAU_X_U8 = 0, 1
AAU_IN_LOGITS, AAU_IN_PROB, AAU_IN_U8, AAU_IN_LOGIT_CUT = 0, 1, 2, 3


class AauConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("in_channels", "num_classes", "base_c", "variant", "use_att", "use_aspp", "att_depth", "act_dtype")]


class AauError(RuntimeError):
    pass


_lib = None

# every symbol include/aau.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("aau_create", C.c_int, [C.POINTER(AauConfig), C.c_int, C.POINTER(C.c_void_p)]),
    ("aau_destroy", C.c_int, [C.c_void_p]),
    ("aau_last_error", C.c_char_p, [C.c_void_p]),
    ("aau_load_tensor", C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    ("aau_commit_weights", C.c_int, [C.c_void_p]),
    ("aau_missing_count", C.c_int, [C.c_void_p]),
    ("aau_unexpected_count", C.c_int, [C.c_void_p]),
    ("aau_num_keys", C.c_int, [C.c_void_p]),
    ("aau_key_name", C.c_char_p, [C.c_void_p, C.c_int]),
    ("aau_key_numel", C.c_int64, [C.c_void_p, C.c_int]),
    ("aau_workspace_bytes", C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("aau_forward", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("aau_frame_scores", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    ("aau_best_frame", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    ("aau_best_frame_mask", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("aau_sigmoid", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    ("aau_flip_w", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    ("aau_tta_prob", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    ("aau_resize_u8", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("aau_tail_masks", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("aau_condition_workspace_bytes", C.c_size_t, [C.c_void_p, C.c_int]),
    ("aau_condition_frames", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("aau_logit_cutoff", C.c_float, [C.c_float]),
    ("aau_round_window_keep_sum", C.c_int, [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_uint16)]),
    ("aau_device_fault", C.c_int, [C.c_void_p]),
    ("aau_num_launches", C.c_int, [C.c_void_p]),
    ("aau_num_ops", C.c_int, [C.c_void_p]),
    ("aau_last_forward_was_graph", C.c_int, [C.c_void_p]),
    ("aau_op_profile", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_float),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("aau_debug_tensor", C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)] + [C.POINTER(C.c_int)] * 6),
    ("aau_set_option", C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
]


def lib():
    """Load libaau.so once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise AauError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  This engine has no CPU or PyTorch fallback.")
        L = C.CDLL(str(LIB_PATH))
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(handle, status: int, what: str):
    if status != 0:
        msg = lib().aau_last_error(handle)
        raise AauError(f"{what} failed ({status}): {msg.decode() if msg else ''}")
