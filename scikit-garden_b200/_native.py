"""ctypes binding of the C ABI declared in ``include/qforest_b200.h``.

The shared library is the product: if it is missing the import fails loudly with
instructions to build it -- there is no Python or CPU fallback behind these calls.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqforest_b200.so")

QF_OK = 0
QF_ERR_INVALID = -1
QF_ERR_CUDA = -2
QF_ERR_LIMIT = -3
QF_ERR_WORKSPACE = -4
QF_MODE_EXACT = 0
QF_MODE_FAST = 1
ABI_VERSION = 2


class QFError(RuntimeError):
    """A C-ABI call returned a negative qf_status."""

    def __init__(self, code, message):
        super().__init__("%s (qf_status %d)" % (message, code))
        self.code = code


class ForestDesc(C.Structure):
    """struct qf_forest_desc"""
    _fields_ = [
        ("device", C.c_int32), ("n_trees", C.c_int32), ("n_features", C.c_int32),
        ("feature_bits", C.c_int32),
        ("n_train", C.c_int64), ("n_nodes", C.c_int64), ("n_members", C.c_int64),
        ("nodes", C.c_void_p), ("tree_offsets", C.c_void_p), ("node_ids", C.c_void_p),
        ("members", C.c_void_p), ("y_sorted", C.c_void_p), ("leaf_values", C.c_void_p),
        ("max_row_members", C.c_int64),
        ("expected_row_members", C.c_double),
    ]


class CsrBuildDesc(C.Structure):
    """struct qf_csr_build_desc"""
    _fields_ = [
        ("device", C.c_int32), ("n_trees", C.c_int32), ("n_features", C.c_int32),
        ("feature_bits", C.c_int32),
        ("n_train", C.c_int64), ("n_nodes", C.c_int64),
        ("nodes_dev", C.c_void_p), ("tree_offsets", C.c_void_p), ("X_train_dev", C.c_void_p),
        ("rank_dev", C.c_void_p), ("counts_dev", C.c_void_p), ("members_dev", C.c_void_p),
        ("members_capacity", C.c_int64),
        ("inbag_out", C.c_void_p), ("max_leaf_members_out", C.c_void_p), ("sum_sq_out", C.c_void_p),
    ]


# every symbol include/qforest_b200.h declares: name -> (restype, argtypes)
_P, _I32, _I64, _U64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
SIGNATURES = {
    "qf_abi_version": (C.c_int, []),
    "qf_last_error": (C.c_char_p, []),
    "qf_pack_tree": (C.c_int, [_I64, _P, _P, _P, _P, _I32, _I64, _P, _P, _P, _U64,
                               _P, _P, _P, _P, _P, _P, _P]),
    "qf_count_inbag": (_I64, [_I64, _P]),
    "qf_build_csr": (C.c_int, [C.POINTER(CsrBuildDesc), C.POINTER(C.c_int64)]),
    "qf_forest_create": (C.c_int, [C.POINTER(ForestDesc), C.POINTER(C.c_void_p)]),
    "qf_forest_destroy": (None, [_P]),
    "qf_forest_device_bytes": (_I64, [_P]),
    "qf_forest_workspace_bytes": (_I64, [_P, _I64, _I32]),
    "qf_forest_apply": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _I64, _P]),
    "qf_forest_predict_quantiles": (C.c_int, [_P, _P, _I64, _I64, _P, _I32, _P, _I32, _P, _I64, _P]),
    "qf_forest_predict_mean": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _I64, _P]),
    "qf_forest_set_leaf_impurity": (C.c_int, [_P, _P]),
    "qf_forest_predict_mean_std": (C.c_int, [_P, _P, _I64, _I64, C.c_double, _P, _P, _P, _I64, _P]),
    "qf_forest_predict_quantiles_host": (C.c_int, [_P, _P, _I64, _I64, _P, _I32, _P, _I32, _I64]),
    "qf_forest_plan": (C.c_int, [_P, _I64, _P]),
    "qf_forest_quantile_kernel_name": (C.c_int, [_P, _I64, _I32, _I32, C.c_char_p, _I32]),
    "qf_forest_last_launches": (C.c_int, [_P]),
    "qf_forest_last_timings": (C.c_int, [_P, _P, _P]),
    "qf_forest_set_option": (C.c_int, [_P, C.c_char_p, _I64]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle of libqforest_b200.so."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libqforest_b200.so is not built (%s). Build it with "
            "`python scikit-garden_b200/build.py` (needs nvcc); this package has no "
            "CPU fallback." % LIB_PATH)
    handle = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(handle, name)          # AttributeError = symbol missing: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    if handle.qf_abi_version() != ABI_VERSION:
        raise ImportError("libqforest_b200.so has ABI %d, expected %d; rebuild it"
                          % (handle.qf_abi_version(), ABI_VERSION))
    _lib = handle
    return _lib


def last_error():
    msg = lib().qf_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code):
    if code != QF_OK:
        raise QFError(code, last_error())
    return code
