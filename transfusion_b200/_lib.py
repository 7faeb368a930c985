"""ctypes binding of libxfusion_sm100a.so (include/xfusion.h).  There is no fallback: if the
library is missing the import of any op fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libxfusion_sm100a.so")

_lib = None


class XfError(RuntimeError):
    pass


class XfGemm(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_ld", C.c_int64),
        ("b", C.c_void_p), ("b_ld", C.c_int64),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("tile_n", C.c_int32), ("split_k", C.c_int32),
        ("bias", C.c_void_p), ("pos_table", C.c_void_p),
        ("rows_in", C.c_int64), ("rows_out", C.c_int64), ("row_off", C.c_int64),
        ("act", C.c_int32),
        ("preact_out", C.c_void_p), ("dact_in", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("out", C.c_void_p), ("ldc", C.c_int64),
        ("out_dtype", C.c_int32), ("accumulate", C.c_int32),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint32), ("drop_stream", C.c_uint32),
        ("drop_first", C.c_int32), ("max_ctas", C.c_int32),
    ]


def lib():
    """Loads the shared library once.  Raises if it has not been built
    (``python -m transfusion_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise XfError(
                f"{LIB_PATH} not found: the CUDA library is the only implementation of this path; "
                "build it with `python -m transfusion_b200.build`")
        L = C.CDLL(LIB_PATH)
        L.xf_version.restype = C.c_int
        L.xf_last_error.restype = C.c_char_p
        L.xf_launch_count.restype = C.c_int64
        for name in EXPORTS:
            if name in ("xf_version", "xf_last_error", "xf_launch_count"):
                continue
            fn = getattr(L, name)
            fn.restype = C.c_int
        _lib = L
    return _lib


# every symbol include/xfusion.h declares (checked by tests/test_cabi.py)
EXPORTS = [
    "xf_version", "xf_last_error", "xf_launch_count", "xf_gemm",
]


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().xf_last_error().decode(errors="replace")
        raise XfError(f"{what} failed (rc={rc}): {msg}")
