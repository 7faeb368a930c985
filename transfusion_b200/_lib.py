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
        ("drop_first", C.c_int32), ("max_ctas", C.c_int32), ("cta_group", C.c_int32),
        ("batch1", C.c_int32), ("batch2", C.c_int32),
        ("a_bs1", C.c_int64), ("a_bs2", C.c_int64), ("b_bs1", C.c_int64), ("b_bs2", C.c_int64),
        ("out_bs1", C.c_int64), ("out_bs2", C.c_int64),
    ]


class XfCastJob(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("lds", C.c_int64),
        ("dst_bf16", C.c_void_p), ("ldd", C.c_int64),
        ("rows", C.c_int32), ("cols", C.c_int32), ("rin", C.c_int32), ("rout", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
    ]


XF_CAST_MAX_JOBS = 32
XF_OPT_MAX_JOBS = 32


class XfRAdamJob(C.Structure):
    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("param_bf16", C.c_void_p), ("n", C.c_int64),
    ]


class XfRAdam(C.Structure):
    _fields_ = [
        ("lr_d", C.c_double), ("beta1_d", C.c_double), ("beta2_d", C.c_double), ("weight_decay_d", C.c_double),
        ("eps", C.c_float),
        ("degenerated_to_sgd", C.c_int32),
        ("step", C.c_int64),
        ("max_grad_norm", C.c_float),
        ("grad_sqnorm", C.c_void_p),
    ]


class XfLayerNorm(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("y", C.c_void_p), ("ldy", C.c_int64),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("rows", C.c_int32), ("D", C.c_int32),
        ("in_rows_in", C.c_int32), ("in_rows_out", C.c_int32), ("in_row_off", C.c_int32),
        ("out_rows_in", C.c_int32), ("out_rows_out", C.c_int32), ("out_row_off", C.c_int32),
        ("eps", C.c_float),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint32), ("drop_stream", C.c_uint32),
    ]


class XfLayerNormBwd(C.Structure):
    _fields_ = [
        ("dy", C.c_void_p), ("lddy", C.c_int64),
        ("x", C.c_void_p), ("ldx", C.c_int64),
        ("gamma", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
        ("dx", C.c_void_p), ("lddx", C.c_int64),
        ("dx2", C.c_void_p),
        ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dbias", C.c_void_p),
        ("rows", C.c_int32), ("D", C.c_int32),
        ("in_rows_in", C.c_int32), ("in_rows_out", C.c_int32), ("in_row_off", C.c_int32),
        ("out_rows_in", C.c_int32), ("out_rows_out", C.c_int32), ("out_row_off", C.c_int32),
        ("dy_drop_p", C.c_float), ("dy_drop_seed", C.c_uint32), ("dy_drop_stream", C.c_uint32),
        ("dx2_drop_p", C.c_float), ("dx2_drop_seed", C.c_uint32), ("dx2_drop_stream", C.c_uint32),
    ]


class XfAttnFwd(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64),
        ("v", C.c_void_p), ("ldv", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("lse", C.c_void_p),
        ("lse_stride", C.c_int32),
        ("key_padding_mask", C.c_void_p),
        ("kpm_start", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("Sq", C.c_int32), ("Sk", C.c_int32), ("dp", C.c_int32),
        ("scale", C.c_float),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint32), ("drop_stream", C.c_uint32),
        ("debug_timeline", C.c_void_p),
    ]


class XfAttnBwd(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64),
        ("v", C.c_void_p), ("ldv", C.c_int64),
        ("d_out", C.c_void_p), ("lddo", C.c_int64),
        ("lse", C.c_void_p), ("delta", C.c_void_p), ("stat_stride", C.c_int32),
        ("dq", C.c_void_p), ("lddq", C.c_int64),
        ("dk", C.c_void_p), ("lddk", C.c_int64),
        ("dv", C.c_void_p), ("lddv", C.c_int64),
        ("key_padding_mask", C.c_void_p),
        ("kpm_start", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("Sq", C.c_int32), ("Sk", C.c_int32), ("dp", C.c_int32),
        ("scale", C.c_float),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint32), ("drop_stream", C.c_uint32),
        ("debug_timeline", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
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
            if name in ("xf_attn_bwd_workspace_bytes", "xf_tmap_cache_stats"):
                getattr(L, name).restype = C.c_int64
                continue
            fn = getattr(L, name)
            fn.restype = C.c_int
        _lib = L
    return _lib


# every symbol include/xfusion.h declares (checked by tests/test_cabi.py)
EXPORTS = [
    "xf_version", "xf_last_error", "xf_launch_count", "xf_tmap_cache_stats", "xf_gemm", "xf_set_gemm_cta_cap",
    "xf_patchify", "xf_fold", "xf_lang_rows_fwd", "xf_lang_rows_bwd",
    "xf_layernorm_fwd", "xf_layernorm_bwd", "xf_colsum", "xf_cast_pad", "xf_cast_pad_multi", "xf_unpad_add", "xf_bf16_to_f32",
    "xf_attn_delta", "xf_attn_fwd", "xf_attn_bwd", "xf_attn_bwd_workspace_bytes", "xf_rows_gather",
    "xf_lm_pool_fwd", "xf_lm_pool_bwd", "xf_rowln_fwd", "xf_rowln_bwd", "xf_small_linear_fwd", "xf_small_linear_bwd",
    "xf_grad_sqnorm", "xf_radam_step", "xf_split3", "xf_softmax_rows_f32",
    "xf_debug_dropout_mask", "xf_debug_attn_dropout_mask",
]


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().xf_last_error().decode(errors="replace")
        raise XfError(f"{what} failed (rc={rc}): {msg}")
