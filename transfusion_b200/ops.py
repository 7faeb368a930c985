"""Torch-facing wrappers over the C ABI (include/xfusion.h).  Torch is only used for device
memory and streams here; every op launches hand-written sm_100a kernels on the current stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import XfGemm, check, lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.XfError(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise _lib.XfError(f"{name}: expected {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int,
         a_mn_major: bool = False, b_mn_major: bool = False,
         bias: Optional[torch.Tensor] = None, pos_table: Optional[torch.Tensor] = None,
         rows_in: int = 0, rows_out: int = 0, row_off: int = 0,
         act: int = 0, preact_out: Optional[torch.Tensor] = None, dact_in: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, accumulate: bool = False, split_k: int = 1,
         tile_n: int = 0, drop_p: float = 0.0, drop_seed: int = 0, drop_stream: int = 0,
         drop_first: bool = False, max_ctas: int = 0) -> torch.Tensor:
    """out[M,N] (+)= A(MxK) @ B(NxK)^T with the fused epilogue of include/xfusion.h:XfGemm.
    a / b / out / residual are 2-D row-major (last stride 1); leading dims come from stride(0)."""
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    g = XfGemm()
    g.a, g.a_ld = a.data_ptr(), a.stride(0)
    g.b, g.b_ld = b.data_ptr(), b.stride(0)
    g.a_mn_major, g.b_mn_major = int(a_mn_major), int(b_mn_major)
    g.M, g.N, g.K = M, N, K
    g.tile_n, g.split_k = tile_n, split_k
    if bias is not None:
        _req(bias, torch.float32, "bias")
        g.bias = bias.data_ptr()
    if pos_table is not None:
        _req(pos_table, torch.float32, "pos_table")
        g.pos_table = pos_table.data_ptr()
    g.rows_in, g.rows_out, g.row_off = rows_in, rows_out, row_off
    g.act = act
    if preact_out is not None:
        _req(preact_out, torch.bfloat16, "preact_out")
        g.preact_out = preact_out.data_ptr()
    if dact_in is not None:
        _req(dact_in, torch.bfloat16, "dact_in")
        g.dact_in = dact_in.data_ptr()
    if residual is not None:
        _req(residual, torch.bfloat16, "residual")
        g.residual, g.ldr = residual.data_ptr(), residual.stride(0)
    g.out, g.ldc = out.data_ptr(), out.stride(0)
    if out.dtype == torch.float32:
        g.out_dtype = 1
    elif out.dtype == torch.bfloat16:
        g.out_dtype = 0
    else:
        raise _lib.XfError(f"out: unsupported dtype {out.dtype}")
    g.accumulate = int(accumulate)
    g.drop_p, g.drop_seed, g.drop_stream, g.drop_first = drop_p, drop_seed, drop_stream, int(drop_first)
    g.max_ctas = max_ctas
    check(lib().xf_gemm(C.byref(g), _stream()), "xf_gemm")
    return out
