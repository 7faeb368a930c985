"""Torch-facing wrappers over the C ABI (include/xfusion.h).  Torch is only used for device
memory and streams here; every op launches hand-written sm_100a kernels on the current stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import XfAttnBwd, XfAttnFwd, XfGemm, XfLayerNorm, XfLayerNormBwd, check, lib


# GEMM tiling default: 0 = library default (CTA pairs), 1 = single-CTA tiles (XF_GEMM_CTA_GROUP env var)
import os as _os
DEFAULT_CTA_GROUP = int(_os.environ.get("XF_GEMM_CTA_GROUP", "0"))

# When set to a list, every op appends (family, algorithmic flops, algorithmic bytes, start event, end event):
# bench.py's per-kernel roofline pass (CUDA events on the launching stream).
PROFILE = None
PROFILE_DETAIL = bool(int(_os.environ.get("XF_PROFILE_DETAIL", "0")))


# dev aid (tools/ablate.sh): XF_ABLATE="fam1,fam2" turns the launches of those op families into no-ops, so the difference in
# step time is the family's TRUE marginal cost inside the overlapped multi-stream schedule (results are garbage, of course)
ABLATE = frozenset(x for x in _os.environ.get("XF_ABLATE", "").split(",") if x)


class _Skip(Exception):
    pass


class _Prof:
    __slots__ = ("fam", "flops", "nbytes", "e0")

    def __init__(self, fam, flops=0.0, nbytes=0.0):
        self.fam, self.flops, self.nbytes = fam, flops, nbytes
        self.e0 = None
        if ABLATE and fam.split("[")[0] in ABLATE:
            raise _Skip()

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None and self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.fam, self.flops, self.nbytes, self.e0, e1))
        return False


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
         drop_first: bool = False, max_ctas: int = 0, cta_group: int = 0,
         batch=None, a_ld: int = 0, b_ld: int = 0, ldc: int = 0) -> torch.Tensor:
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
    g.cta_group = cta_group if cta_group else DEFAULT_CTA_GROUP
    nbatch = 1
    if batch is not None:
        # batch = (n1, n2, (a_bs1, a_bs2), (b_bs1, b_bs2), (out_bs1, out_bs2)): element strides per batch index; a / b / out
        # are then only base pointers and the leading dimensions come from a_ld / b_ld / ldc
        g.batch1, g.batch2 = batch[0], batch[1]
        (g.a_bs1, g.a_bs2), (g.b_bs1, g.b_bs2), (g.out_bs1, g.out_bs2) = batch[2], batch[3], batch[4]
        nbatch = batch[0] * max(1, batch[1])
    if a_ld:
        g.a_ld = a_ld
    if b_ld:
        g.b_ld = b_ld
    if ldc:
        g.ldc = ldc
    fam = "gemm_wgrad" if a_mn_major else ("gemm_dgrad" if b_mn_major else "gemm_fwd")
    if PROFILE_DETAIL:
        fam += f"[M={M},N={N},K={K},split={split_k},act={act},drop={int(drop_p > 0)},res={int(residual is not None)}]"
    with _Prof(fam, 2.0 * M * N * K * nbatch):
        check(lib().xf_gemm(C.byref(g), _stream()), "xf_gemm")
    return out


def set_gemm_cta_cap(ctas: int) -> int:
    """Caps the persistent grid of the GEMMs issued by this host thread (0 = no cap); returns the previous cap."""
    return int(lib().xf_set_gemm_cta_cap(int(ctas)))


def _feat_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return 1
    if t.dtype == torch.bfloat16:
        return 0
    raise _lib.XfError(f"feature map dtype {t.dtype} unsupported (fp32 or bf16)")


def _feat_layout(feat: torch.Tensor, what: str) -> int:
    """0 = contiguous NCHW, 4 = channels_last (NHWC memory); anything else is rejected (no hidden conversion copies)."""
    if feat.is_contiguous():
        return 0
    if feat.dim() == 4 and feat.is_contiguous(memory_format=torch.channels_last):
        return 4
    raise _lib.XfError(f"{what}: feature map must be contiguous NCHW or channels_last")


def patchify(feat: torch.Tensor, p: int, tok: torch.Tensor) -> torch.Tensor:
    """feat [B,C,H,W] (contiguous NCHW or channels_last, fp32/bf16) -> tok bf16 [B*n, C*p*p] (utils.py:35-39 order)."""
    B, Cc, H, W = feat.shape
    layout = _feat_layout(feat, "patchify")
    _req(tok, torch.bfloat16, "tok")
    with _Prof("patchify_fold", 0.0, feat.numel() * (feat.element_size() + 2.0)):
        check(lib().xf_patchify(_ptr(feat), _feat_dtype(feat) | layout, _ptr(tok), C.c_int64(tok.stride(0)), B, Cc, H, W, p, _stream()),
              "xf_patchify")
    return tok


def fold(tok: torch.Tensor, feat: torch.Tensor, p: int, accumulate: bool = False) -> torch.Tensor:
    """tok bf16 [B*n, C*p*p] -> feat [B,C,H,W] (utils.py:42-46)."""
    B, Cc, H, W = feat.shape
    layout = _feat_layout(feat, "fold")
    _req(tok, torch.bfloat16, "tok")
    with _Prof("patchify_fold", 0.0, feat.numel() * (feat.element_size() + 2.0)):
        check(lib().xf_fold(_ptr(tok), C.c_int64(tok.stride(0)), _ptr(feat), _feat_dtype(feat) | layout, int(accumulate), B, Cc, H, W, p,
                            _stream()), "xf_fold")
    return feat


def lang_rows_fwd(lang: torch.Tensor, kind: torch.Tensor, z: torch.Tensor, n: int):
    B, L, D = lang.shape
    S = z.shape[1]
    _req(lang, torch.float32, "lang"); _req(kind, torch.float32, "kind"); _req(z, torch.bfloat16, "z")
    with _Prof("lang_rows", 0.0, 6.0 * B * L * D):
        check(lib().xf_lang_rows_fwd(_ptr(lang), _ptr(kind), _ptr(z), B, L, D, n, S, _stream()), "xf_lang_rows_fwd")


def lang_rows_bwd(dz: torch.Tensor, dlang: Optional[torch.Tensor], dkind: torch.Tensor, B: int, L: int, n: int):
    S, D = dz.shape[1], dz.shape[2]
    _req(dz, torch.bfloat16, "dz"); _req(dkind, torch.float32, "dkind")
    with _Prof("lang_rows", 0.0, 10.0 * B * L * D):
        check(lib().xf_lang_rows_bwd(_ptr(dz), _ptr(dlang), _ptr(dkind), B, L, D, n, S, _stream()), "xf_lang_rows_bwd")


def layernorm_fwd(x, y, gamma, beta, mean, rstd, rows: int, D: int, *, in_map=(0, 0, 0), out_map=(0, 0, 0),
                  eps: float = 1e-5, drop_p: float = 0.0, drop_seed: int = 0, drop_stream: int = 0):
    a = XfLayerNorm()
    a.x, a.ldx = x.data_ptr(), x.stride(-2)
    a.y, a.ldy = y.data_ptr(), y.stride(-2)
    a.gamma, a.beta = gamma.data_ptr(), beta.data_ptr()
    a.mean = mean.data_ptr() if mean is not None else None
    a.rstd = rstd.data_ptr() if rstd is not None else None
    a.rows, a.D = rows, D
    a.in_rows_in, a.in_rows_out, a.in_row_off = in_map
    a.out_rows_in, a.out_rows_out, a.out_row_off = out_map
    a.eps = eps
    a.drop_p, a.drop_seed, a.drop_stream = drop_p, drop_seed, drop_stream
    with _Prof("layernorm_fwd", 0.0, 4.0 * rows * D):
        check(lib().xf_layernorm_fwd(C.byref(a), _stream()), "xf_layernorm_fwd")


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows: int, D: int, *, dbias=None, dx2=None,
                  in_map=(0, 0, 0), out_map=(0, 0, 0), dy_drop=(0.0, 0, 0), dx2_drop=(0.0, 0, 0)):
    a = XfLayerNormBwd()
    a.dy, a.lddy = dy.data_ptr(), dy.stride(-2)
    a.x, a.ldx = x.data_ptr(), x.stride(-2)
    a.gamma, a.mean, a.rstd = gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr()
    a.dx, a.lddx = dx.data_ptr(), dx.stride(-2)
    a.dx2 = dx2.data_ptr() if dx2 is not None else None
    a.dgamma, a.dbeta = dgamma.data_ptr(), dbeta.data_ptr()
    a.dbias = dbias.data_ptr() if dbias is not None else None
    a.rows, a.D = rows, D
    a.in_rows_in, a.in_rows_out, a.in_row_off = in_map
    a.out_rows_in, a.out_rows_out, a.out_row_off = out_map
    a.dy_drop_p, a.dy_drop_seed, a.dy_drop_stream = dy_drop
    a.dx2_drop_p, a.dx2_drop_seed, a.dx2_drop_stream = dx2_drop
    with _Prof("layernorm_bwd", 0.0, (6.0 if dx2 is None else 8.0) * rows * D):
        check(lib().xf_layernorm_bwd(C.byref(a), _stream()), "xf_layernorm_bwd")


def colsum(x: torch.Tensor, out: torch.Tensor, rows: int, cols: int):
    _req(x, torch.bfloat16, "x"); _req(out, torch.float32, "out")
    with _Prof("colsum", 0.0, 2.0 * rows * cols):
        check(lib().xf_colsum(_ptr(x), C.c_int64(x.stride(-2)), rows, cols, _ptr(out), _stream()), "xf_colsum")


def cast_pad(src: torch.Tensor, dst: torch.Tensor, rows: int, cols: int, rin=0, rout=0, cin=0, cout=0):
    _req(src, torch.float32, "src"); _req(dst, torch.bfloat16, "dst")
    with _Prof("cast", 0.0, 6.0 * rows * cols):
      check(lib().xf_cast_pad(_ptr(src), C.c_int64(src.stride(0) if src.dim() > 1 else cols), _ptr(dst),
                            C.c_int64(dst.stride(0) if dst.dim() > 1 else cols), rows, cols, rin, rout, cin, cout, _stream()),
          "xf_cast_pad")


def cast_pad_multi(jobs):
    """jobs: iterable of (src fp32, dst bf16, rows, cols, rin, rout, cin, cout) -- one launch per 32 tensors."""
    jobs = list(jobs)
    for i in range(0, len(jobs), _lib.XF_CAST_MAX_JOBS):
        chunk = jobs[i:i + _lib.XF_CAST_MAX_JOBS]
        arr = (_lib.XfCastJob * len(chunk))()
        nbytes = 0.0
        for j, (src, dst, rows, cols, rin, rout, cin, cout) in zip(arr, chunk):
            _req(src, torch.float32, "src"); _req(dst, torch.bfloat16, "dst")
            j.src, j.lds = src.data_ptr(), (src.stride(0) if src.dim() > 1 else cols)
            j.dst_bf16, j.ldd = dst.data_ptr(), (dst.stride(0) if dst.dim() > 1 else cols)
            j.rows, j.cols, j.rin, j.rout, j.cin, j.cout = rows, cols, rin, rout, cin, cout
            nbytes += 6.0 * rows * cols
        with _Prof("cast", 0.0, nbytes):
            check(lib().xf_cast_pad_multi(arr, len(chunk), _stream()), "xf_cast_pad_multi")


def bf16_to_f32(src: torch.Tensor, dst: torch.Tensor, scale: float = 1.0):
    """dst (fp32, flat) = scale * src (bf16, flat); numel % 8 == 0."""
    _req(src, torch.bfloat16, "src"); _req(dst, torch.float32, "dst")
    n = src.numel()
    with _Prof("cast", 0.0, 6.0 * n):
        check(lib().xf_bf16_to_f32(_ptr(src), _ptr(dst), C.c_int64(n), C.c_float(scale), _stream()), "xf_bf16_to_f32")


def unpad_add(src: torch.Tensor, dst: torch.Tensor, rows: int, cols: int, rin=0, rout=0, cin=0, cout=0):
    _req(src, torch.float32, "src"); _req(dst, torch.float32, "dst")
    with _Prof("cast", 0.0, 12.0 * rows * cols):
      check(lib().xf_unpad_add(_ptr(src), C.c_int64(src.stride(0) if src.dim() > 1 else cols), _ptr(dst),
                             C.c_int64(dst.stride(0) if dst.dim() > 1 else cols), rows, cols, rin, rout, cin, cout, _stream()),
          "xf_unpad_add")


def attn_delta(o: torch.Tensor, d_o: torch.Tensor, delta: torch.Tensor, B: int, S: int, heads: int, dp: int):
    """delta [B, heads, stat_stride] fp32 (stat_stride = delta.shape[-1])."""
    with _Prof("attn_delta", 0.0, 4.0 * B * S * heads * dp):
        check(lib().xf_attn_delta(_ptr(o), _ptr(d_o), C.c_int64(o.stride(-2)), B, S, heads, dp, delta.shape[-1], _ptr(delta),
                                  _stream()), "xf_attn_delta")


def attn_fwd(q, k, v, out, lse, *, B: int, H: int, Sq: int, Sk: int, dp: int, scale: float,
             key_padding_mask: Optional[torch.Tensor] = None, kpm_start: int = 0,
             drop_p: float = 0.0, drop_seed: int = 0, drop_stream: int = 0, debug_timeline: Optional[torch.Tensor] = None):
    """q/k/v/out: bf16 2-D views [B*S, >=H*dp] (may be column slices of one fused qkv buffer)."""
    a = XfAttnFwd()
    a.q, a.ldq = q.data_ptr(), q.stride(0)
    a.k, a.ldk = k.data_ptr(), k.stride(0)
    a.v, a.ldv = v.data_ptr(), v.stride(0)
    a.out, a.ldo = out.data_ptr(), out.stride(0)
    a.lse = lse.data_ptr() if lse is not None else None
    a.lse_stride = lse.shape[-1] if lse is not None else 0
    if key_padding_mask is not None:
        if key_padding_mask.dtype not in (torch.uint8, torch.bool):
            raise _lib.XfError("key_padding_mask must be uint8/bool")
        a.key_padding_mask = key_padding_mask.data_ptr()
    a.kpm_start = kpm_start
    a.debug_timeline = debug_timeline.data_ptr() if debug_timeline is not None else None
    a.B, a.H, a.Sq, a.Sk, a.dp = B, H, Sq, Sk, dp
    a.scale = scale
    a.drop_p, a.drop_seed, a.drop_stream = drop_p, drop_seed, drop_stream
    d_true = 1.0 / (scale * scale)
    with _Prof("attn_fwd", 4.0 * B * H * Sq * Sk * d_true):
        check(lib().xf_attn_fwd(C.byref(a), _stream()), "xf_attn_fwd")


def attn_bwd(q, k, v, d_out, lse, delta, dq, dk, dv, *, B: int, H: int, Sq: int, Sk: int, dp: int, scale: float,
             key_padding_mask: Optional[torch.Tensor] = None, kpm_start: int = 0, drop_p: float = 0.0,
             drop_seed: int = 0, drop_stream: int = 0, debug_timeline: Optional[torch.Tensor] = None,
             workspace: Optional[torch.Tensor] = None):
    """workspace: optional uint8 scratch of attn_bwd_workspace_bytes(...) bytes -> 5-unit backward (include/xfusion.h)."""
    a = XfAttnBwd()
    a.q, a.ldq = q.data_ptr(), q.stride(0)
    a.k, a.ldk = k.data_ptr(), k.stride(0)
    a.v, a.ldv = v.data_ptr(), v.stride(0)
    a.d_out, a.lddo = d_out.data_ptr(), d_out.stride(0)
    a.lse, a.delta, a.stat_stride = lse.data_ptr(), delta.data_ptr(), lse.shape[-1]
    if delta.shape[-1] != lse.shape[-1]:
        raise _lib.XfError("lse and delta must share their last-dim stride")
    a.dq, a.lddq = dq.data_ptr(), dq.stride(0)
    a.dk, a.lddk = dk.data_ptr(), dk.stride(0)
    a.dv, a.lddv = dv.data_ptr(), dv.stride(0)
    if key_padding_mask is not None:
        a.key_padding_mask = key_padding_mask.data_ptr()
    a.kpm_start = kpm_start
    a.debug_timeline = debug_timeline.data_ptr() if debug_timeline is not None else None
    a.B, a.H, a.Sq, a.Sk, a.dp = B, H, Sq, Sk, dp
    a.scale = scale
    a.drop_p, a.drop_seed, a.drop_stream = drop_p, drop_seed, drop_stream
    if workspace is not None:
        a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    d_true = 1.0 / (scale * scale)
    with _Prof("attn_bwd", 8.0 * B * H * Sq * Sk * d_true):
        check(lib().xf_attn_bwd(C.byref(a), _stream()), "xf_attn_bwd")


def attn_bwd_workspace_bytes(B: int, H: int, Sq: int, Sk: int) -> int:
    return int(lib().xf_attn_bwd_workspace_bytes(B, H, Sq, Sk))


def rows_gather(src: torch.Tensor, dst: torch.Tensor, rows: int, D: int, in_map=(0, 0, 0), colsum: Optional[torch.Tensor] = None,
                drop_p: float = 0.0, drop_seed: int = 0, drop_stream: int = 0):
    _req(src, torch.bfloat16, "src"); _req(dst, torch.bfloat16, "dst")
    with _Prof("rows_gather", 0.0, 4.0 * rows * D):
      check(lib().xf_rows_gather(_ptr(src), C.c_int64(src.stride(-2)), _ptr(dst), C.c_int64(dst.stride(-2)), rows, D,
                               in_map[0], in_map[1], in_map[2], _ptr(colsum), C.c_float(drop_p), C.c_uint32(drop_seed),
                               C.c_uint32(drop_stream), _stream()), "xf_rows_gather")


# ---- LM head (fp32) ------------------------------------------------------------------------------
def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _req(t, torch.float32, name)
    return t if t.is_contiguous() else t.contiguous()


def lm_pool_fwd(tok: torch.Tensor, mask: Optional[torch.Tensor], kind: str):
    """tok fp32 [B, L, D]; mask uint8 [B, L] (1 = valid) or None -> (pooled [B, D], argmax int32 [B, D] | None)."""
    tok = _f32c(tok, "tok")
    B, L, D = tok.shape
    t = {"mean": 0, "max": 1}[kind]
    pooled = torch.empty(B, D, device=tok.device, dtype=torch.float32)
    argmax = torch.empty(B, D, device=tok.device, dtype=torch.int32) if t == 1 else None
    with _Prof("lm_head", 0.0, 4.0 * B * L * D):
        check(lib().xf_lm_pool_fwd(_ptr(tok), _ptr(mask), B, L, D, t, _ptr(pooled), _ptr(argmax), _stream()), "xf_lm_pool_fwd")
    return pooled, argmax


def lm_pool_bwd(dpooled: torch.Tensor, mask: Optional[torch.Tensor], argmax: Optional[torch.Tensor], L: int, kind: str):
    dpooled = _f32c(dpooled, "dpooled")
    B, D = dpooled.shape
    t = {"mean": 0, "max": 1}[kind]
    dtok = torch.empty(B, L, D, device=dpooled.device, dtype=torch.float32)
    with _Prof("lm_head", 0.0, 4.0 * B * L * D):
        check(lib().xf_lm_pool_bwd(_ptr(dpooled), _ptr(mask), _ptr(argmax), B, L, D, t, _ptr(dtok), _stream()), "xf_lm_pool_bwd")
    return dtok


def rowln_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float):
    x = _f32c(x, "x")
    R, D = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(R, device=x.device, dtype=torch.float32)
    rstd = torch.empty(R, device=x.device, dtype=torch.float32)
    with _Prof("lm_head", 0.0, 8.0 * R * D):
        check(lib().xf_rowln_fwd(_ptr(x), _ptr(_f32c(gamma, "gamma")), _ptr(_f32c(beta, "beta")), R, D, C.c_float(eps), _ptr(y),
                                 _ptr(mean), _ptr(rstd), _stream()), "xf_rowln_fwd")
    return y, mean, rstd


def rowln_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor):
    dy, x = _f32c(dy, "dy"), _f32c(x, "x")
    R, D = x.shape
    dx = torch.empty_like(x)
    dgamma = torch.zeros(D, device=x.device, dtype=torch.float32)
    dbeta = torch.zeros(D, device=x.device, dtype=torch.float32)
    with _Prof("lm_head", 0.0, 12.0 * R * D):
        check(lib().xf_rowln_bwd(_ptr(dy), _ptr(x), _ptr(_f32c(gamma, "gamma")), _ptr(mean), _ptr(rstd), R, D, _ptr(dx), _ptr(dgamma),
                                 _ptr(dbeta), _stream()), "xf_rowln_bwd")
    return dx, dgamma, dbeta


def small_linear_fwd(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], act: int):
    x, W = _f32c(x, "x"), _f32c(W, "W")
    R, D = x.shape
    Cn = W.shape[0]
    y = torch.empty(R, Cn, device=x.device, dtype=torch.float32)
    with _Prof("lm_head", 2.0 * R * Cn * D):
        check(lib().xf_small_linear_fwd(_ptr(x), _ptr(W), _ptr(None if bias is None else _f32c(bias, "bias")), R, Cn, D, act, _ptr(y),
                                        _stream()), "xf_small_linear_fwd")
    return y


def small_linear_bwd(dy: torch.Tensor, x: torch.Tensor, W: torch.Tensor, act: int, need_dx=True, need_dw=True, has_bias=True):
    dy, x, W = _f32c(dy, "dy"), _f32c(x, "x"), _f32c(W, "W")
    R, D = x.shape
    Cn = W.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    dW = torch.zeros_like(W) if need_dw else None
    db = torch.zeros(Cn, device=x.device, dtype=torch.float32) if (need_dw and has_bias) else None
    with _Prof("lm_head", 4.0 * R * Cn * D):
        check(lib().xf_small_linear_bwd(_ptr(dy), _ptr(x), _ptr(W), R, Cn, D, act, _ptr(dx), _ptr(dW), _ptr(db), _stream()),
              "xf_small_linear_bwd")
    return dx, dW, db


# ---- test aids: the dropout keep masks the kernels recompute on the fly (include/xfusion.h) ------------------------
def debug_dropout_mask(p: float, seed: int, stream: int, row0: int, rows: int, cols: int, device="cuda") -> torch.Tensor:
    out = torch.empty(rows, cols, device=device, dtype=torch.uint8)
    check(lib().xf_debug_dropout_mask(C.c_float(p), C.c_uint32(seed), C.c_uint32(stream), C.c_int64(row0), rows, cols, _ptr(out),
                                      _stream()), "xf_debug_dropout_mask")
    return out


def debug_attn_dropout_mask(p: float, seed: int, stream: int, BH: int, Sq: int, Sk: int, device="cuda") -> torch.Tensor:
    out = torch.empty(BH, Sq, Sk, device=device, dtype=torch.uint8)
    check(lib().xf_debug_attn_dropout_mask(C.c_float(p), C.c_uint32(seed), C.c_uint32(stream), BH, Sq, Sk, _ptr(out), _stream()),
          "xf_debug_attn_dropout_mask")
    return out


# ---- fp32-tolerance mode helpers (csrc/fp32_mode.cu) -------------------------------------------------------------------
def split3(src: torch.Tensor, pattern: int, act: int = 0, bias: Optional[torch.Tensor] = None, bias_cols: int = 0) -> torch.Tensor:
    """fp32 [rows, cols] (row stride src.stride(0)) -> bf16 [rows, 6*cols + bias_cols]: the 3-term split concatenated along
    K (pattern 0: A side, 1: B side); xf_gemm on an A-side and a B-side operand gives the fp32-accurate product.
    bias_cols = 8 appends the bias-carrying columns (A side: ones; B side: the split of bias[row]); act = 1: GELU first."""
    _req(src, torch.float32, "src")
    rows, cols = src.shape
    dst = torch.empty(rows, 6 * cols + bias_cols, device=src.device, dtype=torch.bfloat16)
    with _Prof("fp32_mode", 0.0, 16.0 * rows * cols):
        check(lib().xf_split3(_ptr(src), C.c_int64(src.stride(0)), rows, cols, _ptr(dst), pattern, act, bias_cols, _ptr(bias),
                              _stream()), "xf_split3")
    return dst


def softmax_rows_f32(s: torch.Tensor, Sk: int, kpm: Optional[torch.Tensor], scale: float):
    """in place on fp32 [B, H, Sq, Sp]"""
    _req(s, torch.float32, "s")
    B, H, Sq, Sp = s.shape
    with _Prof("fp32_mode", 0.0, 8.0 * s.numel()):
        check(lib().xf_softmax_rows_f32(_ptr(s), B, H, Sq, Sk, Sp, _ptr(kpm), C.c_float(scale), _stream()), "xf_softmax_rows_f32")


if ABLATE:   # dev aid only: swallow the skip signal of _Prof for the void-returning hot-path ops
    import functools as _ft

    def _ablatable(fn):
        @_ft.wraps(fn)
        def wrapper(*a, **k):
            try:
                return fn(*a, **k)
            except _Skip:
                return None
        return wrapper

    for _name in ("gemm", "patchify", "fold", "lang_rows_fwd", "lang_rows_bwd", "layernorm_fwd", "layernorm_bwd", "colsum", "cast_pad",
                  "cast_pad_multi", "unpad_add", "attn_delta", "attn_fwd", "attn_bwd", "rows_gather"):
        globals()[_name] = _ablatable(globals()[_name])
