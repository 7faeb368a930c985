"""GPU: every C-ABI op against a torch fp32 restatement of the same op on the same bf16 inputs
(kernel-level parity; the path-level parity against the oracle / golden vectors is in
test_gpu_fusion_parity.py).  Tolerances: bf16 output rounding is 2^-9 relative per element
(rel-Frobenius ~1.7e-3); fp32 outputs 1e-5."""
import math

import pytest
import torch

from transfusion_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,split", [
    (128, 128, 64, False, False, 1), (1000, 712, 712, False, False, 1), (4096, 2688, 896, False, False, 1),
    (512, 896, 1792, False, True, 1), (130, 96, 40, False, True, 1), (256, 128, 128, True, False, 1),
    (896, 1792, 4096, True, True, 4), (712, 1424, 1000, True, True, 3), (32, 32, 24, False, False, 1),
])
def test_gemm_operand_layouts(M, N, K, a_mn, b_mn, split):
    torch.manual_seed(0)
    A = torch.randn(M, K, device=DEV).bfloat16()
    B = torch.randn(N, K, device=DEV).bfloat16()
    ref = A.float() @ B.float().t()
    a_st = A.t().contiguous() if a_mn else A
    b_st = B.t().contiguous() if b_mn else B
    if split > 1:
        out = torch.zeros(M, N, device=DEV)
        ops.gemm(a_st, b_st, out, M=M, N=N, K=K, a_mn_major=a_mn, b_mn_major=b_mn, split_k=split, accumulate=True)
        assert rel(out, ref) < 1e-5
    else:
        out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        ops.gemm(a_st, b_st, out, M=M, N=N, K=K, a_mn_major=a_mn, b_mn_major=b_mn)
        assert rel(out.float(), ref) < 3e-3


def test_gemm_epilogue_bias_gelu_residual_remap():
    torch.manual_seed(1)
    Bt, n, S, N, K = 3, 50, 60, 96, 72
    A = torch.randn(Bt * n, K, device=DEV).bfloat16()
    W = torch.randn(N, K, device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV)
    pos = torch.randn(64, N, device=DEV)
    res = torch.randn(Bt * S, N, device=DEV).bfloat16()
    out = torch.zeros(Bt * S, N, device=DEV, dtype=torch.bfloat16)
    pre = torch.zeros(Bt * S, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, out, M=Bt * n, N=N, K=K, bias=bias, pos_table=pos, rows_in=n, rows_out=S, row_off=2, act=1,
             preact_out=pre, residual=res)
    u = (A.float() @ W.float().t() + bias).view(Bt, n, N) + pos[:n]
    ref = torch.nn.functional.gelu(u) + res.float().view(Bt, S, N)[:, 2:2 + n]
    got = out.float().view(Bt, S, N)[:, 2:2 + n]
    assert rel(got, ref) < 4e-3
    assert rel(pre.float().view(Bt, S, N)[:, 2:2 + n], u) < 3e-3
    assert float(out.float().view(Bt, S, N)[:, :2].abs().max()) == 0.0  # rows outside the remap untouched


def test_gemm_dact_matches_gelu_backward():
    torch.manual_seed(2)
    M, N, K = 300, 128, 64
    A = torch.randn(M, K, device=DEV).bfloat16()
    W = torch.randn(N, K, device=DEV).bfloat16()
    u = torch.randn(M, N, device=DEV).bfloat16()
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, out, M=M, N=N, K=K, dact_in=u)
    uu = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uu).backward(A.float() @ W.float().t())
    assert rel(out.float(), uu.grad) < 4e-3


@pytest.mark.parametrize("p", [0.1, 0.15])
def test_gemm_dropout_rate_and_mask_reproducible(p):
    torch.manual_seed(3)
    M, N, K = 2048, 896, 64
    # strictly positive operands: every kept output is non-zero, so (out != 0) is exactly the keep mask
    A = (torch.rand(M, K, device=DEV) + 0.25).bfloat16()
    W = (torch.rand(N, K, device=DEV) + 0.25).bfloat16()
    ref = A.float() @ W.float().t()
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, out, M=M, N=N, K=K, drop_p=p, drop_seed=123, drop_stream=7)
    keep = out != 0
    rate = 1.0 - float(keep.float().mean())
    assert abs(rate - p) < 4 * math.sqrt(p * (1 - p) / (M * N)) + 1e-3, rate
    assert rel(out.float()[keep], ref[keep] / (1 - p)) < 4e-3
    # the backward re-creates the mask from (seed, stream, output index): different operands, same mask
    out2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    A2 = (torch.rand(M, K, device=DEV) + 0.5).bfloat16()
    W2 = (torch.rand(N, K, device=DEV) + 0.5).bfloat16()
    ops.gemm(A2, W2, out2, M=M, N=N, K=K, drop_p=p, drop_seed=123, drop_stream=7, drop_first=True)
    assert torch.equal(out2 != 0, keep)
    out3 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A2, W2, out3, M=M, N=N, K=K, drop_p=p, drop_seed=124, drop_stream=7)
    assert not torch.equal(out3 != 0, keep)
    # column-wise and row-wise keep rates are uniform (no visible structure)
    assert float((keep.float().mean(0) - (1 - p)).abs().max()) < 0.05
    assert float((keep.float().mean(1) - (1 - p)).abs().max()) < 0.07


@pytest.mark.parametrize("M,N,K", [(300, 712, 712), (517, 1424, 712), (1000, 896, 896)])
def test_gemm_specialised_epilogues(M, N, K):
    """The per-kind epilogues (LINEAR / GELU / DGELU; TMA-box I/O) incl. N % 32 != 0 (Ego4Dv1: D = 712), rows that end
    inside a 32-row box, and the same dropout mask in the specialised and the generic (row-remapped) kernels."""
    torch.manual_seed(30)
    A = (0.1 * torch.randn(M, K, device=DEV)).bfloat16()
    W = torch.randn(N, K, device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV).bfloat16()
    z = A.float() @ W.float().t()
    # LINEAR: bias + residual
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, out, M=M, N=N, K=K, bias=bias, residual=res)
    assert rel(out.float(), z + bias + res.float()) < 4e-3
    # GELU: bias, pre-activation store, GELU(erf)
    pre = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, out, M=M, N=N, K=K, bias=bias, act=1, preact_out=pre)
    assert rel(pre.float(), z + bias) < 3e-3
    assert rel(out.float(), torch.nn.functional.gelu(z + bias)) < 4e-3
    # DGELU: x GELU'(saved pre-activation)
    ops.gemm(A, W, out, M=M, N=N, K=K, dact_in=pre)
    uu = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(uu).backward(z)
    assert rel(out.float(), uu.grad) < 4e-3
    # dropout: specialised kernel vs generic kernel (forced by an identity row remap) produce the same mask
    Ap = (torch.rand(M, K, device=DEV) + 0.25).bfloat16()
    Wp = (torch.rand(N, K, device=DEV) + 0.25).bfloat16()
    o1 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    o2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(Ap, Wp, o1, M=M, N=N, K=K, drop_p=0.15, drop_seed=5, drop_stream=2)
    ops.gemm(Ap, Wp, o2, M=M, N=N, K=K, drop_p=0.15, drop_seed=5, drop_stream=2, rows_in=M, rows_out=M, row_off=0)
    assert torch.equal(o1 != 0, o2 != 0)
    assert rel(o1.float(), o2.float()) < 1e-6


# ------------------------------------------------------------------ attention
def attn_ref(q, k, v, H, d, kpm, drop_mask=None, p=0.0):
    B, Sq, _ = q.shape
    Sk = k.shape[1]
    qh = q.reshape(B, Sq, H, d).transpose(1, 2)
    kh = k.reshape(B, Sk, H, d).transpose(1, 2)
    vh = v.reshape(B, Sk, H, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d)
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    a = torch.softmax(s, dim=-1)
    if drop_mask is not None:
        a = a * drop_mask / (1 - p)
    return (a @ vh).transpose(1, 2).reshape(B, Sq, H * d), a


def _pack(src, dp):
    B, S, W, H, d = src.shape
    buf = torch.zeros(B * S, W * H * dp, device=DEV, dtype=torch.bfloat16)
    buf.view(B, S, W, H, dp)[..., :d] = src
    return buf


@pytest.mark.parametrize("B,H,S,d,mask", [(2, 4, 300, 64, True), (2, 4, 832, 224, True), (2, 4, 333, 178, True),
                                          (1, 2, 64, 32, False), (2, 4, 70, 16, True),
                                          # the longest sequence the reference admits: MAX_NUM_PATCHES = 8192 visual tokens
                                          # (cross_f_box_wrapper.py:21) + 64 language tokens
                                          (1, 1, 8256, 32, True)])
def test_attention_forward_backward(B, H, S, d, mask):
    torch.manual_seed(4)
    dp = (d + 31) // 32 * 32
    D = H * d
    src = torch.randn(B, S, 3, H, d, device=DEV).bfloat16()
    qkv = _pack(src, dp)
    Dp = H * dp
    kpm = None
    if mask:
        kpm = torch.zeros(B, S, dtype=torch.uint8, device=DEV)
        for b in range(B):
            kpm[b, S - 3 - 5 * b:] = 1
    Sp = (S + 127) // 128 * 128
    out = torch.zeros(B * S, Dp, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, Sp, device=DEV)
    ops.attn_fwd(qkv[:, :Dp], qkv[:, Dp:2 * Dp], qkv[:, 2 * Dp:], out, lse, B=B, H=H, Sq=S, Sk=S, dp=dp,
                 scale=1 / math.sqrt(d), key_padding_mask=kpm, kpm_start=0)
    qf, kf, vf = (src[:, :, i].float().reshape(B, S, D).requires_grad_(True) for i in range(3))
    ref, _ = attn_ref(qf, kf, vf, H, d, kpm)
    assert rel(out.view(B, S, H, dp)[..., :d].reshape(B, S, D).float(), ref) < 5e-3
    dsrc = torch.randn(B, S, H, d, device=DEV).bfloat16()
    dout = torch.zeros(B * S, Dp, device=DEV, dtype=torch.bfloat16)
    dout.view(B, S, H, dp)[..., :d] = dsrc
    delta = torch.zeros(B, H, Sp, device=DEV)
    ops.attn_delta(out, dout, delta, B, S, H, dp)
    ref.backward(dsrc.float().reshape(B, S, D))
    # both backward schedules: three on-chip passes (no workspace) and the 5-unit path (key-stationary pass + two batched
    # GEMMs over a bf16 [B, H, Sk, Sq] scratch)
    for ws in (None, torch.empty(ops.attn_bwd_workspace_bytes(B, H, S, S), device=DEV, dtype=torch.uint8)):
        dqkv = torch.full((B * S, 3 * Dp), 7.0, device=DEV, dtype=torch.bfloat16)
        ops.attn_bwd(qkv[:, :Dp], qkv[:, Dp:2 * Dp], qkv[:, 2 * Dp:], dout, lse, delta, dqkv[:, :Dp], dqkv[:, Dp:2 * Dp],
                     dqkv[:, 2 * Dp:], B=B, H=H, Sq=S, Sk=S, dp=dp, scale=1 / math.sqrt(d), key_padding_mask=kpm, workspace=ws)
        got = dqkv.view(B, S, 3, H, dp)
        for i, g in enumerate((qf.grad, kf.grad, vf.grad)):
            assert rel(got[:, :, i, :, :d].reshape(B, S, D).float(), g) < 8e-3, ("workspace" if ws is not None else "3-pass", i)
        if dp != d:
            assert float(got[..., d:].float().abs().max()) == 0.0


def test_cross_attention_lq_ne_lk():
    """General Lq != Lk + key padding (the QKVEncoder-style op, cross_qkv_layers.py:70-77)."""
    torch.manual_seed(5)
    B, H, Sq, Sk, d = 2, 2, 200, 333, 32
    qs = torch.randn(B, Sq, 1, H, d, device=DEV).bfloat16()
    ks = torch.randn(B, Sk, 1, H, d, device=DEV).bfloat16()
    vs = torch.randn(B, Sk, 1, H, d, device=DEV).bfloat16()
    kpm = torch.zeros(B, Sk, dtype=torch.uint8, device=DEV)
    kpm[0, 300:] = 1
    kpm[1, 17:40] = 1  # arbitrary (non-suffix) mask
    out = torch.zeros(B * Sq, H * d, device=DEV, dtype=torch.bfloat16)
    ops.attn_fwd(_pack(qs, d), _pack(ks, d), _pack(vs, d), out, None, B=B, H=H, Sq=Sq, Sk=Sk, dp=d, scale=1 / math.sqrt(d),
                 key_padding_mask=kpm)
    ref, _ = attn_ref(qs.float().reshape(B, Sq, -1), ks.float().reshape(B, Sk, -1), vs.float().reshape(B, Sk, -1), H, d, kpm)
    assert rel(out.view(B, Sq, H * d).float(), ref) < 5e-3


def test_attention_dropout_mask_consistent_between_forward_and_backward():
    """V = identity exposes the dropped probabilities P_d as the forward output; the backward must use
    the same mask (checked against a torch reference that applies the extracted mask)."""
    torch.manual_seed(6)
    B, H, S, d, p = 2, 2, 64, 64, 0.15
    src = torch.randn(B, S, 3, H, d, device=DEV).bfloat16()
    src[:, :, 2] = torch.eye(S, d, device=DEV).bfloat16()[None, :, None, :]
    qkv = _pack(src, d)
    Dp = H * d
    out = torch.zeros(B * S, Dp, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, 128, device=DEV)
    kw = dict(B=B, H=H, Sq=S, Sk=S, dp=d, scale=1 / math.sqrt(d), drop_p=p, drop_seed=77, drop_stream=5)
    ops.attn_fwd(qkv[:, :Dp], qkv[:, Dp:2 * Dp], qkv[:, 2 * Dp:], out, lse, **kw)
    pd = out.view(B, S, H, d).permute(0, 2, 1, 3).float()  # [B,H,q,k] dropped probabilities
    qf, kf, vf = (src[:, :, i].float().reshape(B, S, H * d).requires_grad_(True) for i in range(3))
    _, a = attn_ref(qf, kf, vf, H, d, None)
    mask = (pd != 0).float()
    rate = 1 - float(mask.mean())
    assert abs(rate - p) < 0.02, rate
    assert rel(pd, a.detach() * mask / (1 - p)) < 6e-3
    ref, _ = attn_ref(qf, kf, vf, H, d, None, drop_mask=mask, p=p)
    dsrc = torch.randn(B, S, H, d, device=DEV).bfloat16()
    ref.backward(dsrc.float().reshape(B, S, H * d))
    dout = dsrc.reshape(B * S, Dp).contiguous()
    delta = torch.zeros(B, H, 128, device=DEV)
    ops.attn_delta(out, dout, delta, B, S, H, d)
    for ws in (None, torch.empty(ops.attn_bwd_workspace_bytes(B, H, S, S), device=DEV, dtype=torch.uint8)):
        dqkv = torch.zeros(B * S, 3 * Dp, device=DEV, dtype=torch.bfloat16)
        ops.attn_bwd(qkv[:, :Dp], qkv[:, Dp:2 * Dp], qkv[:, 2 * Dp:], dout, lse, delta, dqkv[:, :Dp], dqkv[:, Dp:2 * Dp],
                     dqkv[:, 2 * Dp:], key_padding_mask=None, workspace=ws, **kw)
        got = dqkv.view(B, S, 3, H, d)
        for i, g in enumerate((qf.grad, kf.grad, vf.grad)):
            assert rel(got[:, :, i].reshape(B, S, H * d).float(), g) < 1e-2, ("workspace" if ws is not None else "3-pass", i)


# ------------------------------------------------------------------ LayerNorm / layout / misc
@pytest.mark.parametrize("rows,D", [(999, 896), (300, 712), (129, 64), (50, 1536)])
def test_layernorm_forward_backward(rows, D):
    torch.manual_seed(7)
    x = torch.randn(rows, D, device=DEV).bfloat16()
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    ops.layernorm_fwd(x, y, g, b, mean, rstd, rows, D)
    xr, gr, br = x.float().requires_grad_(True), g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)
    assert rel(y.float(), yr) < 3e-3
    dy = torch.randn(rows, D, device=DEV).bfloat16()
    yr.backward(dy.float())
    dx = torch.empty_like(x)
    dg, db, dbias = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dx, dg, db, rows, D, dbias=dbias)
    assert rel(dx.float(), xr.grad) < 3e-3
    assert rel(dg, gr.grad) < 1e-4 and rel(db, br.grad) < 1e-4
    assert rel(dbias, xr.grad.sum(0)) < 5e-3


def test_layernorm_row_remap_and_dropout_pairing():
    torch.manual_seed(8)
    B, S, n, D, p = 3, 40, 33, 128, 0.1
    x = torch.randn(B * S, D, device=DEV).bfloat16()
    g, b = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    y = torch.empty(B * n, D, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(B * n, device=DEV), torch.empty(B * n, device=DEV)
    ops.layernorm_fwd(x, y, g, b, mean, rstd, B * n, D, in_map=(n, S, 0), drop_p=p, drop_seed=5, drop_stream=9)
    ref = torch.nn.functional.layer_norm(x.float().view(B, S, D)[:, :n], (D,), g, b, 1e-5).reshape(B * n, D)
    keep = y != 0
    assert abs(1 - float(keep.float().mean()) - p) < 0.02
    assert rel(y.float()[keep], ref[keep] / (1 - p)) < 4e-3
    # backward with the same (seed, stream): masked positions contribute nothing
    dy = torch.ones(B * n, D, device=DEV, dtype=torch.bfloat16)
    dx = torch.zeros(B * S, D, device=DEV, dtype=torch.bfloat16)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.layernorm_bwd(dy, x, g, mean, rstd, dx, dg, db, B * n, D, in_map=(n, S, 0), dy_drop=(p, 5, 9))
    assert rel(db, keep.float().sum(0) / (1 - p)) < 5e-3
    assert float(dx.view(B, S, D)[:, n:].float().abs().max()) == 0.0


@pytest.mark.parametrize("B,C,H,W,p,dt", [(2, 256, 16, 24, 4, torch.float32), (2, 512, 8, 12, 4, torch.bfloat16),
                                          (2, 64, 12, 20, 2, torch.float32), (3, 2048, 24, 32, 1, torch.float32),
                                          (2, 8, 16, 24, 4, torch.float32), (1, 16, 16, 272, 4, torch.float32),
                                          (1, 6, 16, 64, 8, torch.float32),
                                          # bf16 maps on the 128-bit path (W % 8 == 0), incl. partial token tiles
                                          (2, 256, 16, 24, 4, torch.bfloat16), (1, 16, 16, 272, 4, torch.bfloat16),
                                          (1, 6, 16, 64, 8, torch.bfloat16), (2, 8, 16, 40, 4, torch.bfloat16)])
def test_patchify_fold_bit_exact(B, C, H, W, p, dt):
    torch.manual_seed(9)
    f = torch.randn(B, C, H, W, device=DEV).to(dt)
    gh, gw = H // p, W // p
    tok = torch.empty(B * gh * gw, C * p * p, device=DEV, dtype=torch.bfloat16)
    ops.patchify(f, p, tok)
    ref = f.float().reshape(B, C, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, C * p * p).bfloat16()
    assert torch.equal(tok, ref)
    out = torch.zeros(B, C, H, W, device=DEV, dtype=dt)
    ops.fold(tok, out, p)
    assert torch.equal(out.float(), f.bfloat16().float())  # fold(patchify(x)) = bf16(x): a pure permutation


@pytest.mark.parametrize("B,C,H,W,p,dt", [(2, 256, 16, 24, 4, torch.float32), (2, 70, 12, 20, 2, torch.bfloat16),
                                          (3, 2048, 6, 8, 1, torch.float32), (1, 48, 16, 272, 4, torch.bfloat16)])
def test_patchify_fold_channels_last_bit_exact(B, C, H, W, p, dt):
    """channels_last (NHWC memory) maps are read and written in place of an NCHW conversion pass."""
    torch.manual_seed(19)
    f = torch.randn(B, C, H, W, device=DEV).to(dt).contiguous(memory_format=torch.channels_last)
    assert not f.is_contiguous() or C == 1
    gh, gw = H // p, W // p
    tok = torch.empty(B * gh * gw, C * p * p, device=DEV, dtype=torch.bfloat16)
    ops.patchify(f, p, tok)
    ref = f.float().reshape(B, C, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, C * p * p).bfloat16()
    assert torch.equal(tok, ref)
    out = torch.zeros(B, C, H, W, device=DEV, dtype=dt).contiguous(memory_format=torch.channels_last)
    ops.fold(tok, out, p)
    assert out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out.float(), f.bfloat16().float())


def test_small_kernels():
    torch.manual_seed(10)
    x = torch.randn(1000, 2688, device=DEV).bfloat16()
    out = torch.zeros(2688, device=DEV)
    ops.colsum(x, out, 1000, 2688)
    assert rel(out, x.float().sum(0)) < 1e-5
    B, L, D, n = 3, 7, 64, 10
    lang, kind = torch.randn(B, L, D, device=DEV), torch.randn(D, device=DEV)
    z = torch.zeros(B, n + L, D, device=DEV, dtype=torch.bfloat16)
    ops.lang_rows_fwd(lang, kind, z, n)
    assert torch.equal(z[:, n:], (lang + kind).bfloat16()) and float(z[:, :n].float().abs().max()) == 0
    dz = torch.randn(B, n + L, D, device=DEV).bfloat16()
    dlang, dkind = torch.zeros(B, L, D, device=DEV), torch.zeros(D, device=DEV)
    ops.lang_rows_bwd(dz, dlang, dkind, B, L, n)
    assert torch.equal(dlang, dz[:, n:].float()) and rel(dkind, dz[:, n:].float().sum((0, 1))) < 1e-5
    w = torch.randn(30, 24, device=DEV)
    wp = torch.zeros(48, 24, device=DEV, dtype=torch.bfloat16)
    ops.cast_pad(w, wp, 30, 24, rin=10, rout=16)
    assert torch.equal(wp.view(3, 16, 24)[:, :10].reshape(30, 24), w.bfloat16())
    assert float(wp.view(3, 16, 24)[:, 10:].float().abs().max()) == 0
    # multi-tensor cast: plain, row-padded, column-padded and an odd-shaped job (single-tensor fallback) in one call
    ws = [torch.randn(64, 48, device=DEV), torch.randn(30, 24, device=DEV), torch.randn(16, 40, device=DEV), torch.randn(5, 7, device=DEV)]
    ds = [torch.zeros(64, 48, device=DEV, dtype=torch.bfloat16), torch.zeros(48, 24, device=DEV, dtype=torch.bfloat16),
          torch.zeros(16, 64, device=DEV, dtype=torch.bfloat16), torch.zeros(5, 7, device=DEV, dtype=torch.bfloat16)]
    ops.cast_pad_multi([(ws[0], ds[0], 64, 48, 0, 0, 0, 0), (ws[1], ds[1], 30, 24, 10, 16, 0, 0),
                        (ws[2], ds[2], 16, 40, 0, 0, 10, 16), (ws[3], ds[3], 5, 7, 0, 0, 0, 0)])
    assert torch.equal(ds[0], ws[0].bfloat16()) and torch.equal(ds[3], ws[3].bfloat16())
    assert torch.equal(ds[1].view(3, 16, 24)[:, :10].reshape(30, 24), ws[1].bfloat16())
    assert torch.equal(ds[2].view(16, 4, 16)[:, :, :10].reshape(16, 40), ws[2].bfloat16())
    assert float(ds[2].view(16, 4, 16)[:, :, 10:].float().abs().max()) == 0
    gsrc, gdst = torch.randn(30, 48, device=DEV), torch.ones(30, 30, device=DEV)
    ops.unpad_add(gsrc, gdst, 30, 30, cin=10, cout=16)
    assert torch.allclose(gdst, 1 + gsrc.view(30, 3, 16)[:, :, :10].reshape(30, 30))
    src = torch.randn(3 * 20, 64, device=DEV).bfloat16()
    dst = torch.zeros(3 * 12, 64, device=DEV, dtype=torch.bfloat16)
    cs = torch.zeros(64, device=DEV)
    ops.rows_gather(src, dst, 36, 64, in_map=(12, 20, 0), colsum=cs)
    assert torch.equal(dst.view(3, 12, 64), src.view(3, 20, 64)[:, :12]) and rel(cs, dst.float().sum(0)) < 1e-5


@pytest.mark.parametrize("kind,ln,repr_size", [("mean", True, None), ("max", True, 48), ("mean", False, 40)])
def test_lm_head_matches_reference_modules(kind, ln, repr_size):
    """PoolPredictor (lm_layers.py:30-81) through the fp32 LM-head kernels vs the same math in stock torch:
    logits and every gradient (tokens, LayerNorm, optional GELU+Linear, noun / verb Linears)."""
    from transfusion_b200.cross_fusion.lm_layers import PoolPredictor
    torch.manual_seed(21)
    B, L, D, nn_, nv = 5, 11, 96, 13, 7
    args = {"type": kind, "ln": ln}
    if repr_size:
        args["repr_size"] = repr_size
    head = PoolPredictor(args, D, nn_, nv).to(DEV)
    tok = torch.randn(B, L, D, device=DEV, requires_grad=True)
    mask = torch.ones(B, L, dtype=torch.bool, device=DEV)
    mask[1, 6:] = False
    mask[3, 1:] = False
    out = head(tok, mask)
    (out["noun_logits"].pow(2).sum() + out["verb_logits"].sum()).backward()
    got = {k: p.grad.clone() for k, p in head.named_parameters()}
    gtok = tok.grad.clone()
    # stock torch restatement of lm_layers.py:59-81 on the same parameters
    head.zero_grad(set_to_none=True)
    tok2 = tok.detach().clone().requires_grad_(True)
    x = tok2 * mask.unsqueeze(2)
    f = x.max(dim=1)[0] if kind == "max" else x.mean(dim=1)
    if head.ln:
        f = head.ln(f)
    if head.repr_mlp:
        f = head.repr_mlp(f)
    noun, verb = head.mlp_noun(f), head.mlp_verb(f)
    (noun.pow(2).sum() + verb.sum()).backward()
    assert rel(out["noun_logits"], noun) < 1e-5 and rel(out["verb_logits"], verb) < 1e-5
    assert rel(gtok, tok2.grad) < 1e-5
    for k, p in head.named_parameters():
        assert rel(got[k], p.grad) < 1e-5, k


# ------------------------------------------------------------------ batched GEMM (per (sample, head) problems)
@pytest.mark.parametrize("Bt,H,Sq,Sk,dp", [(2, 4, 300, 300, 224), (3, 2, 200, 328, 192), (1, 4, 832, 832, 32)])
def test_gemm_batched_head_addressing(Bt, H, Sq, Sk, dp):
    """The two batched products of the attention backward: dQ[b, q, h*dp + e] = sum_k E[b, h, k, q] K[b, k, h*dp + e]
    (A stored [K][M], B stored [K][N], token-major output with heads in columns) and
    dV[b, k, h*dp + e] = sum_q P[b, h, k, q] dO[b, q, h*dp + e] (A stored [M][K]).  Rows / columns outside the per-entry
    extents (the padded pitch of E, the next head's columns) must not leak into the result."""
    torch.manual_seed(40)
    Sqp = (Sq + 63) // 64 * 64
    E = torch.randn(Bt, H, Sk, Sqp, device=DEV).bfloat16()     # garbage in the padded columns on purpose
    kv = torch.randn(Bt * Sk, 3 * H * dp, device=DEV).bfloat16()
    K = kv[:, H * dp:2 * H * dp]
    dq = torch.zeros(Bt * Sq, 3 * H * dp, device=DEV, dtype=torch.bfloat16)
    ld = 3 * H * dp
    ops.gemm(E, K, dq, M=Sq, N=dp, K=Sk, a_mn_major=True, b_mn_major=True, a_ld=Sqp, b_ld=ld, ldc=ld,
             batch=(Bt, H, (H * Sk * Sqp, Sk * Sqp), (Sk * ld, dp), (Sq * ld, dp)))
    ref = torch.einsum("bhkq,bkhe->bqhe", E[..., :Sq].float(), K.float().reshape(Bt, Sk, H, dp)).reshape(Bt * Sq, H * dp)
    assert rel(dq[:, :H * dp].float(), ref) < 3e-3
    assert float(dq[:, H * dp:].abs().max()) == 0.0   # nothing written outside the addressed head columns
    # dV-style: A K-major
    do = torch.randn(Bt * Sq, H * dp, device=DEV).bfloat16()
    dv = torch.zeros(Bt * Sk, H * dp, device=DEV, dtype=torch.bfloat16)
    ops.gemm(E, do, dv, M=Sk, N=dp, K=Sq, a_mn_major=False, b_mn_major=True, a_ld=Sqp, b_ld=H * dp, ldc=H * dp,
             batch=(Bt, H, (H * Sk * Sqp, Sk * Sqp), (Sq * H * dp, dp), (Sk * H * dp, dp)))
    ref = torch.einsum("bhkq,bqhe->bkhe", E[..., :Sq].float(), do.float().reshape(Bt, Sq, H, dp)).reshape(Bt * Sk, H * dp)
    assert rel(dv.float(), ref) < 3e-3
