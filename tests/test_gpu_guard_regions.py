"""GPU: out-of-bounds WRITE check for every kernel family (compute-sanitizer is closed on this pool -- see
profiles/sanitizer_r2_unavailable.txt -- so the memcheck role is played here): every output, statistic and gradient
buffer is carved out of a larger allocation whose surroundings are filled with a sentinel; shapes are chosen so that rows
and columns end inside tiles / TMA boxes; after the launches the sentinels must be intact."""
import math

import pytest
import torch

from transfusion_b200 import ops, optim

pytestmark = pytest.mark.gpu
DEV = "cuda"
PAD = 4096  # elements of sentinel on each side


class Guard:
    def __init__(self):
        self.items = []

    def new(self, *shape, dtype=torch.bfloat16, fill=None):
        n = 1
        for s in shape:
            n *= s
        # keep the payload 256-byte aligned like the caching allocator would
        buf = torch.empty(PAD + n + PAD, device=DEV, dtype=dtype)
        sentinel = 12345.0 if dtype != torch.uint8 else 77
        buf.fill_(sentinel)
        view = buf[PAD:PAD + n].view(*shape)
        if fill is not None:
            view.fill_(fill)
        self.items.append((buf, n, sentinel))
        return view

    def check(self):
        torch.cuda.synchronize()
        for i, (buf, n, sentinel) in enumerate(self.items):
            assert bool((buf[:PAD] == sentinel).all()) and bool((buf[PAD + n:] == sentinel).all()), f"guard {i} of {len(self.items)} damaged"


def test_gemm_epilogues_do_not_write_outside_their_outputs():
    g = Guard()
    torch.manual_seed(0)
    M, N, K = 517, 712, 712      # rows end inside a 32-row box, N % 32 != 0
    A = torch.randn(M, K, device=DEV).bfloat16()
    W = torch.randn(N, K, device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV).bfloat16()
    out = g.new(M, N); ops.gemm(A, W, out, M=M, N=N, K=K, bias=bias, residual=res, drop_p=0.1, drop_seed=1, drop_stream=2)
    out2, pre = g.new(M, N), g.new(M, N)
    ops.gemm(A, W, out2, M=M, N=N, K=K, bias=bias, act=1, preact_out=pre, drop_p=0.1, drop_seed=1, drop_stream=3)
    out3 = g.new(M, N); ops.gemm(A, W, out3, M=M, N=N, K=K, dact_in=pre)
    out4 = g.new(M, N); ops.gemm(A, W, out4, M=M, N=N, K=K, bias=bias, act=2)
    # row-remapped GENERIC epilogue (patch-embed)
    Bt, n, S = 3, 50, 60
    A2 = torch.randn(Bt * n, K, device=DEV).bfloat16()
    pos = torch.randn(64, N, device=DEV)
    z = g.new(Bt * S, N, fill=0.0)
    ops.gemm(A2, W, z, M=Bt * n, N=N, K=K, bias=bias, pos_table=pos, rows_in=n, rows_out=S, drop_p=0.1, drop_seed=1, drop_stream=4)
    # wgrad (split-K fp32 reduction) and fp32 store
    gw = g.new(N, K, dtype=torch.float32, fill=0.0)
    ops.gemm(out, A, gw, M=N, N=K, K=M, a_mn_major=True, b_mn_major=True, accumulate=True, split_k=3)
    f32 = g.new(M, N, dtype=torch.float32)
    ops.gemm(A, W, f32, M=M, N=N, K=K, bias=bias)
    g.check()


def test_attention_kernels_do_not_write_outside_their_outputs():
    g = Guard()
    torch.manual_seed(1)
    B, H, S, d = 2, 4, 333, 178
    dp, Sp = 192, 384
    Dp = H * dp
    qkv = torch.zeros(B * S, 3 * Dp, device=DEV, dtype=torch.bfloat16)
    qkv.view(B * S, 3 * H, dp)[:, :, :d] = torch.randn(B * S, 3 * H, d, device=DEV).bfloat16()
    kpm = torch.zeros(B, S, dtype=torch.uint8, device=DEV)
    kpm[1, S - 9:] = 1
    out = g.new(B * S, Dp)
    lse = g.new(B, H, Sp, dtype=torch.float32)
    kw = dict(B=B, H=H, Sq=S, Sk=S, dp=dp, scale=1 / math.sqrt(d), key_padding_mask=kpm, drop_p=0.15, drop_seed=3, drop_stream=4)
    ops.attn_fwd(qkv[:, :Dp], qkv[:, Dp:2 * Dp], qkv[:, 2 * Dp:], out, lse, kpm_start=0, **kw)
    dout = torch.randn(B * S, Dp, device=DEV).bfloat16()
    delta = g.new(B, H, Sp, dtype=torch.float32)
    ops.attn_delta(out, dout, delta, B, S, H, dp)
    for use_ws in (False, True):
        dqkv = g.new(B * S, 3 * Dp)
        ws = None
        if use_ws:
            ws = g.new(ops.attn_bwd_workspace_bytes(B, H, S, S), dtype=torch.uint8)
        ops.attn_bwd(qkv[:, :Dp], qkv[:, Dp:2 * Dp], qkv[:, 2 * Dp:], dout, lse, delta, dqkv[:, :Dp], dqkv[:, Dp:2 * Dp],
                     dqkv[:, 2 * Dp:], workspace=ws, **kw)
    g.check()


def test_elementwise_kernels_do_not_write_outside_their_outputs():
    g = Guard()
    torch.manual_seed(2)
    rows, D = 333, 712
    x = torch.randn(rows, D, device=DEV).bfloat16()
    gam, bet = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    y = g.new(rows, D)
    mean, rstd = g.new(rows, dtype=torch.float32), g.new(rows, dtype=torch.float32)
    ops.layernorm_fwd(x, y, gam, bet, mean, rstd, rows, D, drop_p=0.1, drop_seed=1, drop_stream=2)
    dx, dx2 = g.new(rows, D), g.new(rows, D)
    dg, db, dbias = (g.new(D, dtype=torch.float32, fill=0.0) for _ in range(3))
    ops.layernorm_bwd(y, x, gam, mean, rstd, dx, dg, db, rows, D, dbias=dbias, dx2=dx2, dx2_drop=(0.15, 1, 3))
    # remapped variant (final LayerNorm on the visual rows)
    Bt, n, S = 3, 50, 61
    xs = torch.randn(Bt * S, D, device=DEV).bfloat16()
    vis = g.new(Bt * n, D)
    m2, r2 = g.new(Bt * n, dtype=torch.float32), g.new(Bt * n, dtype=torch.float32)
    ops.layernorm_fwd(xs, vis, gam, bet, m2, r2, Bt * n, D, in_map=(n, S, 0))
    dxs = g.new(Bt * S, D, fill=0.0)
    ops.layernorm_bwd(vis, xs, gam, m2, r2, dxs, dg, db, Bt * n, D, in_map=(n, S, 0))
    cs = g.new(D, dtype=torch.float32, fill=0.0)
    ops.colsum(x, cs, rows, D)
    gat = g.new(Bt * n, D)
    ops.rows_gather(xs, gat, Bt * n, D, in_map=(n, S, 0), colsum=cs, drop_p=0.1, drop_seed=1, drop_stream=5)
    # layout passes, both memory formats, odd grids
    for fmt in (torch.contiguous_format, torch.channels_last):
        for (B, C, Hh, Ww, p, dt) in ((2, 70, 12, 20, 2, torch.bfloat16), (1, 48, 16, 136, 4, torch.float32), (2, 33, 5, 7, 1, torch.float32)):
            if (C * p * p) % 2:
                continue
            f = torch.randn(B, C, Hh, Ww, device=DEV).to(dt).contiguous(memory_format=fmt)
            tok = g.new(B * (Hh // p) * (Ww // p), C * p * p)
            ops.patchify(f, p, tok)
            back = g.new(B * C * Hh * Ww, dtype=dt).view(B, Hh, Ww, C).permute(0, 3, 1, 2) if fmt == torch.channels_last else g.new(B, C, Hh, Ww, dtype=dt)
            ops.fold(tok, back, p)
    # multi-tensor casts and the optimizer step (ragged sizes)
    srcs = [torch.randn(r, c, device=DEV) for r, c in ((7, 8), (33, 24), (5, 712))]
    dsts = [g.new(*t.shape) for t in srcs]
    ops.cast_pad_multi([(s, d_, s.shape[0], s.shape[1], 0, 0, 0, 0) for s, d_ in zip(srcs, dsts)])
    ps = [torch.nn.Parameter(g.new(n_, dtype=torch.float32, fill=0.5)) for n_ in (1, 7, 897, 4096)]
    opt = optim.FusedRAdam(ps, lr=1e-3, degenerated_to_sgd=True, max_grad_norm=1.0)
    for p_ in ps:
        p_.grad = torch.randn_like(p_)
        opt.state[p_]["step"] = 0
        opt.state[p_]["exp_avg"] = g.new(p_.numel(), dtype=torch.float32, fill=0.0)
        opt.state[p_]["exp_avg_sq"] = g.new(p_.numel(), dtype=torch.float32, fill=0.0)
    opt.step()
    g.check()
