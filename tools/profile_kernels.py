"""One launch of each hot kernel at the Ego4Dv2 level-0 shape (B=13, S=3136, D=896, 4 heads of 224),
for `ncu --set full` captures (profiles/).  Dev tool."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from transfusion_b200 import ops

B, S, D, H = int(os.environ.get("XF_B", 13)), int(os.environ.get("XF_S", 3136)), 896, 4
d = D // H
M = B * S
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, D, device=dev).bfloat16()
w_in = (torch.randn(3 * D, D, device=dev) * 0.03).bfloat16()
b_in = torch.randn(3 * D, device=dev)
qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
w1 = (torch.randn(2 * D, D, device=dev) * 0.03).bfloat16()
b1 = torch.randn(2 * D, device=dev)
u = torch.empty(M, 2 * D, device=dev, dtype=torch.bfloat16)
w2 = (torch.randn(D, 2 * D, device=dev) * 0.03).bfloat16()
wo = (torch.randn(D, D, device=dev) * 0.03).bfloat16()
bo = torch.randn(D, device=dev)
h = torch.empty(M, 2 * D, device=dev, dtype=torch.bfloat16)
for it in range(2):
    # forward GEMMs
    ops.gemm(x, w_in, qkv, M=M, N=3 * D, K=D, bias=b_in)
    ops.gemm(x, w1, h, M=M, N=2 * D, K=D, bias=b1, act=1, preact_out=u, drop_p=0.15, drop_seed=1, drop_stream=2)
    # FFN2 dgrad: GELU' from the saved pre-activation + dropout mask (EPI_DGELU)
    dh = torch.empty(M, 2 * D, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, w2, dh, M=M, N=2 * D, K=D, b_mn_major=True, dact_in=u, drop_p=0.15, drop_seed=1, drop_stream=2)
    # out-proj forward: bias + dropout + residual (EPI_LINEAR)
    y1 = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, wo, y1, M=M, N=D, K=D, bias=bo, drop_p=0.1, drop_seed=1, drop_stream=3, residual=x)
    # dgrad (B operand MN-major) and wgrad (both MN-major, split-K, fp32 atomics)
    dx = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    ops.gemm(h, w1, dx, M=M, N=D, K=2 * D, b_mn_major=True, residual=x)
    gw = torch.zeros(2 * D, D, device=dev)
    ops.gemm(h, x, gw, M=2 * D, N=D, K=M, a_mn_major=True, b_mn_major=True, accumulate=True, split_k=6)
    # attention
    Sp = (S + 127) // 128 * 128
    att = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, Sp, device=dev)
    kpm = torch.zeros(B, S, dtype=torch.uint8, device=dev)
    kpm[:, S - 20:] = 1
    kw = dict(B=B, H=H, Sq=S, Sk=S, dp=d, scale=1 / math.sqrt(d), drop_p=0.15, drop_seed=3, drop_stream=4)
    ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], att, lse, key_padding_mask=kpm, kpm_start=S - 64, **kw)
    datt = torch.randn(M, D, device=dev).bfloat16()
    delta = torch.empty(B, H, Sp, device=dev)
    ops.attn_delta(att, datt, delta, B, S, H, d)
    dqkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    ws = None
    if int(os.environ.get("XF_ATTN_BWD_WS", "1")):
        ws = torch.empty(ops.attn_bwd_workspace_bytes(B, H, S, S), device=dev, dtype=torch.uint8)
    ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], datt, lse, delta, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                 key_padding_mask=kpm, workspace=ws, **kw)
    # HBM-bound
    y = torch.empty_like(x)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    ops.layernorm_fwd(x, y, g, b, mean, rstd, M, D)
    dg, db, dbias = torch.zeros(D, device=dev), torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    ops.layernorm_bwd(y, x, g, mean, rstd, dx, dg, db, M, D, dbias=dbias)
    torch.cuda.synchronize()
print("profile_kernels ok")
