"""The epilogue-heavy K = 896 GEMMs of one encoder layer at the Ego4Dv2 level-0 shape, for `ncu --set full`
captures: FFN1 forward (bias + pre-activation store + GELU + dropout), out-proj forward (bias + dropout +
residual), FFN2 dgrad (GELU' from the saved pre-activation + dropout mask).  Dev tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from transfusion_b200 import ops

B, S, D = int(os.environ.get("XF_B", 13)), 3136, 896
M = B * S
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, D, device=dev).bfloat16()
w1 = (torch.randn(2 * D, D, device=dev) * 0.03).bfloat16()
b1 = torch.randn(2 * D, device=dev)
wo = (torch.randn(D, D, device=dev) * 0.03).bfloat16()
bo = torch.randn(D, device=dev)
w2 = (torch.randn(D, 2 * D, device=dev) * 0.03).bfloat16()
u = torch.empty(M, 2 * D, device=dev, dtype=torch.bfloat16)
h = torch.empty(M, 2 * D, device=dev, dtype=torch.bfloat16)
y = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
dy = torch.randn(M, D, device=dev).bfloat16()
dh = torch.empty(M, 2 * D, device=dev, dtype=torch.bfloat16)
for it in range(int(os.environ.get("XF_ITERS", 2))):
    ops.gemm(x, w1, h, M=M, N=2 * D, K=D, bias=b1, act=1, preact_out=u, drop_p=0.1, drop_seed=1, drop_stream=2)
    ops.gemm(x, wo, y, M=M, N=D, K=D, bias=bo, drop_p=0.1, drop_seed=1, drop_stream=3, residual=x)
    ops.gemm(dy, w2, dh, M=M, N=2 * D, K=D, b_mn_major=True, dact_in=u, drop_p=0.1, drop_seed=1, drop_stream=2)
    torch.cuda.synchronize()
print("profile_gemm_epilogues ok")
