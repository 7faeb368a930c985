"""GPU: the fp32-tolerance mode (CrossFusionBoxWrapper(..., precision="fp32"): 3-way bf16 split GEMMs on the tcgen05 kernel,
fp32 softmax / LayerNorm) against the fp32 CPU oracle.  north_star: "about 1e-5 relative in fp32" -- asserted as
rel-Frobenius <= 1e-5 on the fused features at BASELINE config 1's shape (C5 level, D = 896, L = 64, ragged lengths) and on
a two-level / odd-grid case; forward only (autograd through the mode raises)."""
import copy

import pytest
import torch

from oracle import ref_math
from oracle.ref_loader import FakeRCNN, PassThroughPooling
from tests.fusion_testlib import param_dict, run_module
from tests.golden_utils import rel_fro
from transfusion_b200 import ops
from transfusion_b200.configs import default_fusion_cfg
from transfusion_b200.cross_fusion import CrossFusionBoxWrapper

pytestmark = pytest.mark.gpu
DEV = "cuda"
FP32_REL = 1e-5


def _module(D, shapes, channels, patch, layers, heads, seed, lm=False):
    cfg = default_fusion_cfg(D, n_levels=len(shapes), num_layers=layers, num_heads=heads, patch=patch)
    torch.manual_seed(seed)
    m = CrossFusionBoxWrapper(FakeRCNN(shapes, channels, 9, 6), copy.deepcopy(cfg), {"text_pooling": "x", "train_ep": -1},
                              criterion={"lm": 1 if lm else 0}, narr_pooling_layer=PassThroughPooling(), precision="fp32")
    return m.to(DEV).eval()


def test_split3_gemm_is_fp32_accurate():
    torch.manual_seed(0)
    x = torch.randn(300, 712, device=DEV)
    w = torch.randn(896, 712, device=DEV) * 0.05
    bias = torch.randn(896, device=DEV)
    out = torch.zeros(300, 896, device=DEV)
    ops.gemm(ops.split3(x, 0, bias_cols=8), ops.split3(w, 1, bias=bias, bias_cols=8), out, M=300, N=896, K=6 * 712 + 8, accumulate=True,
             split_k=12)
    ref = (x.double() @ w.double().t() + bias.double()).float()
    assert rel_fro(out, ref) < 2e-6
    one = torch.zeros(300, 896, device=DEV)   # a single 6K-long accumulator chain: the tensor core's truncating adds show
    ops.gemm(ops.split3(x, 0, bias_cols=8), ops.split3(w, 1, bias=bias, bias_cols=8), one, M=300, N=896, K=6 * 712 + 8, accumulate=True)
    assert rel_fro(one, ref) > rel_fro(out, ref)
    bf = torch.empty(300, 896, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x.bfloat16(), w.bfloat16(), bf, M=300, N=896, K=712)
    assert rel_fro(bf.float(), ref) > 1e-3      # the plain bf16 product is three orders of magnitude further away


@pytest.mark.parametrize("D,shapes,channels,patch,layers,B,L,lens", [
    (896, [(24, 32)], [2048], [1], [4], 2, 64, [64, 48]),                    # BASELINE config 1: C5 level, 4 layers (B reduced 8 -> 2)
    (256, [(30, 38), (8, 12)], [16, 24], [2, 1], [2, 1], 2, 16, [0, 16]),    # odd grid (n = 285), a sample without language
])
def test_fp32_mode_matches_fp32_oracle(D, shapes, channels, patch, layers, B, L, lens):
    m = _module(D, shapes, channels, patch, layers, 4, seed=51)
    g = torch.Generator().manual_seed(52)
    feats = {str(i): torch.relu(torch.randn(B, c, h, w, generator=g)) for i, ((h, w), c) in enumerate(zip(shapes, channels))}
    lang = 0.5 * torch.randn(B, L, D, generator=g)
    mask = torch.zeros(B, L, dtype=torch.int64)
    for b, n_ in enumerate(lens):
        mask[b, :n_] = 1
    sd = {k: v.detach().cpu() for k, v in param_dict(m).items()}
    ref, _ = ref_math.cross_fusion_forward({k: v.clone() for k, v in feats.items()}, lang, mask, sd, patch, 4, layers)
    with torch.no_grad():
        out, _ = run_module(m, {k: v.cuda() for k, v in feats.items()}, lang.cuda(), mask.cuda())
    for k in ref:
        r = rel_fro(out[k].float().cpu(), ref[k])
        assert r < FP32_REL, f"fp32 mode features.{k}: rel-Frobenius {r:.3e}"


def test_fp32_mode_is_forward_only():
    m = _module(256, [(8, 12)], [16], [1], [1], 4, seed=53)
    feats = {"0": torch.relu(torch.randn(1, 16, 8, 12)).cuda()}
    with pytest.raises(NotImplementedError):
        run_module(m, feats, torch.randn(1, 4, 256, device=DEV), torch.ones(1, 4, dtype=torch.int64, device=DEV))
