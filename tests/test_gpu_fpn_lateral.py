"""GPU: SURVEY 8f N1 -- the FPN lateral (inner 1x1 conv, faster_rcnn_wrapper.py:419-421 -> torchvision
FeaturePyramidNetwork.inner_blocks) folded into the back-projection GEMM (CrossFusionBoxWrapper.fuse_fpn_inner): FPN
outputs and every gradient (fusion parameters, FPN inner and output convs, inputs) against the CPU oracle's fused maps fed
through the SAME stock torchvision FPN in fp32.  Bounds: rel-Frobenius <= 1e-2."""
import copy
from collections import OrderedDict

import pytest
import torch
from torchvision.ops import FeaturePyramidNetwork
from torchvision.ops.feature_pyramid_network import LastLevelMaxPool

from oracle import ref_math
from tests.fusion_testlib import build_module, param_dict, run_module
from tests.golden_utils import rel_fro

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,patch,channels,out_ch", [(896, [4, 2, 1], [32, 64, 128], 256), (256, [2, 1], [48, 96], 64)])
def test_fused_fpn_lateral_matches_stock_fpn_on_oracle_features(D, patch, channels, out_ch):
    heads, B, L = 4, 2, 12
    image = (128, 192)
    strides = [8, 16, 32][:len(channels)] if len(channels) == 3 else [16, 32]
    shapes = [(image[0] // s, image[1] // s) for s in strides]
    layers = [1] * len(channels)
    m = build_module(D, shapes, channels, patch, layers, heads, seed=41)
    m.train()
    torch.manual_seed(42)
    fpn_cpu = FeaturePyramidNetwork(channels, out_ch, extra_blocks=LastLevelMaxPool())
    fpn_gpu = copy.deepcopy(fpn_cpu).cuda()
    g = torch.Generator().manual_seed(43)
    feats = {str(i): torch.relu(torch.randn(B, c, h, w, generator=g)) for i, ((h, w), c) in enumerate(zip(shapes, channels))}
    lang = 0.5 * torch.randn(B, L, D, generator=g)
    mask = torch.ones(B, L, dtype=torch.int64)
    mask[1, 5:] = 0

    # reference: oracle fused maps -> stock FPN (fp32, CPU)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in param_dict(m).items()}
    f_cpu = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    l_cpu = lang.clone().requires_grad_(True)
    fused, _ = ref_math.cross_fusion_forward(f_cpu, l_cpu, mask, sd, patch, heads, layers)
    ref = fpn_cpu(OrderedDict((k, fused[k]) for k in sorted(fused, key=int)))
    cot = {k: torch.randn(v.shape, generator=g) for k, v in ref.items()}
    sum((ref[k] * cot[k]).sum() for k in ref).backward()

    m.fuse_fpn_inner(fpn_gpu)
    f_gpu = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
    l_gpu = lang.cuda().requires_grad_(True)
    out, _ = run_module(m, f_gpu, l_gpu, mask.cuda())
    assert list(out.keys()) == list(ref.keys())          # incl. the extra "pool" level
    for k in ref:
        assert out[k].shape == ref[k].shape
        assert rel_fro(out[k].detach().float().cpu(), ref[k].detach()) < 1e-2, f"fpn output {k}"
    sum((out[k].float() * cot[k].cuda()).sum() for k in out).backward()
    torch.cuda.synchronize()
    for k in f_gpu:
        assert rel_fro(f_gpu[k].grad.cpu(), f_cpu[k].grad) < 1e-2, f"grad features.{k}"
    assert rel_fro(l_gpu.grad.cpu(), l_cpu.grad) < 1e-2
    worst = ("", 0.0)
    for k, p in param_dict(m).items():
        if k.endswith("heatmap_token"):
            continue
        r = rel_fro(p.grad.cpu(), sd[k].grad)
        if r > worst[1]:
            worst = (k, r)
    assert worst[1] < 1e-2, f"worst fusion param grad {worst}"
    for (k, pg), (_, pc) in zip(fpn_gpu.named_parameters(), fpn_cpu.named_parameters()):
        assert pg.grad is not None, k
        assert rel_fro(pg.grad.cpu(), pc.grad) < 1e-2, f"fpn param grad {k}"
    # undo: the module returns fused maps again and the caller's apply_fpn is used
    m.fuse_fpn_inner(None)
    with torch.no_grad():
        out2, _ = run_module(m, {k: v.cuda() for k, v in feats.items()}, lang.cuda(), mask.cuda())
    assert out2["0"].shape == feats["0"].shape


def test_stock_fpn_on_fused_maps_forward_backward():
    """The reference data flow (no lateral fusion): the stock torchvision FPN consumes the fused maps on the caller's stream
    and its backward feeds the level functions, which replay on their side streams.  FPN outputs, fusion-parameter and FPN
    gradients against the CPU oracle + the same FPN."""
    D, heads, B, L = 256, 4, 2, 12
    image = (128, 192)
    patch, channels, strides, layers = [2, 1], [48, 96], [16, 32], [1, 1]
    shapes = [(image[0] // s, image[1] // s) for s in strides]
    m = build_module(D, shapes, channels, patch, layers, heads, seed=45)
    m.train()
    torch.manual_seed(46)
    fpn_cpu = FeaturePyramidNetwork(channels, 64, extra_blocks=LastLevelMaxPool())
    fpn_gpu = copy.deepcopy(fpn_cpu).cuda()
    m.rcnn_model.apply_fpn = lambda d: {**d, "features": fpn_gpu(OrderedDict((k, d["features"][k].float()) for k in sorted(d["features"], key=int)))}
    g = torch.Generator().manual_seed(47)
    feats = {str(i): torch.relu(torch.randn(B, c, h, w, generator=g)) for i, ((h, w), c) in enumerate(zip(shapes, channels))}
    lang = 0.5 * torch.randn(B, L, D, generator=g)
    mask = torch.ones(B, L, dtype=torch.int64)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in param_dict(m).items()}
    fused, _ = ref_math.cross_fusion_forward({k: v.clone() for k, v in feats.items()}, lang, mask, sd, patch, heads, layers)
    ref = fpn_cpu(OrderedDict((k, fused[k]) for k in sorted(fused, key=int)))
    cot = {k: torch.randn(v.shape, generator=g) for k, v in ref.items()}
    sum((ref[k] * cot[k]).sum() for k in ref).backward()
    for step in range(3):   # several steps: the run-ahead limiter and the side streams are exercised across iterations
        m.zero_grad(set_to_none=True)
        fpn_gpu.zero_grad(set_to_none=True)
        out, _ = run_module(m, {k: v.cuda() for k, v in feats.items()}, lang.cuda(), mask.cuda())
        sum((out[k].float() * cot[k].cuda()).sum() for k in out).backward()
    torch.cuda.synchronize()
    for k in ref:
        assert rel_fro(out[k].detach().float().cpu(), ref[k].detach()) < 1e-2, k
    worst = max((rel_fro(p.grad.cpu(), sd[k].grad), k) for k, p in param_dict(m).items() if not k.endswith("heatmap_token"))
    assert worst[0] < 1e-2, worst
    for (k, pg), (_, pc) in zip(fpn_gpu.named_parameters(), fpn_cpu.named_parameters()):
        assert rel_fro(pg.grad.cpu(), pc.grad) < 1e-2, k
