"""CPU: host logic added in round 2 -- bf16 weight-cache rules, the FPN-from-laterals helper against torchvision, SM shares,
optimizer / box-head / encoder loud failures without CUDA, the staged-reference manifest, and the bench reference arm's JSON
contract (runs the unmodified reference module on one sample when a reference tree is available)."""
import copy
import json
import os
import subprocess
import sys
from collections import OrderedDict

import pytest
import torch
from torch import nn

from oracle import build_ref, ref_loader
from transfusion_b200 import weight_cache
from transfusion_b200.configs import default_fusion_cfg
from transfusion_b200.cross_fusion import CrossFusionBoxWrapper
from transfusion_b200.obj_detection import FusedBoxHead
from transfusion_b200.obj_detection.fpn import fpn_from_laterals

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_weight_cache_trust_rules():
    p = nn.Parameter(torch.randn(6, 8))
    casts = []
    a = weight_cache.bf16_weight(p, 6, 8, casts, trust_version=True, dev=p.device)
    assert len(casts) == 1 and p._xf_bf16[3] == "cast"
    # inference: same version -> reused, no new cast job
    b = weight_cache.bf16_weight(p, 6, 8, casts, trust_version=True, dev=p.device)
    assert b is a and len(casts) == 1
    # training without the fused optimizer: never trusted (p.data writes do not bump the version counter)
    c = weight_cache.bf16_weight(p, 6, 8, casts, trust_version=False, dev=p.device)
    assert len(casts) == 2 and c.data_ptr() == a.data_ptr()          # re-cast into the same buffer
    # a copy tagged by the fused optimizer is consumed exactly once in training
    p._xf_bf16 = (p._version, a, True, "opt")
    d = weight_cache.bf16_weight(p, 6, 8, casts, trust_version=False, dev=p.device)
    assert d is a and len(casts) == 2 and p._xf_bf16[3] == "used"
    e = weight_cache.bf16_weight(p, 6, 8, casts, trust_version=False, dev=p.device)
    assert len(casts) == 3
    # a version bump invalidates in inference too
    with torch.no_grad():
        p.add_(1.0)
    weight_cache.bf16_weight(p, 6, 8, casts, trust_version=True, dev=p.device)
    assert len(casts) == 4
    m = nn.Linear(8, 6)
    m.weight._xf_bf16 = (0, a, True, "cast")
    weight_cache.invalidate(m)
    assert not hasattr(m.weight, "_xf_bf16")


def test_mode_switch_drops_the_weight_cache():
    cfg = default_fusion_cfg(64, n_levels=1, num_layers=[1], patch=[1])
    m = CrossFusionBoxWrapper(ref_loader.FakeRCNN([(4, 4)], [8], 9, 6), cfg, {"text_pooling": "x", "train_ep": -1},
                              criterion={"lm": 0}, narr_pooling_layer=ref_loader.PassThroughPooling())
    w = m.patches_to_token[0].weight
    w._xf_bf16 = (w._version, torch.zeros(1), True, "cast")
    m.eval()
    assert not hasattr(w, "_xf_bf16")
    w._xf_bf16 = (w._version, torch.zeros(1), True, "cast")
    m.eval()                       # no mode change: kept
    assert hasattr(w, "_xf_bf16")
    m.train()
    assert not hasattr(w, "_xf_bf16")
    assert m.precision == "bf16"
    with pytest.raises(ValueError):
        CrossFusionBoxWrapper(ref_loader.FakeRCNN([(4, 4)], [8], 9, 6), copy.deepcopy(default_fusion_cfg(64, n_levels=1, num_layers=[1], patch=[1])),
                              {"text_pooling": "x", "train_ep": -1}, criterion={"lm": 0},
                              narr_pooling_layer=ref_loader.PassThroughPooling(), precision="fp16")


def test_constructor_signature_matches_reference_keywords():
    import inspect
    sig = inspect.signature(CrossFusionBoxWrapper.__init__)
    assert list(sig.parameters)[:5] == ["self", "rcnn_model", "cross_layer_args", "narr_embed_args", "criterion"]   # cross_f_box_wrapper.py:42-44
    assert sig.parameters["criterion"].default is None
    with pytest.raises(ValueError):
        CrossFusionBoxWrapper(ref_loader.FakeRCNN([(4, 4)], [8], 9, 6))   # the reference's defaults are unconstructible


def test_fpn_from_laterals_equals_torchvision_fpn():
    from torchvision.ops import FeaturePyramidNetwork
    from torchvision.ops.feature_pyramid_network import LastLevelMaxPool
    torch.manual_seed(0)
    fpn = FeaturePyramidNetwork([8, 16, 32], 12, extra_blocks=LastLevelMaxPool())
    x = OrderedDict((str(i), torch.randn(2, c, 32 >> i, 48 >> i)) for i, c in enumerate((8, 16, 32)))
    ref = fpn(x)
    lat = OrderedDict((k, fpn.inner_blocks[i](v)) for i, (k, v) in enumerate(x.items()))
    got = fpn_from_laterals(fpn, lat)
    assert list(got.keys()) == list(ref.keys())
    for k in ref:
        assert torch.allclose(got[k], ref[k], atol=1e-6)


def test_fuse_fpn_inner_validates_the_fpn():
    from torchvision.ops import FeaturePyramidNetwork
    cfg = default_fusion_cfg(64, n_levels=2, num_layers=[1, 1], patch=[2, 1])
    m = CrossFusionBoxWrapper(ref_loader.FakeRCNN([(8, 8), (4, 4)], [8, 16], 9, 6), cfg, {"text_pooling": "x", "train_ep": -1},
                              criterion={"lm": 0}, narr_pooling_layer=ref_loader.PassThroughPooling())
    with pytest.raises(ValueError):
        m.fuse_fpn_inner(FeaturePyramidNetwork([8, 16, 32], 12))        # one inner block too many
    bad = FeaturePyramidNetwork([8, 16], 12)
    bad.inner_blocks[0] = nn.Sequential(nn.Conv2d(8, 12, 3, padding=1))
    with pytest.raises(NotImplementedError):
        m.fuse_fpn_inner(bad)
    m.fuse_fpn_inner(FeaturePyramidNetwork([8, 16], 12))
    assert "_xf_fpn" in m.__dict__ and "_xf_fpn" not in dict(m.named_modules())    # not re-registered: state_dict unchanged
    m.fuse_fpn_inner(None)


def test_new_modules_fail_loudly_without_cuda():
    from transfusion_b200._lib import XfError
    from transfusion_b200.narration_embeds import XfLinear, bert_encoder_forward
    from transfusion_b200.optim import FusedRAdam, grad_sqnorm
    p = nn.Parameter(torch.randn(8))
    p.grad = torch.randn(8)
    with pytest.raises(XfError):
        FusedRAdam([p], lr=1e-3).step()
    with pytest.raises(XfError):
        grad_sqnorm([p.grad], out=torch.zeros(1))
    with pytest.raises(RuntimeError):
        FusedBoxHead(16, 8, 3, 2)(torch.randn(2, 16))
    with pytest.raises(RuntimeError):
        XfLinear(8, 8)(torch.randn(2, 8))
    with pytest.raises(RuntimeError):
        bert_encoder_forward(None, torch.zeros(1, 4, dtype=torch.long))
    with pytest.raises(ValueError):
        FusedRAdam([p], lr=-1.0)


def test_staged_reference_manifest_is_intact():
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref")):
        pytest.skip("oracle/_ref not staged (no reference tree in this environment)")
    assert build_ref.verify()
    with open(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")) as f:
        man = json.load(f)["files"]
    assert "modeling/cross_fusion/ego_fusion/cross_f_box_wrapper.py" in man and len(man) >= 12
    # nothing under oracle/_ref is tracked: no reference source enters the history
    tracked = subprocess.run(["git", "ls-files", "oracle/_ref"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    assert tracked == ""


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")
def test_bench_reference_arm_json_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-batch", "1", "--lang-len", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "cross_fusion fwd+bwd samples/sec" and line["unit"] == "samples/s"
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"]["per_gpu_batch"] == 13 and "ego4dv2" in line["config"]["workload"]
