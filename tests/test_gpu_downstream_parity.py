"""GPU: parity of the quantities the north star names downstream of the fused feature maps —
noun/verb logits and box regressions — obtained by running STOCK torchvision FPN -> MultiScaleRoIAlign
-> TwoMLPHead -> box_regressor / noun_classifier / verb_classifier (the structure of the reference's
roi_wrappers.py:194-214 and faster_rcnn_wrapper.py:419-421, with fixed seeded proposals instead of
RPN/NMS) on top of (a) the CPU oracle's fused features and (b) the CUDA path's fused features
(SURVEY §8a row A10).  The head is not part of the hot path; it only propagates the parity check."""
from collections import OrderedDict

import pytest
import torch
from torch import nn
from torchvision.models.detection.faster_rcnn import TwoMLPHead
from torchvision.ops import FeaturePyramidNetwork, MultiScaleRoIAlign

from oracle import ref_math
from tests.fusion_testlib import build_module, param_dict, run_module
from tests.golden_utils import rel_fro

pytestmark = pytest.mark.gpu


class Head(nn.Module):
    def __init__(self, channels, nouns=129, verbs=82, rep=256):
        super().__init__()
        self.fpn = FeaturePyramidNetwork(channels, 64)
        self.pool = MultiScaleRoIAlign([str(i) for i in range(len(channels))], 7, 2)
        self.box_head = TwoMLPHead(64 * 7 * 7, rep)
        self.box_regressor = nn.Linear(rep, 4 * nouns)
        self.noun_classifier = nn.Linear(rep, nouns)
        self.verb_classifier = nn.Linear(rep, verbs)

    def forward(self, feats, proposals, image_shapes):
        f = self.fpn(OrderedDict((k, feats[k]) for k in sorted(feats, key=int)))
        x = self.box_head(self.pool(f, proposals, image_shapes))
        return {"class_logits": self.noun_classifier(x), "verb_logits": self.verb_classifier(x),
                "box_regression": self.box_regressor(x)}


def test_roi_logits_and_box_regression_parity():
    D, heads, B, L = 896, 4, 2, 16
    image = (128, 192)
    strides, channels, patch, layers = [8, 16, 32], [32, 64, 128], [4, 2, 1], [1, 1, 1]
    shapes = [(image[0] // s, image[1] // s) for s in strides]
    m = build_module(D, shapes, channels, patch, layers, heads, seed=11)
    m.train()
    g = torch.Generator().manual_seed(12)
    feats = {str(i): torch.relu(torch.randn(B, c, h, w, generator=g)) for i, ((h, w), c) in enumerate(zip(shapes, channels))}
    lang = 0.5 * torch.randn(B, L, D, generator=g)
    mask = torch.ones(B, L, dtype=torch.int64)
    mask[1, 7:] = 0
    torch.manual_seed(13)
    head = Head(channels)
    proposals = []
    for b in range(B):
        xy = torch.rand(24, 2, generator=g) * torch.tensor([image[1] - 40.0, image[0] - 40.0])
        wh = 16 + torch.rand(24, 2, generator=g) * 24
        proposals.append(torch.cat([xy, xy + wh], dim=1))
    image_shapes = [image] * B

    with torch.no_grad():
        sd = {k: v.detach().cpu() for k, v in param_dict(m).items()}
        ref_feats, _ = ref_math.cross_fusion_forward(feats, lang, mask, sd, patch, heads, layers)
        ref_out = head(ref_feats, proposals, image_shapes)
        got_feats, _ = run_module(m, {k: v.cuda() for k, v in feats.items()}, lang.cuda(), mask.cuda())
        got_out = head({k: v.float().cpu() for k, v in got_feats.items()}, proposals, image_shapes)
    for k in ("class_logits", "verb_logits", "box_regression"):
        assert ref_out[k].shape[0] == 48
        r = rel_fro(got_out[k], ref_out[k])
        ma = float((got_out[k] - ref_out[k]).abs().max())
        rms = float(ref_out[k].pow(2).mean().sqrt())
        assert r < 1e-2, f"{k}: rel-Frobenius {r:.3e}"
        assert ma < 5e-2 * max(1.0, rms), f"{k}: max-abs {ma:.3e}"


def test_roi_logits_with_fused_box_head_at_shipped_widths():
    """The same propagation at the REAL head widths (FPN 256 channels -> 12544-wide fc6, R = 1280, 129 / 82 classes,
    ego_vis_det_ego4dv2.yml:6) with the dense head running on xf_gemm (transfusion_b200.obj_detection.FusedBoxHead sharing
    the stock modules' parameters): CUDA fusion path + CUDA box head vs CPU oracle + stock torch head."""
    import types
    from transfusion_b200.obj_detection import FusedBoxHead

    D, heads, B, L = 896, 4, 2, 16
    image = (128, 192)
    strides, channels, patch, layers = [8, 16, 32], [32, 64, 128], [4, 2, 1], [1, 1, 1]
    shapes = [(image[0] // s, image[1] // s) for s in strides]
    m = build_module(D, shapes, channels, patch, layers, heads, seed=21)
    m.train()
    g = torch.Generator().manual_seed(22)
    feats = {str(i): torch.relu(torch.randn(B, c, h, w, generator=g)) for i, ((h, w), c) in enumerate(zip(shapes, channels))}
    lang = 0.5 * torch.randn(B, L, D, generator=g)
    mask = torch.ones(B, L, dtype=torch.int64)
    mask[0, 11:] = 0
    torch.manual_seed(23)
    head = Head(channels, rep=1280)
    head.fpn = FeaturePyramidNetwork(channels, 256)
    head.box_head = TwoMLPHead(256 * 7 * 7, 1280)
    proposals = []
    for b in range(B):
        xy = torch.rand(64, 2, generator=g) * torch.tensor([image[1] - 40.0, image[0] - 40.0])
        wh = 16 + torch.rand(64, 2, generator=g) * 24
        proposals.append(torch.cat([xy, xy + wh], dim=1))
    image_shapes = [image] * B
    with torch.no_grad():
        sd = {k: v.detach().cpu() for k, v in param_dict(m).items()}
        ref_feats, _ = ref_math.cross_fusion_forward(feats, lang, mask, sd, patch, heads, layers)
        ref_out = head(ref_feats, proposals, image_shapes)
        got_feats, _ = run_module(m, {k: v.cuda() for k, v in feats.items()}, lang.cuda(), mask.cuda())
        head_gpu = head.cuda()
        f = head_gpu.fpn(OrderedDict((k, got_feats[k].float()) for k in sorted(got_feats, key=int)))
        pooled = head_gpu.pool(f, [p.cuda() for p in proposals], image_shapes)
        roi = types.SimpleNamespace(roi_head_wrap=types.SimpleNamespace(box_head=head_gpu.box_head), dropout_1=nn.Identity(),
                                    classif_dropout=nn.Identity(), box_regressor=nn.Sequential(nn.Identity(), head_gpu.box_regressor),
                                    noun_classifier=head_gpu.noun_classifier, verb_classifier=head_gpu.verb_classifier)
        fused_head = FusedBoxHead.from_roi_heads(roi).eval()
        box, noun, verb = fused_head(pooled)
    got_out = {"class_logits": noun.cpu(), "verb_logits": verb.cpu(), "box_regression": box.cpu()}
    for k in ("class_logits", "verb_logits", "box_regression"):
        assert ref_out[k].shape[0] == 128
        r = rel_fro(got_out[k], ref_out[k])
        assert r < 1e-2, f"{k}: rel-Frobenius {r:.3e}"
