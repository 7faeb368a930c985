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
