"""Shared helpers for the GPU parity tests: build the CUDA-backed CrossFusionBoxWrapper with a fake
rcnn_model / pass-through pooling layer (same stand-ins the oracle loader uses)."""
import copy

import torch

from oracle.ref_loader import FakeRCNN, PassThroughPooling
from transfusion_b200.configs import default_fusion_cfg
from transfusion_b200.cross_fusion import CrossFusionBoxWrapper


def build_module(token_dim, shapes, channels, patch, layers, heads, lm=False, dropout=False, device="cuda",
                 noun_classes=9, verb_classes=6, seed=0, use_lm_f=True, forward_language_f=False):
    kw = {} if dropout else dict(patch_dropout=0.0, token_dropout=0.0, backproj_dropout=0.0)
    kw.update(use_lm_f=use_lm_f, forward_language_f=forward_language_f)
    cfg = default_fusion_cfg(token_dim, n_levels=len(shapes), num_layers=layers, num_heads=heads, patch=patch, **kw)
    torch.manual_seed(seed)
    rcnn = FakeRCNN(shapes, channels, noun_classes, verb_classes)
    m = CrossFusionBoxWrapper(rcnn, copy.deepcopy(cfg), {"text_pooling": "x", "train_ep": -1},
                              criterion={"lm": 1 if lm else 0}, narr_pooling_layer=PassThroughPooling())
    return m.to(device)


def run_module(m, features, lang, att_mask):
    m.rcnn_model.features = features
    out = m({"image": None, "language_f": (lang, att_mask)})
    return out["features"], out.get("lm")


def param_dict(m):
    return {k: p for k, p in m.named_parameters() if not k.startswith(("rcnn_model", "narr_pooling_layer"))}
