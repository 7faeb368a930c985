"""Stand-ins for the two neighbours of the fusion path that are outside it: the
Faster-R-CNN wrapper that produces the C2..C5 feature maps and consumes the fused ones, and the
language-context pooling layer.  Used by bench.py and the examples to drive
``CrossFusionBoxWrapper.forward`` with synthetic features; not needed inside the reference tree."""
from __future__ import annotations

import copy

import torch
from torch import nn

from .configs import WORKLOADS, default_fusion_cfg, level_shapes


class FeatureProviderRCNN(nn.Module):
    """Implements what CrossFusionBoxWrapper needs from ``rcnn_model``: ``forward_features`` hands
    back the feature maps stored in ``self.features``; FPN / RPN / RoI are identity so the fused maps
    are returned directly."""

    def __init__(self, shapes, channels, noun_classes=129, verb_classes=82):
        super().__init__()
        self._shapes, self._channels = list(shapes), list(channels)
        self.noun_classes, self.verb_classes = noun_classes, verb_classes
        self.features = None

    def get_dsampled_shapes(self):
        return self._shapes

    def get_features_out_channels(self):
        return self._channels

    def forward_features(self, images, targets=None):
        return {"features": dict(self.features)}

    def apply_fpn(self, d):
        return d

    def apply_rpn_roi_on_features(self, d):
        return d

    def call_model_epoch_triggers(self, epoch):
        pass


class PassThroughPooling(nn.Module):
    """Returns (embeddings [B,L,D], None, att_mask [B,L] 1 = valid): SBertLayer's contract
    (narr_pooling_layers.py:199-202) with the embeddings supplied by the caller."""

    def forward(self, lang, pad_mask=False):
        emb, mask = lang
        return emb, None, (mask if pad_mask else None)


def build_workload_module(name: str, device="cuda", dropout: bool = True, lm: bool = False, seed: int = 0):
    from .cross_fusion import CrossFusionBoxWrapper

    w = WORKLOADS[name]
    kw = {} if dropout else dict(patch_dropout=0.0, token_dropout=0.0, backproj_dropout=0.0)
    cfg = default_fusion_cfg(w["token_dim"], n_levels=len(w["channels"]), num_layers=w["num_layers"],
                             num_heads=w["num_heads"], patch=w["patch"], **kw)
    torch.manual_seed(seed)
    rcnn = FeatureProviderRCNN(level_shapes(w), w["channels"], w["noun_classes"], w["verb_classes"])
    m = CrossFusionBoxWrapper(rcnn, copy.deepcopy(cfg), {"text_pooling": "synthetic", "train_ep": -1},
                              criterion={"lm": 1 if lm else 0}, narr_pooling_layer=PassThroughPooling())
    return m.to(device)


def synthetic_inputs(name: str, batch: int, lang_len: int, seed: int, device="cpu", feat_dtype=torch.float32,
                     pin: bool = False):
    """features relu(randn) (rms ~0.7), language ~ N(0, 0.25), valid lengths ~U{L/2..L}
    with at least one full-length sample."""
    w = WORKLOADS[name]
    g = torch.Generator().manual_seed(seed)
    feats = {}
    for i, ((h, ww), c) in enumerate(zip(level_shapes(w), w["channels"])):
        t = torch.relu(torch.randn(batch, c, h, ww, generator=g)).to(feat_dtype)
        feats[str(i)] = t.pin_memory() if pin else t
    lang = 0.5 * torch.randn(batch, lang_len, w["token_dim"], generator=g)
    lens = torch.randint(lang_len // 2, lang_len + 1, (batch,), generator=g)
    lens[0] = lang_len
    mask = (torch.arange(lang_len)[None, :] < lens[:, None]).to(torch.int64)
    if pin:
        lang, mask = lang.pin_memory(), mask.pin_memory()
    if device != "cpu":
        feats = {k: v.to(device) for k, v in feats.items()}
        lang, mask = lang.to(device), mask.to(device)
    return feats, lang, mask
