"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference hot path in-process.

Works where ``/root/reference`` exists (the build container) or where its byte-identical staged copy
``oracle/_ref`` (oracle/build_ref.py; git-ignored, travels with the gpurun snapshot) does.  It is used to
 (1) pin ``oracle/ref_math.py`` (tests/test_oracle_vs_reference.py, skipped elsewhere) and
 (2) generate the committed golden fixtures (oracle/make_golden.py).

The reference's ``cross_f_box_wrapper`` transitively imports two modules whose real
versions need detectron2 / natsort / sentence_transformers (absent here and unused by the
live math).  They are pre-seeded in ``sys.modules`` with minimal stand-ins:

  * ``modeling.commons``                       (real: modeling/commons.py:33,44)
  * ``modeling.narration_embeds.narr_pooling_layers`` (real: narr_pooling_layers.py:23-33)

No reference source is copied; the reference classes run as they are.
"""
from __future__ import annotations

import copy
import os
import sys
import types

import torch
from torch import nn

FUSION_YAML = "modeling/cross_fusion/ego_fusion/cross_fusion_config_sym_ego_res50.yml"
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/build_ref.py (unmodified copies)


def _pick_root() -> str:
    """The reference tree itself where it exists (build container), else the byte-identical staged copy oracle/_ref
    (git-ignored; travels to the GPU box), else the path that will fail ``reference_available()``."""
    env = os.environ.get("XF_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(cand, FUSION_YAML)):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, FUSION_YAML))


def _seed_stub_modules() -> None:
    if "modeling.commons" not in sys.modules:
        commons = types.ModuleType("modeling.commons")

        class NaoABC(nn.Module):  # stand-in for modeling/commons.py:44
            pass

        commons.NaoABC = NaoABC
        commons.freeze_all_but_bn = lambda m: None  # modeling/commons.py:33
        sys.modules["modeling.commons"] = commons
    name = "modeling.narration_embeds.narr_pooling_layers"
    if name not in sys.modules:
        npl = types.ModuleType(name)
        npl.get_narr_pooling_layer = lambda typ: (lambda narr_args, out_mode: nn.Identity())
        sys.modules[name] = npl


def import_reference_wrapper():
    """Returns the reference ``CrossFusionBoxWrapper`` class (cross_f_box_wrapper.py:41)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _seed_stub_modules()
    from modeling.cross_fusion.ego_fusion.cross_f_box_wrapper import CrossFusionBoxWrapper

    return CrossFusionBoxWrapper


def load_fusion_yaml() -> dict:
    import yaml

    with open(os.path.join(REFERENCE_ROOT, FUSION_YAML)) as f:
        return yaml.safe_load(f)


class FakeRCNN(nn.Module):
    """The five things ``CrossFusionBoxWrapper`` needs from ``rcnn_model``.

    ``forward_features`` hands back pre-made feature maps; FPN / RPN / RoI are identity so
    the fused feature maps are exposed directly.
    """

    def __init__(self, shapes, channels, noun_classes=129, verb_classes=82):
        super().__init__()
        self._shapes = list(shapes)
        self._channels = list(channels)
        self.noun_classes = noun_classes
        self.verb_classes = verb_classes
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
    """Emulates SBertLayer's return contract (narr_pooling_layers.py:199-202)."""

    def forward(self, lang, pad_mask=False):
        emb, mask = lang
        return emb, None, (mask if pad_mask else None)


def build_fusion_cfg(token_dim, n_levels=4, num_layers=None, num_heads=4, patch=None,
                     dropout=0.0, base_cfg=None, use_lm_f=None, forward_language_f=None) -> dict:
    """Fusion config as ``run_experiment.update_config`` would hand it over
    (run_experiment.py:75-77,100), with dropout probabilities overridable (0.0 = all off, anything else =
    the yml's values) and the two language-routing switches (lm_args.use_lm_f, forward_language_f)."""
    cfg = copy.deepcopy(base_cfg) if base_cfg is not None else load_fusion_yaml()
    patch = list(patch) if patch is not None else cfg["patch_h"][:n_levels]
    cfg["patch_h"] = list(patch)
    cfg["patch_w"] = list(patch)
    cfg["fpn_features"] = list(range(n_levels))
    cfg["replace_fpn_features"] = True
    cfg["args"]["input_f_size"] = token_dim
    cfg["args"]["num_heads"] = num_heads
    cfg["args"]["num_layers"] = list(num_layers) if num_layers is not None else [4] * n_levels
    cfg["args"]["patch_dropout"] = dropout if dropout == 0.0 else cfg["args"]["patch_dropout"]
    cfg["args"]["token_dropout"] = dropout if dropout == 0.0 else cfg["args"]["token_dropout"]
    cfg["backproj_dropout"] = dropout if dropout == 0.0 else cfg["backproj_dropout"]
    if use_lm_f is not None:
        cfg["lm_args"]["use_lm_f"] = bool(use_lm_f)
    if forward_language_f is not None:
        cfg["forward_language_f"] = forward_language_f
    return cfg


def build_reference_module(cfg, shapes, channels, lm=False, seed=0,
                           noun_classes=129, verb_classes=82):
    """Constructs the reference wrapper under ``torch.manual_seed(seed)``."""
    Wrapper = import_reference_wrapper()
    torch.manual_seed(seed)
    rcnn = FakeRCNN(shapes, channels, noun_classes, verb_classes)
    m = Wrapper(rcnn, copy.deepcopy(cfg), {"text_pooling": "x", "train_ep": -1},
                criterion={"lm": 1 if lm else 0})
    m.narr_pooling_layer = PassThroughPooling()
    return m


def run_reference(m, features: dict, lang: torch.Tensor, att_mask: torch.Tensor):
    """Forward of the unmodified reference.  Returns (fused feature dict, lm dict or None)."""
    m.rcnn_model.features = features
    out = m({"image": None, "language_f": (lang, att_mask)})
    return out["features"], out.get("lm")


class recorded_dropout:
    """Context manager that makes the UNMODIFIED reference's dropout deterministic and observable: every
    ``F.dropout`` call (nn.Dropout modules, cross_f_box_layers.py:74, the attention-probability dropout) draws
    its keep mask from a seeded CPU generator and appends it to ``self.masks`` in call order.  torch >= 2 routes
    ``nn.MultiheadAttention`` through the fused ``scaled_dot_product_attention`` builtin whose dropout cannot be
    observed, so that one function is replaced by its documented math (softmax(q k^T * scale + mask) -> dropout
    -> @ v), which is the reference's own torch18_adapters.py:789-798."""

    def __init__(self, seed: int):
        self.gen = torch.Generator().manual_seed(seed)
        self.masks = []

    def _dropout(self, x, p=0.5, training=True, inplace=False):
        if not training or p <= 0.0:
            return x
        keep = (torch.rand(x.shape, generator=self.gen) >= p)
        self.masks.append(keep)
        return x * keep.to(x.dtype) / (1.0 - p)

    def _sdpa(self, q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, **kw):
        import math
        assert not is_causal
        s = (q * (scale if scale is not None else 1.0 / math.sqrt(q.shape[-1]))) @ k.transpose(-2, -1)
        if attn_mask is not None:
            s = s.masked_fill(attn_mask, float("-inf")) if attn_mask.dtype == torch.bool else s + attn_mask
        a = torch.softmax(s, dim=-1)
        a = self._dropout(a, dropout_p, True)
        return a @ v

    def __enter__(self):
        import torch.nn.functional as F
        self._saved = (F.dropout, F.scaled_dot_product_attention)
        F.dropout = self._dropout
        F.scaled_dot_product_attention = self._sdpa
        return self

    def __exit__(self, *exc):
        import torch.nn.functional as F
        F.dropout, F.scaled_dot_product_attention = self._saved
        return False


def masks_by_site(recorded, n_levels: int, num_layers, level_order=None):
    """Maps the call-ordered mask list of ``recorded_dropout`` to oracle/ref_math.py's site names: per level
    (reference order 0..n-1): patch, then per layer attn, drop1, ffn, drop2, then backproj."""
    it = iter(recorded)
    out = {}
    for i in (level_order if level_order is not None else range(n_levels)):
        d = {"patch": next(it)}
        for l in range(num_layers[i]):
            for site in ("attn", "drop1", "ffn", "drop2"):
                d[f"l{l}.{site}"] = next(it)
        d["backproj"] = next(it)
        out[str(i)] = d
    rest = list(it)
    assert not rest, f"{len(rest)} unconsumed dropout masks"
    return out
