"""Configuration of the fusion path as the reference ships it.

``default_fusion_cfg`` restates the live keys of
modeling/cross_fusion/ego_fusion/cross_fusion_config_sym_ego_res50.yml:1-41,85-91 plus what
runner/run_experiment.update_config injects (:75-77,100) and the run YAML adds
(runner/nao/configs/ego_nao_res50_ego4dv2.yml:86-93).  Shapes of the two shipped workloads are in
``WORKLOADS``."""
from __future__ import annotations

import copy


def default_fusion_cfg(token_dim: int, n_levels: int = 4, num_layers=None, num_heads: int = 4, patch=None,
                       patch_dropout: float = 0.1, token_dropout: float = 0.15, backproj_dropout: float = 0.1,
                       use_lm_f: bool = True, forward_language_f=False) -> dict:
    patch = list(patch) if patch is not None else [4, 4, 2, 1][:n_levels]
    cfg = {
        "model": "cross_f",
        "type": "cross_transformer",
        "share_encoders": False,
        "narr_out_mode": "tokens",
        "patch_h": list(patch),
        "patch_w": list(patch),
        "backproj_dropout": backproj_dropout,
        "backproj_activ_f": None,
        "patch_norm": {"visual": None, "language": None},
        "pos_embedding": "sin1d",
        "forward_language_f": forward_language_f,
        "vis_mask_type": "global",
        "args": {
            "patch_dropout": patch_dropout,
            "num_layers": list(num_layers) if num_layers is not None else [4] * n_levels,
            "num_heads": num_heads,
            "fforward_multiplier": 2,
            "token_dropout": token_dropout,
            "back_to_img_fn": "regroup",
            "activ_f": "gelu",
            "final_norm": "ln",
            "input_f_size": token_dim,
        },
        "lm_args": {"pooling": {"type": "mean", "ln": True, "repr_size": 0}, "multi": False, "use_lm_f": bool(use_lm_f)},
        "fpn_features": list(range(n_levels)),
        "replace_fpn_features": True,
    }
    return copy.deepcopy(cfg)


# the two shipped shapes (padded batch image, C2..C5 strides 4/8/16/32)
WORKLOADS = {
    "ego4dv2": dict(token_dim=896, image=(768, 1024), channels=[256, 512, 1024, 2048], strides=[4, 8, 16, 32],
                    patch=[4, 4, 2, 1], num_layers=[4, 4, 4, 4], num_heads=4, train_batch=13, eval_batch=74, lang_len=64,
                    noun_classes=129, verb_classes=82),
    "ego4dv1": dict(token_dim=712, image=(800, 1280), channels=[256, 512, 1024, 2048], strides=[4, 8, 16, 32],
                    patch=[4, 4, 2, 1], num_layers=[4, 4, 4, 4], num_heads=4, train_batch=18, eval_batch=36, lang_len=64,
                    noun_classes=88, verb_classes=75),
}


def level_shapes(workload: dict):
    H, W = workload["image"]
    return [(H // s, W // s) for s in workload["strides"]]
