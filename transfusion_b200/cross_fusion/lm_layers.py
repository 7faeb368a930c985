"""Mirror of modeling/cross_fusion/ego_fusion/lm_layers.py (PoolPredictor, mean pooling + LN).
The head is [B, D] -> a few hundred logits: negligible work, kept as the stock torch modules the
reference uses (it is off in the shipped configs: criterion.lm = 0)."""
from __future__ import annotations

import torch
from torch import nn


class PoolPredictor(nn.Module):
    """lm_layers.py:30-81."""

    def __init__(self, pooling_args, token_dim, no_nouns, no_verbs):
        super().__init__()
        self.pooling_args = pooling_args
        self.token_dim = token_dim
        self.repr_size = token_dim
        self.ln = None
        self.repr_mlp = None
        self.mlp_verb = None
        if pooling_args.get("ln", None):
            self.ln = nn.LayerNorm(token_dim)
        if pooling_args.get("repr_size", None):
            self.repr_mlp = nn.Sequential(nn.GELU(), nn.Linear(token_dim, pooling_args["repr_size"]))
            self.repr_size = pooling_args["repr_size"]
        self.mlp_noun = nn.Linear(self.repr_size, no_nouns)
        if no_verbs:
            self.mlp_verb = nn.Linear(self.repr_size, no_verbs)

    def forward(self, fused_l_tokens, att_mask=None):
        if att_mask is not None:
            fused_l_tokens = fused_l_tokens * att_mask.unsqueeze(2)
        if self.pooling_args["type"] == "max":
            features = fused_l_tokens.max(dim=1)[0]
        elif self.pooling_args["type"] == "mean":
            features = fused_l_tokens.mean(dim=1)  # divides by the padded length (lm_layers.py:65-66)
        else:
            raise NotImplementedError
        if self.ln:
            features = self.ln(features)
        if self.repr_mlp:
            features = self.repr_mlp(features)
        noun_logits = self.mlp_noun(features)
        verb_logits = self.mlp_verb(features) if self.mlp_verb else None
        return {"noun_logits": noun_logits, "verb_logits": verb_logits}


def get_lm_layer(wrapper):
    """lm_layers.py:5-27 (single-scale PoolPredictor only; multi-scale variants are off in the
    shipped config: lm_args.multi = False)."""
    args = wrapper.cross_encoder_args
    no_nouns = wrapper.rcnn_model.noun_classes - 1
    no_verbs = wrapper.rcnn_model.verb_classes - 1
    if args["lm_args"]["pooling"]["type"] in {"mean", "max"} and not args["lm_args"].get("multi", False):
        return PoolPredictor(args["lm_args"]["pooling"], wrapper.token_dim, no_nouns, no_verbs)
    raise NotImplementedError("only the single-scale mean/max PoolPredictor is supported")
