"""Mirror of modeling/cross_fusion/utils.py for the live box-model path: parameter containers and
constants only — the arithmetic runs in the CUDA library (see level_fn.py)."""
from __future__ import annotations

import math

import torch
from torch import nn


def get_visual_token_mask(img_shape, mask_type):
    """utils.py:9-32.  Only the shipped ``global`` setting (no visual-token mask) is implemented."""
    if mask_type == "global":
        return None
    raise NotImplementedError(f"vis_mask_type={mask_type!r}: only 'global' (the shipped config) is supported")


def get_sin1d_embed(no_embeds: int, dim: int) -> torch.Tensor:
    """utils.py:267-273 — the sin1d table [1, no_embeds, dim] (fp32)."""
    position = torch.arange(no_embeds).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2) * (-math.log(10000.0) / dim))
    pe = torch.zeros(no_embeds, dim)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


class PositionalEmbeddingLayer(nn.Module):
    """utils.py:172-218.  Holds the ``pos_embedding`` buffer / parameter under the reference's
    state_dict key; the add itself is fused into the patch-embed GEMM epilogue."""

    def __init__(self, embedding_type, num_patches, token_dim, temporal_dim=0):
        super().__init__()
        if temporal_dim:
            raise NotImplementedError("temporal positional embeddings are not part of the box-model path")
        self.embedding_type = embedding_type
        self.num_patches = num_patches
        self.token_dim = token_dim
        self.temporal_dim = temporal_dim
        if embedding_type == "learned":
            self.pos_embedding = nn.Parameter(torch.randn(1, num_patches, token_dim))
        elif embedding_type == "zero":
            self.pos_embedding = nn.Parameter(torch.zeros(1, num_patches, token_dim))
        elif embedding_type == "sin1d":
            self.register_buffer("pos_embedding", get_sin1d_embed(num_patches, token_dim))
        else:
            raise ValueError(f"{embedding_type=} is not recognized")

    def table(self) -> torch.Tensor:
        return self.pos_embedding[0]


class RegroupPatchesLayerBox(nn.Module):
    """utils.py:84-119.  Container for the back-projection ``linear`` (+ ``init_h/init_w`` the wrapper
    sets per call, cross_f_box_wrapper.py:180-181)."""

    def __init__(self, token_dim, init_h, init_w, patch_h, patch_w, out_channels, backproj_dropout=0.1,
                 activ_f=None, final_norm=False):
        super().__init__()
        if activ_f is not None:
            raise NotImplementedError("backproj_activ_f other than null is not part of the shipped config")
        if final_norm:
            raise NotImplementedError("RegroupPatchesLayerBox.final_norm is not part of the shipped config")
        if patch_h != patch_w:
            raise NotImplementedError("non-square patches")
        self.init_h = init_h
        self.init_w = init_w
        self.patch_h = patch_h
        self.patch_w = patch_w
        self.out_channels = out_channels
        self.back_dropout = nn.Dropout(backproj_dropout)
        self.linear = nn.Linear(token_dim, patch_h * patch_w * out_channels)
