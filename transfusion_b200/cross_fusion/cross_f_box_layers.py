"""Mirror of modeling/cross_fusion/ego_fusion/cross_f_box_layers.py:13-108 (CrossTransformerModuleBox).

Same constructor, forward signature, return tuple and state_dict keys.  The submodules are the same
torch classes the reference instantiates, used purely as *parameter containers* (identical
initialisation under a given seed, identical key names); their ``forward`` is never called — the
arithmetic is FusionLevelFunction over the CUDA library."""
from __future__ import annotations

import torch
from torch import nn


class CrossTransformerModuleBox(nn.Module):
    def __init__(self, no_patches, patch_dropout, input_f_size, pos_embedding_layer, num_layers=2, num_heads=4,
                 classif_token=False, fforward_multiplier=2, token_dropout=0.1, back_to_img_fn="token",
                 activ_f="relu", patch_norm=False, final_norm=False, lang_pos_embedding=None):
        super().__init__()
        if classif_token:
            raise NotImplementedError("classif_token is dead code in the reference box model (SURVEY Appendix D)")
        if activ_f != "gelu":
            raise NotImplementedError(f"activ_f={activ_f!r}: the CUDA path implements the shipped 'gelu' (erf) activation")
        if lang_pos_embedding:
            raise NotImplementedError("lang_pos_embedding is absent from the shipped config")
        if final_norm != "ln":
            raise NotImplementedError("final_norm must be 'ln' (shipped config)")
        if input_f_size % num_heads != 0 or input_f_size % 8 != 0:
            raise ValueError("input_f_size must be a multiple of num_heads and of 8")
        self.no_patches = no_patches
        self.classif_token = classif_token
        self.back_to_img_fn = back_to_img_fn
        self.patch_norm = patch_norm
        self.final_norm = final_norm
        self.token_dim = input_f_size
        self.num_heads = num_heads
        self.num_layers = num_layers
        self.token_dropout = token_dropout
        self.pos_embedding_layer = pos_embedding_layer
        self.image_kind_embedding = nn.Parameter(torch.randn(1, 1, self.token_dim))
        self.lang_kind_embedding = nn.Parameter(torch.randn(1, 1, self.token_dim))
        self.lang_pos_embedding = lang_pos_embedding
        self.heatmap_token = nn.Parameter(torch.randn(1, 1, self.token_dim))  # registered, unused (reference :43)
        self.patch_dropout = patch_dropout
        self.register_buffer("padding_mask", torch.zeros(size=(1,), dtype=torch.bool))
        t_encoder_layer = nn.TransformerEncoderLayer(
            d_model=self.token_dim, nhead=num_heads, dim_feedforward=int(self.token_dim * fforward_multiplier),
            batch_first=True, dropout=token_dropout, activation=activ_f)
        self.t_encoder = nn.TransformerEncoder(t_encoder_layer, num_layers, enable_nested_tensor=False)
        self.final_norm_layer = nn.LayerNorm(self.token_dim)

    def level_params(self):
        """Parameter tensors in the order FusionLevelFunction expects (encoder part)."""
        out = []
        for layer in self.t_encoder.layers:
            out += [layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias, layer.self_attn.out_proj.weight,
                    layer.self_attn.out_proj.bias, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight,
                    layer.linear2.bias, layer.norm1.weight, layer.norm1.bias, layer.norm2.weight, layer.norm2.bias]
        return out

    def forward(self, x, language_tokens, language_tokens_att_maks, vis_tokens_mask=None):
        raise RuntimeError(
            "CrossTransformerModuleBox is executed as part of the fused level schedule "
            "(CrossFusionBoxWrapper.forward -> FusionLevelFunction); call the wrapper, or "
            "transfusion_b200.cross_fusion.cross_f_box_wrapper.run_level for a single level")
