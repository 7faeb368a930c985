"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32/fp64, explicit math) of the
reference cross_fusion hot path.  Not product code; see oracle/__init__.py.

Every function cites the reference file:line it follows (paths relative to the
reference root).  Pinned against the unmodified reference module by
``tests/test_oracle_vs_reference.py`` (build container) and against the committed
golden vectors ``tests/golden/*.npz`` (everywhere).

The functions take a flat ``sd`` mapping with the reference's ``state_dict`` key names
 so the same weights drive the reference, the oracle and the
CUDA module.  Dropout is deterministic: every dropout site takes an explicit keep mask
(``masks`` dicts, 1 = keep) and applies ``x * mask / (1 - p)`` exactly like ``F.dropout``; without
masks the probabilities are 0.  Site names per level: ``patch`` [B,n,D]
(cross_f_box_layers.py:74), ``l{t}.attn`` [B,H,S,S] (torch18_adapters.py:796-797), ``l{t}.drop1``
[B,S,D] (:109), ``l{t}.ffn`` [B,S,2D] (:111), ``l{t}.drop2`` [B,S,D] (:112), ``backproj`` [B,n,D]
(utils.py:115).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

LN_EPS = 1e-5  # nn.LayerNorm default; torch18_adapters.py:63,79-80


def sin1d_table(n: int, dim: int, dtype=torch.float32) -> torch.Tensor:
    """modeling/cross_fusion/utils.py:267-273 (get_sin1d_embed), rows 0..n-1 -> [n, dim]."""
    position = torch.arange(n).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2) * (-math.log(10000.0) / dim))
    pe = torch.zeros(n, dim)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(dtype)


def patchify(feat: torch.Tensor, p: int) -> torch.Tensor:
    """[B,C,h,w] -> [B, n, C*p*p] with column order (c,u,v) and row-major token order
    over (gh,gw) — the im2col of ``nn.Conv2d(k=stride=p)`` (cross_f_box_wrapper.py:268-274)
    followed by ``patchify_image(.,1,1)`` (utils.py:35-39)."""
    B, C, h, w = feat.shape
    gh, gw = h // p, w // p
    x = feat.reshape(B, C, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B, gh * gw, C * p * p)


def fold(y: torch.Tensor, C: int, p: int, gh: int, gw: int) -> torch.Tensor:
    """utils.py:42-46 (regroup_patches: transpose + F.fold, kernel = stride = p):
    out[b,c,i*p+u,j*p+v] = y[b, i*gw+j, c*p*p+u*p+v]."""
    B = y.shape[0]
    x = y.reshape(B, gh, gw, C, p, p).permute(0, 3, 1, 4, 2, 5)
    return x.reshape(B, C, gh * p, gw * p)


def layer_norm(x, w, b):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS) * w + b


def gelu_erf(x):
    """F.gelu exact erf form (torch18_adapters.py:10 _get_activation_fn('gelu'))."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _drop(x, mask, p: float):
    """F.dropout with an explicit keep mask (1 = keep): x * mask / (1 - p)."""
    if mask is None or p <= 0.0:
        return x
    return x * mask.to(x.dtype) / (1.0 - p)


def attention(q, k, v, key_pad: Optional[torch.Tensor], num_heads: int, drop_mask=None, drop_p: float = 0.0):
    """torch18_adapters.py:544-555 (head split), :578-597 (key-padding -> -inf),
    :789-798 (_scaled_dot_product_attention), :607 (head merge).
    q:[B,Sq,D] k,v:[B,Sk,D]; key_pad:[B,Sk] bool, True = ignore.  General Sq != Sk."""
    B, Sq, D = q.shape
    Sk = k.shape[1]
    d = D // num_heads
    qh = q.reshape(B, Sq, num_heads, d).transpose(1, 2)
    kh = k.reshape(B, Sk, num_heads, d).transpose(1, 2)
    vh = v.reshape(B, Sk, num_heads, d).transpose(1, 2)
    s = (qh / math.sqrt(d)) @ kh.transpose(-1, -2)
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    a = _drop(a, drop_mask, drop_p)  # torch18_adapters.py:796-797 (dropout on the probabilities)
    o = a @ vh
    return o.transpose(1, 2).reshape(B, Sq, D)


def encoder_layer(x, key_pad, sd, pre: str, num_heads: int, masks=None, tag: str = "", drop_p: float = 0.0):
    """Post-LN encoder layer, torch18_adapters.py:108-113; in-proj :685; out-proj :608.
    masks: optional dict with keep masks ``{tag}.attn / .drop1 / .ffn / .drop2`` (token_dropout sites)."""
    D = x.shape[-1]
    mk = (lambda n: masks.get(f"{tag}.{n}")) if masks else (lambda n: None)
    qkv = x @ sd[pre + "self_attn.in_proj_weight"].t() + sd[pre + "self_attn.in_proj_bias"]
    q, k, v = qkv.split(D, dim=-1)
    o = attention(q, k, v, key_pad, num_heads, mk("attn"), drop_p)
    o = o @ sd[pre + "self_attn.out_proj.weight"].t() + sd[pre + "self_attn.out_proj.bias"]
    x = layer_norm(x + _drop(o, mk("drop1"), drop_p), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"])
    f = gelu_erf(x @ sd[pre + "linear1.weight"].t() + sd[pre + "linear1.bias"])
    f = _drop(f, mk("ffn"), drop_p) @ sd[pre + "linear2.weight"].t() + sd[pre + "linear2.bias"]
    return layer_norm(x + _drop(f, mk("drop2"), drop_p), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"])


def fusion_level(feat, lang, lang_pad, sd, i: int, p: int, num_heads: int, num_layers: int, masks=None,
                 drop=(0.0, 0.0, 0.0)):
    """One FPN level: cross_f_box_wrapper.py:177-212 + cross_f_box_layers.py:69-108.
    Returns (fused feature map [B,C,h,w], language tokens out [B,L,D]).
    masks: optional keep masks of this level (module docstring); drop = (patch, token, backproj) probabilities."""
    p_patch, p_tok, p_back = drop
    B, C, h, w = feat.shape
    gh, gw = h // p, w // p
    n = gh * gw
    D = lang.shape[-1]
    enc = f"cross_fusion_encoders.{i}."
    # cross_f_box_wrapper.py:183,185 — patch-embed conv (no bias) + token order
    wpe = sd[f"patches_to_token.{i}.weight"].reshape(D, C * p * p)
    x = patchify(feat, p) @ wpe.t()
    # cross_f_box_layers.py:72-73 ; utils.py:209-214
    x = x + sin1d_table(n, D, x.dtype) + sd[enc + "image_kind_embedding"].reshape(D)
    x = _drop(x, masks.get("patch") if masks else None, p_patch)  # cross_f_box_layers.py:74
    # cross_f_box_layers.py:76
    lg = lang + sd[enc + "lang_kind_embedding"].reshape(D)
    # cross_f_box_layers.py:80-86
    key_pad = None
    if lang_pad is not None:
        key_pad = torch.cat([torch.zeros(B, n, dtype=torch.bool), lang_pad], dim=1)
    z = torch.cat([x, lg], dim=1)
    for l in range(num_layers):  # cross_f_box_layers.py:97
        z = encoder_layer(z, key_pad, sd, enc + f"t_encoder.layers.{l}.", num_heads, masks, f"l{l}", p_tok)
    # cross_f_box_layers.py:104-108
    vis = layer_norm(z[:, :n], sd[enc + "final_norm_layer.weight"], sd[enc + "final_norm_layer.bias"])
    lang_out = z[:, n:]
    # utils.py:114-119
    vis = _drop(vis, masks.get("backproj") if masks else None, p_back)  # utils.py:115
    y = vis @ sd[f"tokens_to_features.{i}.linear.weight"].t() + sd[f"tokens_to_features.{i}.linear.bias"]
    return fold(y, C, p, gh, gw), lang_out


def lm_head(lang, att_mask_bool, sd):
    """lm_layers.py:59-81 (PoolPredictor, mean pooling over the padded length, LN)."""
    t = lang * att_mask_bool.unsqueeze(2).to(lang.dtype)
    f = t.mean(dim=1)
    f = layer_norm(f, sd["lm_layer.ln.weight"], sd["lm_layer.ln.bias"])
    noun = f @ sd["lm_layer.mlp_noun.weight"].t() + sd["lm_layer.mlp_noun.bias"]
    verb = f @ sd["lm_layer.mlp_verb.weight"].t() + sd["lm_layer.mlp_verb.bias"]
    return {"noun_logits": noun, "verb_logits": verb}


def cross_fusion_forward(features: Dict[str, torch.Tensor], lang, att_mask, sd,
                         patch: Sequence[int], num_heads: int, num_layers: Sequence[int],
                         lm: bool = False, use_lm_f: bool = True, forward_language_f=False,
                         masks: Optional[Dict[str, Dict[str, torch.Tensor]]] = None, drop=(0.0, 0.0, 0.0)):
    """cross_f_box_wrapper.py:165-230 with identity FPN/RoI (fused maps exposed).
    att_mask: [B,L] int, 1 = valid (narr_pooling_layers.py:199-202); inverted at
    cross_f_box_wrapper.py:196.  ``forward_language_f`` in (False, "direct", "sum") chains the fused
    language tokens into the next level (:203-209); ``use_lm_f`` False feeds the LM head the LAST
    level's fused tokens instead of the input language features (:224-227).
    masks[level key][site] are optional dropout keep masks (module docstring)."""
    lang_pad = ~(att_mask.bool())
    out = {}
    fused_l = None
    for i, key in enumerate(sorted(features.keys(), key=int)):
        fused, fused_l = fusion_level(features[key], lang, lang_pad, sd, i, patch[i], num_heads, num_layers[i],
                                      masks.get(key) if masks else None, drop)
        if forward_language_f == "direct":
            lang = fused_l
        elif forward_language_f == "sum":
            lang = lang + fused_l  # the reference adds in place (:206); same value
        elif forward_language_f:
            raise NotImplementedError(forward_language_f)
        out[key] = fused
    lm_out = lm_head(lang if use_lm_f else fused_l, att_mask.bool(), sd) if lm else None
    return out, lm_out


def algorithmic_flops_fwd(shapes, channels, patch, D, L, num_layers) -> float:
    """F_fwd = sum_l [4 n K D + layers*(16 S D^2 + 4 S^2 D)] per sample."""
    total = 0.0
    for (h, w), C, p, nl in zip(shapes, channels, patch, num_layers):
        n = (h // p) * (w // p)
        K = C * p * p
        S = n + L
        total += 4.0 * n * K * D + nl * (16.0 * S * D * D + 4.0 * S * S * D)
    return total
