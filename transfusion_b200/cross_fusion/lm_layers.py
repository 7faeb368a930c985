"""Mirror of modeling/cross_fusion/ego_fusion/lm_layers.py (PoolPredictor: mask-multiply, mean / max pooling,
LayerNorm, optional GELU + Linear, noun / verb Linears).  The nn.LayerNorm / nn.Linear submodules are parameter
containers with the reference's names and initialisation; the math runs in the fp32 LM-head kernels of
csrc/lm_head.cu through the C ABI (xf_lm_pool_*, xf_rowln_*, xf_small_linear_*), forward and backward.
There is no PyTorch fallback: CPU tensors raise."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class _PoolFn(torch.autograd.Function):
    """lm_layers.py:60-66."""

    @staticmethod
    def forward(ctx, tok, mask_u8, kind):
        pooled, argmax = ops.lm_pool_fwd(tok, mask_u8, kind)
        ctx.kind, ctx.L = kind, tok.shape[1]
        ctx.save_for_backward(*(t for t in (mask_u8, argmax) if t is not None))
        ctx.has = (mask_u8 is not None, argmax is not None)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        saved = list(ctx.saved_tensors)
        mask = saved.pop(0) if ctx.has[0] else None
        argmax = saved.pop(0) if ctx.has[1] else None
        return ops.lm_pool_bwd(dpooled, mask, argmax, ctx.L, ctx.kind), None, None


class _RowLnFn(torch.autograd.Function):
    """nn.LayerNorm on [B, D] rows (lm_layers.py:68-69)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        y, mean, rstd = ops.rowln_fwd(x, gamma, beta, eps)
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dx, dg, db = ops.rowln_bwd(dy, x, gamma, mean, rstd)
        return dx, dg, db, None


class _LinearFn(torch.autograd.Function):
    """y = act(x) W^T + b with act = identity or GELU(erf) on the input (lm_layers.py:43-55)."""

    @staticmethod
    def forward(ctx, x, W, bias, act):
        ctx.act, ctx.has_bias = act, bias is not None
        ctx.save_for_backward(x, W)
        return ops.small_linear_fwd(x, W, bias, act)

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dx, dW, db = ops.small_linear_bwd(dy, x, W, ctx.act, need_dx=ctx.needs_input_grad[0], need_dw=ctx.needs_input_grad[1],
                                          has_bias=ctx.has_bias)
        return dx, dW, (db if ctx.has_bias else None), None


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
        kind = self.pooling_args["type"]
        if kind not in ("max", "mean"):
            raise NotImplementedError
        if not fused_l_tokens.is_cuda:
            raise ops._lib.XfError("PoolPredictor: expected CUDA tensors (there is no CPU path)")
        mask_u8 = None if att_mask is None else att_mask.to(torch.uint8).contiguous()
        features = _PoolFn.apply(fused_l_tokens.float(), mask_u8, kind)   # mean divides by the padded length (:65-66)
        if self.ln:
            features = _RowLnFn.apply(features, self.ln.weight, self.ln.bias, self.ln.eps)
        if self.repr_mlp:
            lin = self.repr_mlp[1]
            features = _LinearFn.apply(features, lin.weight, lin.bias, 1)
        noun_logits = _LinearFn.apply(features, self.mlp_noun.weight, self.mlp_noun.bias, 0)
        verb_logits = _LinearFn.apply(features, self.mlp_verb.weight, self.mlp_verb.bias, 0) if self.mlp_verb else None
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
