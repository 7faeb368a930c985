"""SURVEY 8f N2 — the language-context producer of the fusion path on the path's kernels.

The reference's ``SBertLayer`` (modeling/narration_embeds/narr_pooling_layers.py:75-202) tokenises the narration strings on
the CPU, runs sentence-transformers' MiniLM-L12-H384 (a HuggingFace ``BertModel``: 12 post-LN layers, hidden 384, 12 heads
of 32, FFN 1536, GELU(erf), LayerNorm eps 1e-12) to get ``token_embeddings`` [B, L, 384] (:163), and maps them to the
fusion width with ``out_mlp = Linear(384, D)`` (:95-96,189-190); the encoder is frozen in both shipped configs
(``train_ep: -1``, ego_vis_det_ego4dv2.yml:3; ``freeze_all_but_bn`` :87), ``out_mlp`` trains.

A BERT layer is the same post-LN block as the fusion encoder layer (torch18_adapters.py:108-113), so the encoder forward
runs on exactly the fusion path's kernels: ``xf_gemm`` (fused QKV projection, out-proj + residual, FFN1 + GELU, FFN2 +
residual), ``xf_attn_fwd`` (key-padding mask from ``attention_mask``, head_dim 32) and ``xf_layernorm_fwd``.  The HF module
stays the parameter container (checkpoints load as before); tokenisation and the embedding-table gathers stay in torch.
Inference forward, and -- in the state the reference trains the encoder in (``freeze_all_but_bn``: matrices frozen,
LayerNorms trainable, dropouts on) -- forward + backward (``_BertEncoderFn``: LayerNorm gradients through all 12 layers on
the same dgrad GEMM / attention-backward / LayerNorm-backward kernels); other trainable encoder weights raise.
``out_mlp`` has forward and backward (``XfLinear``).  No CPU fallback."""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from .. import ops
from ..weight_cache import bf16_weight


def _layer_weights(layer, D: int, dev):
    """bf16 copies of one BertLayer's matrices, QKV concatenated to [3D, D]; cached on the layer while the parameters'
    version counters are unchanged (the encoder is frozen)."""
    att, inter, out = layer.attention, layer.intermediate, layer.output
    prms = [att.self.query.weight, att.self.key.weight, att.self.value.weight, att.output.dense.weight, inter.dense.weight,
            out.dense.weight, att.self.query.bias, att.self.key.bias, att.self.value.bias]
    key = tuple(p._version for p in prms) + (str(dev),)
    cache = layer.__dict__.get("_xf_w")
    if cache is not None and cache[0] == key:
        return cache[1]
    bf = torch.bfloat16
    F = inter.dense.weight.shape[0]
    wqkv = torch.empty(3 * D, D, device=dev, dtype=bf)
    wo = torch.empty(D, D, device=dev, dtype=bf)
    w1 = torch.empty(F, D, device=dev, dtype=bf)
    w2 = torch.empty(D, F, device=dev, dtype=bf)
    casts = [(prms[i].detach(), wqkv[i * D:(i + 1) * D], D, D, 0, 0, 0, 0) for i in range(3)]
    casts += [(prms[3].detach(), wo, D, D, 0, 0, 0, 0), (prms[4].detach(), w1, F, D, 0, 0, 0, 0), (prms[5].detach(), w2, D, F, 0, 0, 0, 0)]
    ops.cast_pad_multi(casts)
    bqkv = torch.cat([prms[6].detach(), prms[7].detach(), prms[8].detach()]).float().contiguous()
    val = (wqkv, bqkv, wo, w1, w2)
    layer.__dict__["_xf_w"] = (key, val)
    return val


def bert_encoder_forward(bert, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                         token_type_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``bert(input_ids, attention_mask, token_type_ids).last_hidden_state`` (fp32 [B, L, hidden]) for a HuggingFace
    ``BertModel``, computed by the fusion path's CUDA kernels.

    * eval / no_grad: the inference forward (no dropout, nothing saved);
    * training (``bert.training`` and autograd on) in the state the reference trains it in -- ``freeze_all_but_bn``
      (modeling/commons.py:33-42): every Linear / embedding frozen, the LayerNorm parameters trainable, dropouts on --
      ``_BertEncoderFn`` runs forward and backward (LayerNorm gradients; no weight gradients are needed) with the counter-
      based dropout of the fusion path;
    * any other trainable parameter (``unfreeze_embeddings``, train_ep >= 0) raises: those weight gradients are not built."""
    if not input_ids.is_cuda:
        raise RuntimeError("transfusion_b200: the MiniLM encoder has no CPU implementation (CUDA tensors required)")
    trainable = [k for k, p in bert.named_parameters() if p.requires_grad]
    if torch.is_grad_enabled() and trainable:
        bad = [k for k in trainable if "LayerNorm" not in k]
        if bad:
            raise NotImplementedError(f"bert_encoder_forward trains LayerNorm parameters only (freeze_all_but_bn); trainable weights "
                                      f"such as {bad[0]} need weight gradients that are not built")
        return _bert_encoder_train(bert, input_ids, attention_mask, token_type_ids)
    with torch.no_grad():
        return _bert_encoder_forward(bert, input_ids, attention_mask, token_type_ids)


_SITE_EMB, _SITE_ATTN, _SITE_DROP1, _SITE_DROP2 = 0x4000, 0x4001, 0x4002, 0x4003


def _site(layer: int, site: int) -> int:
    return (site + 16 * layer) & 0xFFFFFFFF


class _BertEncoderFn(torch.autograd.Function):
    """forward + backward of the BERT encoder stack with frozen matrices and trainable LayerNorms.
    apply(meta, x0 [M, D] bf16 (embedding sum), emb_ln_w, emb_ln_b, ln1_w_0, ln1_b_0, ln2_w_0, ln2_b_0, ...) -> fp32 [M, D]."""

    @staticmethod
    def forward(ctx, meta, x0, *ln):
        B, L, D, H, eps, kpm, seed, p_h, p_a, weights = meta
        dev, bf, f32 = x0.device, torch.bfloat16, torch.float32
        M, d = B * L, D // H
        scale = 1.0 / math.sqrt(d)
        Lp = (L + 127) // 128 * 128

        def empty(*shape, dtype=bf):
            return torch.empty(*shape, device=dev, dtype=dtype)

        mean0, rstd0 = empty(M, dtype=f32), empty(M, dtype=f32)
        x = empty(M, D)
        ops.layernorm_fwd(x0, x, ln[0], ln[1], mean0, rstd0, M, D, eps=eps, drop_p=p_h, drop_seed=seed, drop_stream=_site(0, _SITE_EMB))
        saved = []
        for li, (wqkv, bqkv, wo, bo, w1, b1, w2, b2) in enumerate(weights):
            n1w, n1b, n2w, n2b = ln[2 + 4 * li: 6 + 4 * li]
            F = w1.shape[0]
            qkv = empty(M, 3 * D)
            ops.gemm(x, wqkv, qkv, M=M, N=3 * D, K=D, bias=bqkv)
            att, lse = empty(M, D), empty(B, H, Lp, dtype=f32)
            ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], att, lse, B=B, H=H, Sq=L, Sk=L, dp=d, scale=scale,
                         key_padding_mask=kpm, kpm_start=0, drop_p=p_a, drop_seed=seed, drop_stream=_site(li, _SITE_ATTN))
            y1 = empty(M, D)
            ops.gemm(att, wo, y1, M=M, N=D, K=D, bias=bo, residual=x, drop_p=p_h, drop_seed=seed, drop_stream=_site(li, _SITE_DROP1))
            x1, mean1, rstd1 = empty(M, D), empty(M, dtype=f32), empty(M, dtype=f32)
            ops.layernorm_fwd(y1, x1, n1w, n1b, mean1, rstd1, M, D, eps=eps)
            u, h = empty(M, F), empty(M, F)
            ops.gemm(x1, w1, h, M=M, N=F, K=D, bias=b1, act=1, preact_out=u)
            y2 = empty(M, D)
            ops.gemm(h, w2, y2, M=M, N=D, K=F, bias=b2, residual=x1, drop_p=p_h, drop_seed=seed, drop_stream=_site(li, _SITE_DROP2))
            x2, mean2, rstd2 = empty(M, D), empty(M, dtype=f32), empty(M, dtype=f32)
            ops.layernorm_fwd(y2, x2, n2w, n2b, mean2, rstd2, M, D, eps=eps)
            saved.append((qkv, att, lse, y1, mean1, rstd1, u, y2, mean2, rstd2))
            x = x2
        ctx.meta, ctx.saved, ctx.x0, ctx.stat0, ctx.ln = meta, saved, x0, (mean0, rstd0), ln
        return x.float()

    @staticmethod
    def backward(ctx, d_out):
        B, L, D, H, eps, kpm, seed, p_h, p_a, weights = ctx.meta
        ln = ctx.ln
        dev, bf, f32 = d_out.device, torch.bfloat16, torch.float32
        M, d = B * L, D // H
        scale = 1.0 / math.sqrt(d)
        Lp = (L + 127) // 128 * 128

        def empty(*shape, dtype=bf):
            return torch.empty(*shape, device=dev, dtype=dtype)

        dcur = empty(M, D)
        ops.cast_pad(d_out.reshape(M, D).float().contiguous(), dcur, M, D)
        grads = [None] * len(ln)
        ws = torch.empty(ops.attn_bwd_workspace_bytes(B, H, L, L), device=dev, dtype=torch.uint8)
        for li in reversed(range(len(weights))):
            wqkv, bqkv, wo, bo, w1, b1, w2, b2 = weights[li]
            qkv, att, lse, y1, mean1, rstd1, u, y2, mean2, rstd2 = ctx.saved[li]
            n1w, n1b, n2w, n2b = ln[2 + 4 * li: 6 + 4 * li]
            F = w1.shape[0]
            g_n2w, g_n2b = torch.zeros(D, device=dev, dtype=f32), torch.zeros(D, device=dev, dtype=f32)
            dy2 = empty(M, D)
            g2 = empty(M, D) if p_h > 0 else None
            ops.layernorm_bwd(dcur, y2, n2w, mean2, rstd2, dy2, g_n2w, g_n2b, M, D, dx2=g2, dx2_drop=(p_h, seed, _site(li, _SITE_DROP2)))
            G2 = g2 if g2 is not None else dy2
            du = empty(M, F)
            ops.gemm(G2, w2, du, M=M, N=F, K=D, b_mn_major=True, dact_in=u)
            dx1 = empty(M, D)
            ops.gemm(du, w1, dx1, M=M, N=D, K=F, b_mn_major=True, residual=dy2)
            g_n1w, g_n1b = torch.zeros(D, device=dev, dtype=f32), torch.zeros(D, device=dev, dtype=f32)
            dy1 = empty(M, D)
            g1 = empty(M, D) if p_h > 0 else None
            ops.layernorm_bwd(dx1, y1, n1w, mean1, rstd1, dy1, g_n1w, g_n1b, M, D, dx2=g1, dx2_drop=(p_h, seed, _site(li, _SITE_DROP1)))
            G1 = g1 if g1 is not None else dy1
            datt = empty(M, D)
            ops.gemm(G1, wo, datt, M=M, N=D, K=D, b_mn_major=True)
            delta = empty(B, H, Lp, dtype=f32)
            ops.attn_delta(att, datt, delta, B, L, H, d)
            dqkv = empty(M, 3 * D)
            ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], datt, lse, delta, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:],
                         B=B, H=H, Sq=L, Sk=L, dp=d, scale=scale, key_padding_mask=kpm, kpm_start=0, drop_p=p_a, drop_seed=seed,
                         drop_stream=_site(li, _SITE_ATTN), workspace=ws)
            dxin = empty(M, D)
            ops.gemm(dqkv, wqkv, dxin, M=M, N=D, K=3 * D, b_mn_major=True, residual=dy1)
            grads[2 + 4 * li: 6 + 4 * li] = [g_n1w, g_n1b, g_n2w, g_n2b]
            dcur = dxin
            ctx.saved[li] = None
        g0w, g0b = torch.zeros(D, device=dev, dtype=f32), torch.zeros(D, device=dev, dtype=f32)
        mean0, rstd0 = ctx.stat0
        dx0 = empty(M, D)
        ops.layernorm_bwd(dcur, ctx.x0, ln[0], mean0, rstd0, dx0, g0w, g0b, M, D, dy_drop=(p_h, seed, _site(0, _SITE_EMB)))
        grads[0], grads[1] = g0w, g0b
        return (None, None, *[g if p.requires_grad else None for g, p in zip(grads, ln)])


def _bert_encoder_train(bert, input_ids, attention_mask, token_type_ids):
    cfg = bert.config
    if cfg.hidden_act not in ("gelu",) or getattr(cfg, "position_embedding_type", "absolute") != "absolute":
        raise NotImplementedError("only GELU(erf) / absolute-position BERT encoders (MiniLM-L12-H384) are supported")
    dev, bf = input_ids.device, torch.bfloat16
    B, L = input_ids.shape
    D, H = cfg.hidden_size, cfg.num_attention_heads
    if (D // H) % 32 or D % 8:
        raise NotImplementedError("head_dim must be a multiple of 32")
    emb = bert.embeddings
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    with torch.no_grad():
        pos = torch.arange(L, device=dev)
        e = emb.word_embeddings.weight[input_ids] + emb.position_embeddings.weight[pos][None] + emb.token_type_embeddings.weight[token_type_ids]
        x0 = e.reshape(B * L, D).to(bf).contiguous()
        kpm = (attention_mask == 0).to(torch.uint8).contiguous() if attention_mask is not None else None
        weights = []
        for layer in bert.encoder.layer:
            wqkv, bqkv, wo, w1, w2 = _layer_weights(layer, D, dev)
            weights.append((wqkv, bqkv, wo, layer.attention.output.dense.bias.detach().float(), w1,
                            layer.intermediate.dense.bias.detach().float(), w2, layer.output.dense.bias.detach().float()))
    train = bert.training
    p_h = float(cfg.hidden_dropout_prob) if train else 0.0
    p_a = float(cfg.attention_probs_dropout_prob) if train else 0.0
    seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if (p_h > 0 or p_a > 0) else 0
    ln = [emb.LayerNorm.weight, emb.LayerNorm.bias]
    for layer in bert.encoder.layer:
        ln += [layer.attention.output.LayerNorm.weight, layer.attention.output.LayerNorm.bias,
               layer.output.LayerNorm.weight, layer.output.LayerNorm.bias]
    meta = (B, L, D, H, float(cfg.layer_norm_eps), kpm, seed, p_h, p_a, weights)
    return _BertEncoderFn.apply(meta, x0, *ln).view(B, L, D)


def _bert_encoder_forward(bert, input_ids, attention_mask, token_type_ids):
    cfg = bert.config
    if cfg.hidden_act not in ("gelu",) or getattr(cfg, "position_embedding_type", "absolute") != "absolute":
        raise NotImplementedError("only GELU(erf) / absolute-position BERT encoders (MiniLM-L12-H384) are supported")
    dev, bf = input_ids.device, torch.bfloat16
    B, L = input_ids.shape
    D, H = cfg.hidden_size, cfg.num_attention_heads
    d = D // H
    if d % 32 or D % 8:
        raise NotImplementedError(f"head_dim {d} must be a multiple of 32")
    eps = float(cfg.layer_norm_eps)
    emb = bert.embeddings
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    pos = torch.arange(L, device=dev)
    # embedding-table gathers (index plumbing) in torch; everything from the LayerNorm on runs on the C ABI
    e = emb.word_embeddings.weight[input_ids] + emb.position_embeddings.weight[pos][None] + emb.token_type_embeddings.weight[token_type_ids]
    M = B * L
    x0 = e.reshape(M, D).to(bf).contiguous()
    x = torch.empty(M, D, device=dev, dtype=bf)
    ops.layernorm_fwd(x0, x, emb.LayerNorm.weight.detach().float(), emb.LayerNorm.bias.detach().float(), None, None, M, D, eps=eps)
    kpm = None
    if attention_mask is not None:
        kpm = (attention_mask == 0).to(torch.uint8).contiguous()
    scale = 1.0 / math.sqrt(d)
    for layer in bert.encoder.layer:
        wqkv, bqkv, wo, w1, w2 = _layer_weights(layer, D, dev)
        F = w1.shape[0]
        att_o, inter, out = layer.attention.output, layer.intermediate, layer.output
        qkv = torch.empty(M, 3 * D, device=dev, dtype=bf)
        ops.gemm(x, wqkv, qkv, M=M, N=3 * D, K=D, bias=bqkv)
        att = torch.empty(M, D, device=dev, dtype=bf)
        ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], att, None, B=B, H=H, Sq=L, Sk=L, dp=d, scale=scale,
                     key_padding_mask=kpm, kpm_start=0)
        y1 = torch.empty(M, D, device=dev, dtype=bf)
        ops.gemm(att, wo, y1, M=M, N=D, K=D, bias=att_o.dense.bias.detach().float(), residual=x)
        x1 = torch.empty(M, D, device=dev, dtype=bf)
        ops.layernorm_fwd(y1, x1, att_o.LayerNorm.weight.detach().float(), att_o.LayerNorm.bias.detach().float(), None, None, M, D, eps=eps)
        h = torch.empty(M, F, device=dev, dtype=bf)
        ops.gemm(x1, w1, h, M=M, N=F, K=D, bias=inter.dense.bias.detach().float(), act=1)
        y2 = torch.empty(M, D, device=dev, dtype=bf)
        ops.gemm(h, w2, y2, M=M, N=D, K=F, bias=out.dense.bias.detach().float(), residual=x1)
        x = torch.empty(M, D, device=dev, dtype=bf)
        ops.layernorm_fwd(y2, x, out.LayerNorm.weight.detach().float(), out.LayerNorm.bias.detach().float(), None, None, M, D, eps=eps)
    return x.float().view(B, L, D)


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on xf_gemm (bf16 operands, fp32 accumulate, fp32 output); backward: dgrad (if needed), wgrad, bias grad."""

    @staticmethod
    def forward(ctx, x, W, b, training):
        if not x.is_cuda:
            raise RuntimeError("transfusion_b200: XfLinear has no CPU implementation (CUDA tensors required)")
        dev, bf = x.device, torch.bfloat16
        lead, K = x.shape[:-1], x.shape[-1]
        N = W.shape[0]
        x2 = x.reshape(-1, K)
        R = x2.shape[0]
        casts = []
        need_grad = any(ctx.needs_input_grad)
        w_b = bf16_weight(W, N, K, casts, not (training and need_grad), dev)
        x_b = torch.empty(R, K, device=dev, dtype=bf)
        casts.append((x2.detach().float().contiguous(), x_b, R, K, 0, 0, 0, 0))
        ops.cast_pad_multi(casts)
        y = torch.empty(R, N, device=dev, dtype=torch.float32)
        ops.gemm(x_b, w_b, y, M=R, N=N, K=K, bias=b.detach().float() if b is not None else None)
        if need_grad:
            ctx.save_for_backward(x_b, w_b)
            ctx.meta = (lead, K, N, b is not None, x.dtype)
        return y.view(*lead, N)

    @staticmethod
    def backward(ctx, dy):
        x_b, w_b = ctx.saved_tensors
        lead, K, N, has_b, x_dtype = ctx.meta
        dev, bf, f32 = x_b.device, torch.bfloat16, torch.float32
        R = x_b.shape[0]
        dy_b = torch.empty(R, N, device=dev, dtype=bf)
        ops.cast_pad(dy.reshape(R, N).float().contiguous(), dy_b, R, N)
        gW = gb = dx = None
        if ctx.needs_input_grad[1]:
            gW = torch.zeros(N, K, device=dev, dtype=f32)
            ops.gemm(dy_b, x_b, gW, M=N, N=K, K=R, a_mn_major=True, b_mn_major=True, accumulate=True, split_k=max(1, min(8, R // 512)))
        if has_b and ctx.needs_input_grad[2]:
            gb = torch.zeros(N, device=dev, dtype=f32)
            ops.colsum(dy_b, gb, R, N)
        if ctx.needs_input_grad[0]:
            dxb = torch.empty(R, K, device=dev, dtype=bf)
            ops.gemm(dy_b, w_b, dxb, M=R, N=K, K=N, b_mn_major=True)
            dx = dxb.to(x_dtype).view(*lead, K)
        return dx, gW, gb, None


class XfLinear(nn.Linear):
    """nn.Linear whose forward / backward run on xf_gemm (N, K multiples of 8).  Same parameters, same state_dict."""

    def forward(self, x):
        return _LinearFn.apply(x, self.weight, self.bias, self.training)


class SBertTokensXf(nn.Module):
    """The tensor part of ``SBertLayer.forward`` in token mode (narr_pooling_layers.py:160-202) for already tokenised input:
    BERT encoder -> token_embeddings -> out_mlp -> (tanh) -> dropout; returns ``(embeddings [B, L, D], None, attention_mask)``
    like the reference with ``pad_mask=True``.  ``bert`` is the HuggingFace model sentence-transformers wraps
    (``encoder[0].auto_model``); ``out_mlp`` an ``nn.Linear(384, D)`` whose parameters are shared, not copied."""

    def __init__(self, bert, out_mlp: Optional[nn.Linear] = None, out_tanh: bool = False, out_dropout: float = 0.0):
        super().__init__()
        self.bert = bert
        self.out_mlp = out_mlp
        self.use_out_tanh = out_tanh
        self.out_dropout = nn.Dropout(out_dropout)

    def forward(self, tokenized, pad_mask: bool = False):
        emb = bert_encoder_forward(self.bert, tokenized["input_ids"], tokenized.get("attention_mask"), tokenized.get("token_type_ids"))
        if self.out_mlp is not None:
            emb = _LinearFn.apply(emb, self.out_mlp.weight, self.out_mlp.bias, self.training)
        if self.use_out_tanh:
            emb = torch.tanh(emb)
        return self.out_dropout(emb), None, (tokenized.get("attention_mask") if pad_mask else None)
