"""The per-FPN-level schedule of the fusion hot path as one ``torch.autograd.Function`` over the
C-ABI CUDA library (cross_f_box_wrapper.py:177-212 + cross_f_box_layers.py:69-108 + the encoder
layers of torch18_adapters.py:108-113).  Torch supplies device memory, streams and the autograd
edge; every FLOP and every byte moved on this path is a hand-written sm_100a kernel.

Data layout in HBM (bf16 activations, fp32 statistics / gradients):
  tok   [B*n, C*p*p]   patchified feature map (A operand of the patch-embed GEMM)
  z     [B, S, D]      token sequence, S = n + L; visual rows first, language rows last
  qkv   [B*S, 3*H*dp]  fused in-proj output, head hd of Q/K/V at columns (w*H + hd)*dp (dp = head
                       dim padded to a multiple of 32 with zero weights, e.g. 178 -> 192)
  att   [B*S, H*dp]    attention output, heads merged (A operand of out_proj)
  u, h  [B*S, 2D]      FFN pre-activation / activation
  lse   [B, H, Sp]     log2-domain logsumexp (Sp = S rounded up to 128)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from .. import ops
from ..weight_cache import bf16_weight

LN_EPS = 1e-5

# Attention backward schedule: 1 = score tiles computed once, dQ / dV as batched GEMMs over a transient bf16 [B, H, S, S]
# scratch (5 GEMM units); 0 = three on-chip passes that recompute the scores (8 units, no S x S bytes in HBM)
import os as _os
ATTN_BWD_WORKSPACE = bool(int(_os.environ.get("XF_ATTN_BWD_WS", "1")))

# dropout sites inside one encoder layer (stream ids for the counter-based RNG)
SITE_ATTN, SITE_DROP1, SITE_FFN, SITE_DROP2 = 0, 1, 2, 3
SITE_PATCH, SITE_BACKPROJ = 60, 61


@dataclass
class LevelConfig:
    level: int
    patch: int
    num_heads: int
    num_layers: int
    training: bool
    patch_dropout: float
    token_dropout: float
    backproj_dropout: float
    seed: int = 0
    need_lang_out: bool = False
    # stream that consumes the level's outputs when the level itself runs on a side stream: the outputs are then
    # allocated from THAT stream's pool (no record_stream, so no deferred frees that make the allocator's needs
    # depend on how far the host runs ahead)
    out_stream: object = None
    # share of the SMs this level's persistent GEMM grids may occupy when the levels run concurrently on side streams
    # (0 = whole machine)
    gemm_ctas: int = 0
    # the level's FPN lateral (inner 1x1 conv, C -> Co) is folded into the back-projection; the function then
    # returns the [B, Co, h, w] lateral map instead of the fused [B, C, h, w] map and takes two more parameters
    lateral: bool = False

    def stream(self, layer: int, site: int) -> int:
        return ((self.level * 16 + layer) * 64 + site) & 0xFFFFFFFF


# dev aid: when set to a dict, the forward stores clones of its intermediates in it (tools/gpu_probe.py)
DEBUG_SINK = None


def _dbg(name, t):
    if DEBUG_SINK is not None and t is not None:
        DEBUG_SINK[name] = t.detach().float().clone()


N_HEAD_PARAMS = 4       # patch-embed weight, image_kind, lang_kind, pos table (buffer, no grad)
N_LAYER_PARAMS = 12
N_TAIL_PARAMS = 4       # final LN weight/bias, back-projection weight/bias
N_LATERAL_PARAMS = 2    # optional: FPN inner 1x1 conv weight [Co, C, 1, 1] / bias [Co] folded into the back-projection (8f N1)


N_CLUSTERS = 74  # CTA pairs on a 148-SM B200


def _split_k_for(tiles: int, total_kb: int, clusters: int = N_CLUSTERS) -> int:
    """Split-K factor for a wgrad GEMM with `tiles` output tiles and `total_kb` 64-token k-blocks: minimise
    waves x (k-blocks per item + fixed per-item cost) over the persistent grid of CTA pairs."""
    best, best_cost = 1, None
    for s in range(1, max(1, min(32, total_kb // 8)) + 1):
        items = tiles * s
        waves = -(-items // clusters)
        cost = waves * (-(-total_kb // s) + 6)
        if best_cost is None or cost < best_cost:
            best, best_cost = s, cost
    return best


def _wgrad(dy: torch.Tensor, x: torch.Tensor, out_f32: torch.Tensor, n_out: int, k_in: int, tokens: int, ctas: int = 0):
    """out_f32[n_out, k_in] += dy[tokens, n_out]^T @ x[tokens, k_in]  (both operands MN-major)."""
    tn = next((c for c in (256, 224, 192, 160, 128) if k_in % c == 0), 256)   # mirrors pick_tile_n in gemm.cu
    tiles = ((n_out + 255) // 256) * ((k_in + tn - 1) // tn)
    split = _split_k_for(tiles, (tokens + 63) // 64, max(1, ctas // 2) if ctas > 0 else N_CLUSTERS)
    ops.gemm(dy, x, out_f32, M=n_out, N=k_in, K=tokens, a_mn_major=True, b_mn_major=True, accumulate=True,
             split_k=split)


class FusionLevelFunction(torch.autograd.Function):
    """fused, lang_out = f(feat, lang, key_pad, *params).  See module docstring."""

    @staticmethod
    def forward(ctx, cfg: LevelConfig, feat: torch.Tensor, lang: torch.Tensor, key_pad: Optional[torch.Tensor], *params):
        if not feat.is_cuda:
            raise RuntimeError("transfusion_b200: the fusion path has no CPU implementation (CUDA tensors required)")
        prev_cap = ops.set_gemm_cta_cap(cfg.gemm_ctas)
        try:
            return FusionLevelFunction._forward(ctx, cfg, feat, lang, key_pad, *params)
        finally:
            ops.set_gemm_cta_cap(prev_cap)

    @staticmethod
    def _forward(ctx, cfg: LevelConfig, feat: torch.Tensor, lang: torch.Tensor, key_pad: Optional[torch.Tensor], *params):
        dev = feat.device
        B, C, Hf, Wf = feat.shape
        p = cfg.patch
        gh, gw = Hf // p, Wf // p
        n = gh * gw
        L = lang.shape[1]
        D = lang.shape[2]
        S = n + L
        H = cfg.num_heads
        d = D // H
        dp = (d + 31) // 32 * 32
        Dp = H * dp
        F = params[N_HEAD_PARAMS + 4].shape[0]  # linear1.weight [F, D]
        K = C * p * p
        M = B * S
        Sp = (S + 127) // 128 * 128
        nl = cfg.num_layers
        train = cfg.training
        need_grad = any(ctx.needs_input_grad)  # grad mode is off inside Function.forward; this is the autograd edge
        bf = torch.bfloat16
        scale = 1.0 / math.sqrt(d)

        wpe, img_kind, lang_kind, pos_table = params[:N_HEAD_PARAMS]
        layer_params = [params[N_HEAD_PARAMS + i * N_LAYER_PARAMS: N_HEAD_PARAMS + (i + 1) * N_LAYER_PARAMS] for i in range(nl)]
        tail = params[N_HEAD_PARAMS + nl * N_LAYER_PARAMS:]
        lnf_w, lnf_b, wbp, bbp = tail[:N_TAIL_PARAMS]
        lat_w, lat_b = (tail[N_TAIL_PARAMS:] if cfg.lateral else (None, None))

        def empty(*shape, dtype=bf):
            return torch.empty(*shape, device=dev, dtype=dtype)

        # channels_last maps (a channels_last backbone) are consumed and produced as they are: the layout
        # kernels read / write NHWC memory directly, no conversion pass; any other non-contiguous layout is made NCHW
        cl = feat.dim() == 4 and not feat.is_contiguous() and feat.is_contiguous(memory_format=torch.channels_last)
        feat_c = feat if cl else feat.contiguous()
        if feat_c.dtype not in (torch.float32, torch.bfloat16):
            feat_c = feat_c.float()
        mem_fmt = torch.channels_last if cl else torch.contiguous_format
        lang_c = lang.contiguous().float()
        kpm = None
        if key_pad is not None:
            kpm = torch.zeros(B, S, device=dev, dtype=torch.uint8)
            kpm[:, n:] = key_pad.to(torch.uint8)

        pd_patch = cfg.patch_dropout if train else 0.0
        pd_tok = cfg.token_dropout if train else 0.0
        pd_back = cfg.backproj_dropout if train else 0.0
        seed = cfg.seed

        # ---- bf16 weight copies (head dim padded d -> dp with zero rows / columns): one launch for the level; cached per
        # parameter (weight_cache.py: produced by FusedRAdam in training, version-keyed in inference)
        casts = []
        trust_version = not (train and need_grad)

        def bf16_of(prm, rows, cols, pad=None):
            return bf16_weight(prm, rows, cols, casts, trust_version, dev, pad)

        wpe_b = bf16_of(wpe, D, K)
        wbp_b = bf16_of(wbp, K, D)
        lw = []
        for (in_w, in_b, out_w, out_b, w1, b1, w2, b2, n1w, n1b, n2w, n2b) in layer_params:
            if dp != d:
                win_b = bf16_of(in_w, 3 * D, D, pad=(d, dp, 0, 0, 3 * Dp, D))
                bin_p = torch.zeros(3 * H, dp, device=dev, dtype=torch.float32)
                bin_p[:, :d] = in_b.reshape(3 * H, d)
                bin_p = bin_p.reshape(3 * Dp)
                wo_b = bf16_of(out_w, D, D, pad=(0, 0, d, dp, D, Dp))
            else:
                win_b = bf16_of(in_w, 3 * D, D)
                bin_p = in_b
                wo_b = bf16_of(out_w, D, D)
            w1_b = bf16_of(w1, F, D)
            w2_b = bf16_of(w2, D, F)
            lw.append((win_b, bin_p, wo_b, w1_b, w2_b))
        if casts:
            ops.cast_pad_multi(casts)

        # ---- patch embedding + positional / kind embeddings, language rows   (K1-K4)
        tok = empty(B * n, K)
        ops.patchify(feat_c, p, tok)
        z = empty(B, S, D)
        z2 = z.view(M, D)
        ops.gemm(tok, wpe_b, z2, M=B * n, N=D, K=K, bias=img_kind.reshape(D), pos_table=pos_table,
                 rows_in=n, rows_out=S, drop_p=pd_patch, drop_seed=seed, drop_stream=cfg.stream(0, SITE_PATCH))
        ops.lang_rows_fwd(lang_c, lang_kind.reshape(D).contiguous(), z, n)
        _dbg("tok", tok); _dbg("z0", z)

        saved_layers = []
        x = z2
        for l in range(nl):
            (in_w, in_b, out_w, out_b, w1, b1, w2, b2, n1w, n1b, n2w, n2b) = layer_params[l]
            win_b, bin_p, wo_b, w1_b, w2_b = lw[l]
            qkv = torch.zeros(M, 3 * Dp, device=dev, dtype=bf) if False else empty(M, 3 * Dp)
            ops.gemm(x, win_b, qkv, M=M, N=3 * Dp, K=D, bias=bin_p)                                   # K5
            att = empty(M, Dp)
            lse = empty(B, H, Sp, dtype=torch.float32) if need_grad else None
            ops.attn_fwd(qkv[:, :Dp], qkv[:, Dp:2 * Dp], qkv[:, 2 * Dp:], att, lse, B=B, H=H, Sq=S, Sk=S, dp=dp,
                         scale=scale, key_padding_mask=kpm, kpm_start=n, drop_p=pd_tok, drop_seed=seed,
                         drop_stream=cfg.stream(l, SITE_ATTN))                                       # K6
            y1 = empty(M, D)
            ops.gemm(att, wo_b, y1, M=M, N=D, K=Dp, bias=out_b, residual=x, drop_p=pd_tok, drop_seed=seed,
                     drop_stream=cfg.stream(l, SITE_DROP1))                                           # K7
            x1 = empty(M, D)
            mean1 = empty(M, dtype=torch.float32) if need_grad else None
            rstd1 = empty(M, dtype=torch.float32) if need_grad else None
            ops.layernorm_fwd(y1, x1, n1w, n1b, mean1, rstd1, M, D, eps=LN_EPS)
            u = empty(M, F) if need_grad else None
            h = empty(M, F)
            ops.gemm(x1, w1_b, h, M=M, N=F, K=D, bias=b1, act=1, preact_out=u, drop_p=pd_tok, drop_seed=seed,
                     drop_stream=cfg.stream(l, SITE_FFN))                                             # K8
            y2 = empty(M, D)
            ops.gemm(h, w2_b, y2, M=M, N=D, K=F, bias=b2, residual=x1, drop_p=pd_tok, drop_seed=seed,
                     drop_stream=cfg.stream(l, SITE_DROP2))                                           # K9
            x2 = empty(M, D)
            mean2 = empty(M, dtype=torch.float32) if need_grad else None
            rstd2 = empty(M, dtype=torch.float32) if need_grad else None
            ops.layernorm_fwd(y2, x2, n2w, n2b, mean2, rstd2, M, D, eps=LN_EPS)
            for nm, t in (("qkv", qkv), ("att", att), ("y1", y1), ("x1", x1), ("h", h), ("y2", y2), ("x2", x2)):
                _dbg(f"l{l}.{nm}", t)
            if need_grad:
                saved_layers.append((x, qkv, att, lse, y1, mean1, rstd1, x1, u, h, y2, mean2, rstd2))
            x = x2

        # ---- final LN on the visual rows, back-projection, fold   (K10-K12)
        vis = empty(B * n, D)
        meanf = empty(B * n, dtype=torch.float32) if need_grad else None
        rstdf = empty(B * n, dtype=torch.float32) if need_grad else None
        ops.layernorm_fwd(x, vis, lnf_w, lnf_b, meanf, rstdf, B * n, D, in_map=(n, S, 0), eps=LN_EPS,
                          drop_p=pd_back, drop_seed=seed, drop_stream=cfg.stream(0, SITE_BACKPROJ))
        w2_b = wl_b = None
        Co, Ko = C, K
        if cfg.lateral:
            # FPN lateral folded into the back-projection (both linear, nothing in between: utils.py:114-119 ->
            # torchvision FeaturePyramidNetwork.inner_blocks[i], faster_rcnn_wrapper.py:419-421):
            #   W2[(o,u,v), d] = sum_c Wl[o,c] Wbp[(c,u,v), d],   b2[(o,u,v)] = sum_c Wl[o,c] bbp[(c,u,v)] + bl[o]
            # One weight-space GEMM per step ([Co, C] x [C, p^2 D]) replaces the [B,C,h,w] fused map, its fold pass and the
            # 1x1 convolution over it; the token GEMM shrinks from N = C p^2 to N = Co p^2 (8x at C5).
            Co = lat_w.shape[0]
            Ko = Co * p * p
            wl_b = empty(Co, C)
            ops.cast_pad(lat_w.reshape(Co, C), wl_b, Co, C)
            w2_b = empty(Co, p * p * D)
            ops.gemm(wl_b, wbp_b.view(C, p * p * D), w2_b, M=Co, N=p * p * D, K=C, b_mn_major=True)
            b2 = (lat_w.reshape(Co, C).float() @ bbp.reshape(C, p * p).float() + lat_b.float()[:, None]).reshape(Ko)   # [Co, p^2]: tiny, weight space
            yb = empty(B * n, Ko)
            ops.gemm(vis, w2_b.view(Ko, D), yb, M=B * n, N=Ko, K=D, bias=b2.detach())
        else:
            yb = empty(B * n, K)
            ops.gemm(vis, wbp_b, yb, M=B * n, N=K, K=D, bias=bbp)
        _dbg("vis", vis); _dbg("yb", yb)
        if cfg.out_stream is not None:
            with torch.cuda.stream(cfg.out_stream):
                fused = torch.empty(B, Co, Hf, Wf, device=dev, dtype=feat_c.dtype, memory_format=mem_fmt)
                lang_out = torch.empty(B, L, D, device=dev, dtype=torch.float32) if cfg.need_lang_out else lang.new_zeros(())
        else:
            fused = torch.empty(B, Co, Hf, Wf, device=dev, dtype=feat_c.dtype, memory_format=mem_fmt)
            lang_out = torch.empty(B, L, D, device=dev, dtype=torch.float32) if cfg.need_lang_out else lang.new_zeros(())
        ops.fold(yb, fused, p)
        if cfg.need_lang_out:
            lang_out.copy_(x.view(B, S, D)[:, n:])

        if need_grad:
            ctx.cfg = cfg
            ctx.dims = (B, C, Hf, Wf, p, n, L, D, S, H, d, dp, Dp, F, K, M, Sp)
            ctx.pd = (pd_patch, pd_tok, pd_back)
            ctx.kpm = kpm
            ctx.tok = tok
            ctx.xfinal = x
            ctx.vis = vis
            ctx.statf = (meanf, rstdf)
            ctx.saved_layers = saved_layers
            ctx.lw = lw
            ctx.wpe_b, ctx.wbp_b = wpe_b, wbp_b
            ctx.lateral = (w2_b, wl_b, Co, Ko)
            ctx.param_refs = params
            ctx.needs = (feat.requires_grad, lang.requires_grad)
            ctx.feat_dtype = feat_c.dtype
            ctx.mem_fmt = mem_fmt
        return fused, lang_out

    @staticmethod
    def backward(ctx, d_fused, d_lang_out):
        prev_cap = ops.set_gemm_cta_cap(ctx.cfg.gemm_ctas)
        try:
            return FusionLevelFunction._backward(ctx, d_fused, d_lang_out)
        finally:
            ops.set_gemm_cta_cap(prev_cap)

    @staticmethod
    def _backward(ctx, d_fused, d_lang_out):
        cfg: LevelConfig = ctx.cfg
        (B, C, Hf, Wf, p, n, L, D, S, H, d, dp, Dp, F, K, M, Sp) = ctx.dims
        pd_patch, pd_tok, pd_back = ctx.pd
        params = ctx.param_refs
        nl = cfg.num_layers
        seed = cfg.seed
        dev = d_fused.device
        bf = torch.bfloat16
        f32 = torch.float32
        scale = 1.0 / math.sqrt(d)
        kpm = ctx.kpm
        wpe, img_kind, lang_kind, pos_table = params[:N_HEAD_PARAMS]
        layer_params = [params[N_HEAD_PARAMS + i * N_LAYER_PARAMS: N_HEAD_PARAMS + (i + 1) * N_LAYER_PARAMS] for i in range(nl)]
        tail = params[N_HEAD_PARAMS + nl * N_LAYER_PARAMS:]
        lnf_w, lnf_b, wbp, bbp = tail[:N_TAIL_PARAMS]
        lat_w, lat_b = (tail[N_TAIL_PARAMS:] if cfg.lateral else (None, None))
        w2_b, wl_b, Co, Ko = ctx.lateral

        def empty(*shape, dtype=bf):
            return torch.empty(*shape, device=dev, dtype=dtype)

        # fp32 gradient buffers (accumulated into by split-K wgrads, column sums and LayerNorm backward) are carved
        # from one zero-filled arena per level: one memset launch instead of ~57 tiny fills.  autograd adopts the
        # views as .grad without copying.
        arena = {"buf": None, "off": 0}
        arena_cap = sum(int(q.numel()) for q in params if q is not None and q.requires_grad) + 8 * len(params) * 64 + (1 << 16)
        if dp != d:   # head-padded scratch gradients (g_win_p, g_bin_p, g_wo_p) live in the arena as well
            arena_cap += nl * (3 * Dp * D + D * Dp + 3 * Dp + 3 * 64)

        def zeros(*shape, dtype=f32):
            if dtype != f32:
                return torch.zeros(*shape, device=dev, dtype=dtype)
            numel = 1
            for s_ in shape:
                numel *= int(s_)
            need = (numel + 63) // 64 * 64   # 256-byte aligned slices
            if arena["buf"] is None or arena["off"] + need > arena["buf"].numel():
                # one buffer per level by construction (arena_cap): a second one would silently break the in-place
                # data-parallel reduction of the arena (parallel.BucketedGradAllReduce)
                assert arena["buf"] is None, "gradient arena overflow: arena_cap underestimates the level's gradients"
                arena["buf"] = torch.zeros(max(need, arena_cap), device=dev, dtype=f32)
                arena["off"] = 0
            out = arena["buf"][arena["off"]:arena["off"] + numel].view(*shape)
            arena["off"] += need
            return out

        grads: List[Optional[torch.Tensor]] = [None] * len(params)

        # the incoming gradients may have been produced on another stream (the caller's): tell the caching allocator
        # that this stream reads them (the autograd engine has already ordered the streams)
        if d_fused.is_cuda:
            cur_stream = torch.cuda.current_stream()
            d_fused.record_stream(cur_stream)
            if isinstance(d_lang_out, torch.Tensor) and d_lang_out.is_cuda:
                d_lang_out.record_stream(cur_stream)
        # ---- fold^T, back-projection
        d_cl = d_fused.dim() == 4 and not d_fused.is_contiguous() and d_fused.is_contiguous(memory_format=torch.channels_last)
        d_fused_c = d_fused if d_cl else d_fused.contiguous()
        if d_fused_c.dtype not in (torch.float32, torch.bfloat16):
            d_fused_c = d_fused_c.float()
        g_lat_w = g_lat_b = None
        dyb = empty(B * n, Ko)
        ops.patchify(d_fused_c, p, dyb)
        dvis = empty(B * n, D)
        if cfg.lateral:
            # gradients of the folded weights, then the chain rule back to (Wl, bl) and (Wbp, bbp) in weight space
            g_b2 = torch.zeros(Ko, device=dev, dtype=f32); ops.colsum(dyb, g_b2, B * n, Ko)
            g_w2 = torch.zeros(Ko, D, device=dev, dtype=f32); _wgrad(dyb, ctx.vis, g_w2, Ko, D, B * n, ctas=cfg.gemm_ctas)
            ops.gemm(dyb, w2_b.view(Ko, D), dvis, M=B * n, N=D, K=Ko, b_mn_major=True)
            pp = p * p
            g_w2_b = empty(Co, pp * D)
            ops.cast_pad(g_w2.view(Co, pp * D), g_w2_b, Co, pp * D)
            # dWl[o,c] = sum_{u,v,d} dW2[(o,u,v),d] Wbp[(c,u,v),d]  (+ the bias term below)
            g_lat_w = torch.zeros(Co, C, device=dev, dtype=f32)
            ops.gemm(g_w2_b, ctx.wbp_b.view(C, pp * D), g_lat_w, M=Co, N=C, K=pp * D, accumulate=True,
                     split_k=max(1, min(32, (pp * D) // 512)))   # few output tiles, long K: spread it over the SMs
            # dWbp[(c,u,v),d] = sum_o Wl[o,c] dW2[(o,u,v),d]
            g_wbp = zeros(K, D)
            ops.gemm(wl_b, g_w2_b, g_wbp.view(C, pp * D), M=C, N=pp * D, K=Co, a_mn_major=True, b_mn_major=True, accumulate=True)
            gb2 = g_b2.view(Co, pp)
            g_lat_w += gb2 @ bbp.detach().reshape(C, pp).float().t()            # tiny weight-space terms ([Co, p^2] x [p^2, C])
            g_lat_b = gb2.sum(1)
            g_bbp = zeros(K)
            g_bbp.view(C, pp).copy_(lat_w.detach().reshape(Co, C).float().t() @ gb2)
            g_lat_w = g_lat_w.view_as(lat_w)
        else:
            g_bbp = zeros(K); ops.colsum(dyb, g_bbp, B * n, K)
            g_wbp = zeros(K, D); _wgrad(dyb, ctx.vis, g_wbp, K, D, B * n, ctas=cfg.gemm_ctas)
            ops.gemm(dyb, ctx.wbp_b, dvis, M=B * n, N=D, K=K, b_mn_major=True)
        del dyb
        # ---- final LN backward into the visual rows of dz; language rows from d_lang_out (or zero)
        dx = empty(B, S, D)
        if cfg.need_lang_out and d_lang_out is not None and d_lang_out.dim() == 3:
            dx[:, n:] = d_lang_out.to(bf)
        else:
            dx[:, n:].zero_()
        dxf = dx.view(M, D)
        g_lnf_w, g_lnf_b = zeros(D), zeros(D)
        meanf, rstdf = ctx.statf
        ops.layernorm_bwd(dvis, ctx.xfinal, lnf_w, meanf, rstdf, dxf, g_lnf_w, g_lnf_b, B * n, D,
                          in_map=(n, S, 0), dy_drop=(pd_back, seed, cfg.stream(0, SITE_BACKPROJ)))
        base_tail = N_HEAD_PARAMS + nl * N_LAYER_PARAMS
        grads[base_tail + 0], grads[base_tail + 1] = g_lnf_w, g_lnf_b
        grads[base_tail + 2], grads[base_tail + 3] = g_wbp, g_bbp
        if cfg.lateral:   # FPN parameters: not part of the level's arena / bucket
            grads[base_tail + 4], grads[base_tail + 5] = g_lat_w, g_lat_b
        del dvis

        # scratch of the 5-unit attention backward (E = scale dS^T as bf16 [B, H, S, S]); one buffer for all layers
        attn_ws = None
        if ATTN_BWD_WORKSPACE:
            attn_ws = torch.empty(ops.attn_bwd_workspace_bytes(B, H, S, S), device=dev, dtype=torch.uint8)
        dcur = dxf  # gradient w.r.t. the layer output x2
        for l in reversed(range(nl)):
            (in_w, in_b, out_w, out_b, w1, b1, w2, b2, n1w, n1b, n2w, n2b) = layer_params[l]
            win_b, bin_p, wo_b, w1_b, w2_b = ctx.lw[l]
            (x, qkv, att, lse, y1, mean1, rstd1, x1, u, h, y2, mean2, rstd2) = ctx.saved_layers[l]
            base = N_HEAD_PARAMS + l * N_LAYER_PARAMS
            # LN2 backward: dY2 (+ dropout2-masked copy G2 feeding linear2's gradients), b2 grad
            dy2 = empty(M, D)
            g2 = empty(M, D) if pd_tok > 0 else None
            g_n2w, g_n2b, g_b2 = zeros(D), zeros(D), zeros(D)
            ops.layernorm_bwd(dcur, y2, n2w, mean2, rstd2, dy2, g_n2w, g_n2b, M, D, dbias=g_b2, dx2=g2,
                              dx2_drop=(pd_tok, seed, cfg.stream(l, SITE_DROP2)))
            G2 = g2 if g2 is not None else dy2
            # linear2: wgrad, dgrad fused with dropout(ffn) mask and GELU'
            g_w2 = zeros(D, F); _wgrad(G2, h, g_w2, D, F, M, ctas=cfg.gemm_ctas)
            du = empty(M, F)
            ops.gemm(G2, w2_b, du, M=M, N=F, K=D, b_mn_major=True, dact_in=u, drop_p=pd_tok, drop_seed=seed,
                     drop_stream=cfg.stream(l, SITE_FFN), drop_first=True)
            # linear1
            g_b1 = zeros(F); ops.colsum(du, g_b1, M, F)
            g_w1 = zeros(F, D); _wgrad(du, x1, g_w1, F, D, M, ctas=cfg.gemm_ctas)
            dx1 = empty(M, D)
            ops.gemm(du, w1_b, dx1, M=M, N=D, K=F, b_mn_major=True, residual=dy2)
            del du
            # LN1 backward
            dy1 = empty(M, D)
            g1 = empty(M, D) if pd_tok > 0 else None
            g_n1w, g_n1b, g_bo = zeros(D), zeros(D), zeros(D)
            ops.layernorm_bwd(dx1, y1, n1w, mean1, rstd1, dy1, g_n1w, g_n1b, M, D, dbias=g_bo, dx2=g1,
                              dx2_drop=(pd_tok, seed, cfg.stream(l, SITE_DROP1)))
            G1 = g1 if g1 is not None else dy1
            # out_proj
            g_wo_p = zeros(D, Dp); _wgrad(G1, att, g_wo_p, D, Dp, M, ctas=cfg.gemm_ctas)
            datt = empty(M, Dp)
            ops.gemm(G1, wo_b, datt, M=M, N=Dp, K=D, b_mn_major=True)
            # attention backward
            delta = empty(B, H, Sp, dtype=f32)
            ops.attn_delta(att, datt, delta, B, S, H, dp)
            dqkv = empty(M, 3 * Dp)
            ops.attn_bwd(qkv[:, :Dp], qkv[:, Dp:2 * Dp], qkv[:, 2 * Dp:], datt, lse, delta,
                         dqkv[:, :Dp], dqkv[:, Dp:2 * Dp], dqkv[:, 2 * Dp:], B=B, H=H, Sq=S, Sk=S, dp=dp, scale=scale,
                         key_padding_mask=kpm, kpm_start=n, drop_p=pd_tok, drop_seed=seed, drop_stream=cfg.stream(l, SITE_ATTN),
                         workspace=attn_ws)
            del datt
            # in_proj
            g_bin_p = zeros(3 * Dp); ops.colsum(dqkv, g_bin_p, M, 3 * Dp)
            g_win_p = zeros(3 * Dp, D); _wgrad(dqkv, x, g_win_p, 3 * Dp, D, M, ctas=cfg.gemm_ctas)
            dxin = empty(M, D)
            ops.gemm(dqkv, win_b, dxin, M=M, N=D, K=3 * Dp, b_mn_major=True, residual=dy1)
            del dqkv
            if dp != d:
                g_win = zeros(3 * D, D); ops.unpad_add(g_win_p, g_win, 3 * D, D, rin=d, rout=dp)
                g_bin = g_bin_p.view(3 * H, dp)[:, :d].reshape(3 * D).contiguous()
                g_wo = zeros(D, D); ops.unpad_add(g_wo_p, g_wo, D, D, cin=d, cout=dp)
            else:
                g_win, g_bin, g_wo = g_win_p, g_bin_p, g_wo_p
            grads[base: base + N_LAYER_PARAMS] = [g_win, g_bin, g_wo, g_bo, g_w1, g_b1, g_w2, g_b2, g_n1w, g_n1b, g_n2w, g_n2b]
            dcur = dxin
            ctx.saved_layers[l] = None

        # ---- sequence assembly backward: language rows and visual rows
        dz0 = dcur.view(B, S, D)
        g_lang_kind = zeros(D)
        # not from the arena: the arena holds parameter gradients only (a data-parallel reducer sums it in place)
        d_lang = torch.zeros(B, L, D, device=dev, dtype=f32) if ctx.needs[1] else None
        ops.lang_rows_bwd(dz0, d_lang, g_lang_kind, B, L, n)
        dz0v = empty(B * n, D)
        g_img_kind = zeros(D)
        ops.rows_gather(dz0, dz0v, B * n, D, in_map=(n, S, 0), colsum=g_img_kind, drop_p=pd_patch, drop_seed=seed,
                        drop_stream=cfg.stream(0, SITE_PATCH))
        g_wpe = zeros(D, K); _wgrad(dz0v, ctx.tok, g_wpe, D, K, B * n, ctas=cfg.gemm_ctas)
        d_feat = None
        if ctx.needs[0]:
            dtok = empty(B * n, K)
            ops.gemm(dz0v, ctx.wpe_b, dtok, M=B * n, N=K, K=D, b_mn_major=True)
            d_feat = torch.empty(B, C, Hf, Wf, device=dev, dtype=ctx.feat_dtype, memory_format=ctx.mem_fmt)
            ops.fold(dtok, d_feat, p)
        grads[0] = g_wpe.view_as(wpe)
        grads[1] = g_img_kind.view_as(img_kind)
        grads[2] = g_lang_kind.view_as(lang_kind)
        grads[3] = None  # positional table: buffer (sin1d)
        out_grads = []
        for prm, g in zip(params, grads):
            if g is None or not prm.requires_grad:
                out_grads.append(None)
            else:
                out_grads.append(g.view_as(prm) if g.shape != prm.shape else g)
        # every parameter gradient of the level lives in the arena: a data-parallel reducer can all-reduce that one
        # buffer in place (parallel.BucketedGradAllReduce) instead of flattening ~70 tensors
        # With gradient accumulation (a .grad already exists) autograd adds this backward's gradients IN PLACE into the first
        # micro-step's arena views, so that first arena stays the one to reduce: only a fresh .grad adopts the new arena.
        if arena["buf"] is not None:
            used = arena["buf"][:arena["off"]]
            for prm in params:
                if prm is not None and prm.requires_grad and (prm.grad is None or getattr(prm, "_xf_grad_arena", None) is None):
                    prm._xf_grad_arena = used
        return (None, d_feat, d_lang, None, *out_grads)
