"""fp32-tolerance mode of one fusion level (north_star: "about 1e-5 relative in fp32"; the reference's Ego4Dv2 config runs
``precision: 32``, runner/nao/configs/ego_nao_res50_ego4dv2.yml:124).  FORWARD only -- a validation / inference mode, selected
with ``CrossFusionBoxWrapper(..., precision="fp32")`` or ``XF_PRECISION=fp32``; autograd through it raises.

Same op sequence as ``level_fn.FusionLevelFunction`` (cross_f_box_wrapper.py:177-212, cross_f_box_layers.py:69-108,
torch18_adapters.py:108-113), activations kept in fp32.  Every contraction still runs on the tcgen05 GEMM kernel: fp32
operands are split into three bf16 terms concatenated along K (``ops.split3``: K' = 6K, fp32 accumulate in TMEM, fp32
output), attention materialises the fp32 score tensor per (sample, head) -- this mode trades memory and 6x the MMA work
for accuracy -- with the softmax in ``xf_softmax_rows_f32`` and LayerNorm in the fp32 ``xf_rowln_fwd`` kernel.  Torch is
used for layout permutations, zero fills and residual pre-fills only (no arithmetic)."""
from __future__ import annotations

import math

import torch

from .. import ops
from .level_fn import LN_EPS, N_HEAD_PARAMS, N_LAYER_PARAMS, N_TAIL_PARAMS, LevelConfig


def _chains(k_elems: int) -> int:
    """split-K factor: accumulator chains of <= ~20 MMAs (the tensor core truncates on every add into its fp32 accumulator)."""
    kb = (k_elems + 63) // 64
    return max(1, min(32, kb // 5))


def _gemm32(x, W, bias=None, act_in=0, into=None):
    """fp32-accurate act(x) [R, K] @ W [N, K]^T (+ bias) -> fp32 [R, N], accumulated ONTO `into` when given (residual add,
    positional pre-fill).  The bias rides inside the GEMM (8 extra K columns), the sum is split over short chains."""
    R, K = x.shape
    N = W.shape[0]
    bc = 8 if bias is not None else 0
    xa = ops.split3(x, 0, act=act_in, bias_cols=bc)
    wb = ops.split3(W, 1, bias=bias, bias_cols=bc)
    out = into if into is not None else torch.zeros(R, N, device=x.device, dtype=torch.float32)
    ops.gemm(xa, wb, out, M=R, N=N, K=6 * K + bc, accumulate=True, split_k=_chains(6 * K + bc))
    return out


def _attention32(qkv, B, S, H, d, kpm):
    """softmax(Q K^T / sqrt(d) + key padding) V per (sample, head) in fp32: two batched split-GEMMs around the softmax."""
    dev = qkv.device
    D = H * d
    q, k, v = (qkv[:, i * D:(i + 1) * D].reshape(B, S, H, d).permute(0, 2, 1, 3).contiguous() for i in range(3))   # [B,H,S,d]
    Sp = (S + 7) // 8 * 8
    qa = ops.split3(q.view(B * H * S, d), 0)                                   # [BH*S, 6d]
    kp = torch.zeros(B * H, Sp, d, device=dev, dtype=torch.float32)
    kp[:, :S] = k.view(B * H, S, d)
    kb = ops.split3(kp.view(B * H * Sp, d), 1)                                 # [BH*Sp, 6d]
    scores = torch.zeros(B, H, S, Sp, device=dev, dtype=torch.float32)
    ops.gemm(qa, kb, scores, M=S, N=Sp, K=6 * d, accumulate=True, split_k=_chains(6 * d), a_ld=6 * d, b_ld=6 * d, ldc=Sp,
             batch=(B * H, 1, (S * 6 * d, 0), (Sp * 6 * d, 0), (S * Sp, 0)))
    ops.softmax_rows_f32(scores, S, kpm, 1.0 / math.sqrt(d))
    pa = ops.split3(scores.view(B * H * S, Sp), 0)                             # [BH*S, 6Sp]
    vt = torch.zeros(B * H, d, Sp, device=dev, dtype=torch.float32)
    vt[:, :, :S] = v.view(B * H, S, d).transpose(1, 2)
    vb = ops.split3(vt.view(B * H * d, Sp), 1)                                 # [BH*d, 6Sp]
    o = torch.zeros(B * H, S, d, device=dev, dtype=torch.float32)
    ops.gemm(pa, vb, o, M=S, N=d, K=6 * Sp, accumulate=True, split_k=_chains(6 * Sp), a_ld=6 * Sp, b_ld=6 * Sp, ldc=d,
             batch=(B * H, 1, (S * 6 * Sp, 0), (d * 6 * Sp, 0), (S * d, 0)))
    return o.view(B, H, S, d).permute(0, 2, 1, 3).reshape(B * S, D).contiguous()


def fusion_level_forward_fp32(cfg: LevelConfig, feat, lang, key_pad, *params):
    """Returns (fused [B, C, h, w] fp32, lang_out [B, L, D] fp32)."""
    if not feat.is_cuda:
        raise RuntimeError("transfusion_b200: the fusion path has no CPU implementation (CUDA tensors required)")
    if torch.is_grad_enabled() and (feat.requires_grad or lang.requires_grad or any(p is not None and p.requires_grad for p in params)):
        raise NotImplementedError("precision='fp32' is a forward-only validation / inference mode: run it under torch.no_grad()")
    if cfg.training and (cfg.patch_dropout > 0 or cfg.token_dropout > 0 or cfg.backproj_dropout > 0):
        raise NotImplementedError("precision='fp32' runs without dropout (eval mode)")
    dev = feat.device
    B, C, Hf, Wf = feat.shape
    p = cfg.patch
    gh, gw = Hf // p, Wf // p
    n = gh * gw
    L, D = lang.shape[1], lang.shape[2]
    S = n + L
    H = cfg.num_heads
    d = D // H
    if d % 8:
        raise NotImplementedError("precision='fp32': head_dim must be a multiple of 8")
    nl = cfg.num_layers
    K = C * p * p
    wpe, img_kind, lang_kind, pos_table = params[:N_HEAD_PARAMS]
    layer_params = [params[N_HEAD_PARAMS + i * N_LAYER_PARAMS: N_HEAD_PARAMS + (i + 1) * N_LAYER_PARAMS] for i in range(nl)]
    lnf_w, lnf_b, wbp, bbp = params[N_HEAD_PARAMS + nl * N_LAYER_PARAMS:][:N_TAIL_PARAMS]

    # token matrix (layout permutation of nn.Conv2d(k = stride = p)'s im2col, utils.py:35-39)
    tok = feat.float().reshape(B, C, gh, p, gw, p).permute(0, 2, 4, 1, 3, 5).reshape(B * n, K).contiguous()
    z = torch.empty(B, S, D, device=dev, dtype=torch.float32)
    z[:, n:] = lang.float() + lang_kind.reshape(D).float()         # cross_f_box_layers.py:76 (one broadcast add: plumbing-sized)
    z2 = z.view(B * S, D)
    # patch embedding accumulated onto the sin1d positions + image kind embedding (a constant [n, D] table, utils.py:209-214)
    emb = (pos_table[:n].float() + img_kind.reshape(D).float())[None].expand(B, n, D).reshape(B * n, D).contiguous()
    z[:, :n] = _gemm32(tok, wpe.reshape(D, K).float(), into=emb).view(B, n, D)
    kpm = None
    if key_pad is not None:
        kpm = torch.zeros(B, S, device=dev, dtype=torch.uint8)
        kpm[:, n:] = key_pad.to(torch.uint8)
    x = z2
    M = B * S
    for (in_w, in_b, out_w, out_b, w1, b1, w2, b2, n1w, n1b, n2w, n2b) in layer_params:
        qkv = _gemm32(x, in_w.float(), bias=in_b.float())
        att = _attention32(qkv, B, S, H, d, kpm)
        y1 = _gemm32(att, out_w.float(), bias=out_b.float(), into=x.clone())                 # x + out_proj(att)
        x1, _, _ = ops.rowln_fwd(y1, n1w.float(), n1b.float(), LN_EPS)
        u = _gemm32(x1, w1.float(), bias=b1.float())
        y2 = _gemm32(u, w2.float(), bias=b2.float(), act_in=1, into=x1.clone())              # x1 + linear2(gelu(linear1(x1)))
        x, _, _ = ops.rowln_fwd(y2, n2w.float(), n2b.float(), LN_EPS)
    xs = x.view(B, S, D)
    vis, _, _ = ops.rowln_fwd(xs[:, :n].reshape(B * n, D).contiguous(), lnf_w.float(), lnf_b.float(), LN_EPS)
    yb = _gemm32(vis, wbp.float(), bias=bbp.float())
    fused = yb.view(B, gh, gw, C, p, p).permute(0, 3, 1, 4, 2, 5).reshape(B, C, Hf, Wf).contiguous()   # F.fold (utils.py:42-46)
    lang_out = xs[:, n:].contiguous()
    return fused, lang_out
