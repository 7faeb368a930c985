"""SURVEY 8f N1 — the dense part of the RoI box head on the path's tensor-core GEMM.

Mirror of what ``RoIHeadsWrapper.forward`` (modeling/obj_detection/roi_wrappers.py:198-214) runs after RoIAlign:
``box_head`` = torchvision ``TwoMLPHead(256*7*7, R)`` (fc6 + ReLU, fc7 + ReLU), then ``box_regressor`` =
``Sequential(box_dropout, Linear(R, 4*nouns))`` (faster_rcnn_wrapper.py:93), ``noun_classifier = Linear(R, nouns)`` and
``verb_classifier = Linear(R, verbs)``; R = 1024 (v1) / 1280 (v2), ΣR_i = 128 boxes per image in training.  The three
dropouts around it are p = 0 in both shipped configs (ego_vis_det_ego4dv2.yml:4-5, ego_nao_res50_ego4dv2.yml:137);
a non-zero one raises instead of being silently skipped.

The nn.Linear modules stay the parameter containers (same names, same init, checkpoints load unchanged); the math is ONE
autograd node over ``xf_gemm``: bias + ReLU in the GEMM epilogue, the ReLU mask of the backward in the dgrad epilogue, the
three predictor Linears as one concatenated GEMM with fp32 logits, weight gradients by split-K fp32 reduction.
No PyTorch fallback: CPU tensors raise."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import ops
from ..weight_cache import bf16_weight, invalidate


def _split_k(tokens: int) -> int:
    return max(1, min(8, tokens // 512))


class _BoxHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w6, b6, w7, b7, wr, br, wn, bn, wv, bv, training):
        if not x.is_cuda:
            raise RuntimeError("transfusion_b200: FusedBoxHead has no CPU implementation (CUDA tensors required)")
        dev, bf = x.device, torch.bfloat16
        R_, K = x.shape
        Rp = w6.shape[0]
        need_grad = any(ctx.needs_input_grad)
        heads = [(wr, br)] + ([(wn, bn)] if wn is not None else []) + ([(wv, bv)] if wv is not None else [])
        sizes = [w.shape[0] for w, _ in heads]
        ncat = sum(sizes)
        ncat_p = (ncat + 7) // 8 * 8
        casts = []
        trust = not (training and need_grad)
        w6_b = bf16_weight(w6, Rp, K, casts, trust, dev)
        w7_b = bf16_weight(w7, Rp, Rp, casts, trust, dev)
        # the predictor weights live in ONE [ncat_p, R] bf16 matrix (rows beyond ncat are zero): a single GEMM, fp32 logits
        wcat_b = torch.zeros(ncat_p, Rp, device=dev, dtype=bf) if ncat_p != ncat else torch.empty(ncat_p, Rp, device=dev, dtype=bf)
        bcat = torch.zeros(ncat_p, device=dev, dtype=torch.float32)
        off = 0
        for (w, b), n in zip(heads, sizes):
            casts.append((w, wcat_b[off:off + n], n, Rp, 0, 0, 0, 0))
            bcat[off:off + n] = b.detach()
            off += n
        # the activations entering fc6 (fp32 RoIAlign output) -> bf16 through the same multi-tensor cast launch
        if x.dtype == bf:
            x_b = x.contiguous()
        else:
            x_b = torch.empty(R_, K, device=dev, dtype=bf)
            casts.append((x.detach().float().contiguous(), x_b, R_, K, 0, 0, 0, 0))
        ops.cast_pad_multi(casts)
        h1 = torch.empty(R_, Rp, device=dev, dtype=bf)
        ops.gemm(x_b, w6_b, h1, M=R_, N=Rp, K=K, bias=b6.detach(), act=2)          # fc6 + ReLU
        h2 = torch.empty(R_, Rp, device=dev, dtype=bf)
        ops.gemm(h1, w7_b, h2, M=R_, N=Rp, K=Rp, bias=b7.detach(), act=2)          # fc7 + ReLU
        logits = torch.empty(R_, ncat_p, device=dev, dtype=torch.float32)
        ops.gemm(h2, wcat_b, logits, M=R_, N=ncat_p, K=Rp, bias=bcat)              # box_regressor | noun | verb
        if need_grad:
            ctx.save_for_backward(x_b, h1, h2, w6_b, w7_b, wcat_b)
            ctx.meta = (sizes, ncat, ncat_p, x.dtype, wn is not None, wv is not None)
        outs, off = [], 0
        for n in sizes:
            outs.append(logits[:, off:off + n].contiguous())
            off += n
        box = outs[0]
        noun = outs[1] if wn is not None else None
        verb = outs[-1] if wv is not None else None
        return box, noun, verb

    @staticmethod
    def backward(ctx, d_box, d_noun, d_verb):
        x_b, h1, h2, w6_b, w7_b, wcat_b = ctx.saved_tensors
        sizes, ncat, ncat_p, x_dtype, has_n, has_v = ctx.meta
        dev, bf, f32 = x_b.device, torch.bfloat16, torch.float32
        R_, K = x_b.shape
        Rp = h1.shape[1]
        # d(logits) of the concatenated predictor GEMM, bf16 (zero where a head received no gradient / in the pad columns)
        dcat = torch.zeros(R_, ncat_p, device=dev, dtype=bf)
        parts = [d_box] + ([d_noun] if has_n else []) + ([d_verb] if has_v else [])
        off = 0
        for g, n in zip(parts, sizes):
            if g is not None:
                dcat[:, off:off + n] = g
            off += n
        sk = _split_k(R_)
        g_wcat = torch.zeros(ncat_p, Rp, device=dev, dtype=f32)
        ops.gemm(dcat, h2, g_wcat, M=ncat_p, N=Rp, K=R_, a_mn_major=True, b_mn_major=True, accumulate=True, split_k=sk)
        g_bcat = torch.zeros(ncat_p, device=dev, dtype=f32)
        ops.colsum(dcat, g_bcat, R_, ncat_p)
        dpre7 = torch.empty(R_, Rp, device=dev, dtype=bf)
        ops.gemm(dcat, wcat_b, dpre7, M=R_, N=Rp, K=ncat_p, b_mn_major=True, dact_in=h2, act=2)   # (dlogits Wcat) o relu'(h2)
        g_w7 = torch.zeros(Rp, Rp, device=dev, dtype=f32)
        ops.gemm(dpre7, h1, g_w7, M=Rp, N=Rp, K=R_, a_mn_major=True, b_mn_major=True, accumulate=True, split_k=sk)
        g_b7 = torch.zeros(Rp, device=dev, dtype=f32)
        ops.colsum(dpre7, g_b7, R_, Rp)
        dpre6 = torch.empty(R_, Rp, device=dev, dtype=bf)
        ops.gemm(dpre7, w7_b, dpre6, M=R_, N=Rp, K=Rp, b_mn_major=True, dact_in=h1, act=2)
        g_w6 = torch.zeros(Rp, K, device=dev, dtype=f32)
        ops.gemm(dpre6, x_b, g_w6, M=Rp, N=K, K=R_, a_mn_major=True, b_mn_major=True, accumulate=True, split_k=sk)
        g_b6 = torch.zeros(Rp, device=dev, dtype=f32)
        ops.colsum(dpre6, g_b6, R_, Rp)
        dx = None
        if ctx.needs_input_grad[0]:
            dxb = torch.empty(R_, K, device=dev, dtype=bf)
            ops.gemm(dpre6, w6_b, dxb, M=R_, N=K, K=Rp, b_mn_major=True)
            dx = dxb.to(x_dtype)
        gw, gb, off = [], [], 0
        for n in sizes:
            gw.append(g_wcat[off:off + n])
            gb.append(g_bcat[off:off + n])
            off += n
        i = 1
        g_wn = g_bn = g_wv = g_bv = None
        if has_n:
            g_wn, g_bn = gw[i], gb[i]
            i += 1
        if has_v:
            g_wv, g_bv = gw[i], gb[i]
        return dx, g_w6, g_b6, g_w7, g_b7, gw[0], gb[0], g_wn, g_bn, g_wv, g_bv, None


def _p_of(drop) -> float:
    return float(getattr(drop, "p", 0.0)) if drop is not None and not isinstance(drop, nn.Identity) else 0.0


class FusedBoxHead(nn.Module):
    """box_head (fc6, fc7) + box_regressor + noun_classifier + verb_classifier of the reference's RoI heads on xf_gemm.

    ``FusedBoxHead.from_roi_heads(roi_heads_wrapper)`` shares the reference modules' parameters (drop-in: the wrapper keeps
    owning them, state_dict keys unchanged); ``forward(box_features)`` takes the RoIAlign output ``[ΣR_i, 256, 7, 7]`` (or
    flattened) and returns ``(box_regression [ΣR_i, 4*nouns], class_logits [ΣR_i, nouns], verb_logits [ΣR_i, verbs] | None)``
    in fp32, like roi_wrappers.py:198-214."""

    def __init__(self, in_features: int, representation_size: int, noun_classes: int, verb_classes: Optional[int]):
        super().__init__()
        self.fc6 = nn.Linear(in_features, representation_size)
        self.fc7 = nn.Linear(representation_size, representation_size)
        self.box_regressor = nn.Sequential(nn.Identity(), nn.Linear(representation_size, 4 * noun_classes))
        self.noun_classifier = nn.Linear(representation_size, noun_classes)
        self.verb_classifier = nn.Linear(representation_size, verb_classes) if verb_classes else None
        nn.init.normal_(self.box_regressor[1].weight, std=0.01)   # roi_wrappers.py:90-91,104-105
        nn.init.constant_(self.box_regressor[1].bias, 0)
        nn.init.normal_(self.noun_classifier.weight, std=0.01)
        nn.init.constant_(self.noun_classifier.bias, 0)

    @classmethod
    def from_roi_heads(cls, roi):
        """roi: the reference's RoIHeadsWrapper (attributes roi_head_wrap.box_head, box_regressor, noun_classifier,
        verb_classifier, dropout_1, classif_dropout)."""
        for name in ("dropout_1", "classif_dropout"):
            if _p_of(getattr(roi, name, None)) > 0:
                raise NotImplementedError(f"FusedBoxHead: {name} > 0 is not fused (both shipped configs use 0)")
        if _p_of(roi.box_regressor[0]) > 0:
            raise NotImplementedError("FusedBoxHead: box_2_dropout > 0 is not fused (both shipped configs use 0)")
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        bh = roi.roi_head_wrap.box_head
        self.fc6, self.fc7 = bh.fc6, bh.fc7
        self.box_regressor = roi.box_regressor
        self.noun_classifier = roi.noun_classifier
        self.verb_classifier = roi.verb_classifier
        return self

    def train(self, mode: bool = True):
        if mode != self.training:
            invalidate(self)
        return super().train(mode)

    def forward(self, box_features: torch.Tensor):
        x = box_features.flatten(start_dim=1)
        lin = self.box_regressor[1]
        wn = self.noun_classifier
        wv = self.verb_classifier
        return _BoxHeadFn.apply(x, self.fc6.weight, self.fc6.bias, self.fc7.weight, self.fc7.bias, lin.weight, lin.bias,
                                wn.weight if wn is not None else None, wn.bias if wn is not None else None,
                                wv.weight if wv is not None else None, wv.bias if wv is not None else None, self.training)
