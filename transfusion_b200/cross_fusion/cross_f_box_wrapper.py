"""Drop-in mirror of modeling/cross_fusion/ego_fusion/cross_f_box_wrapper.py:41-303
(``CrossFusionBoxWrapper``): same constructor arguments, ``forward(x, targets)`` contract, helper
methods and state_dict keys, so ``modeling/model_factory.get_fusion_model`` and
the ``runner/nao`` trainer pick it up unchanged.

Per FPN level the reference runs Conv2d patch-embed -> CrossTransformerModuleBox -> RegroupPatchesLayerBox
as ~60 ATen calls; here each level is ONE autograd node (FusionLevelFunction) that schedules the
hand-written sm_100a kernels of libxfusion_sm100a.so.  There is no PyTorch fallback."""
from __future__ import annotations

import os as _os

import torch
from torch import nn

from .cross_f_box_layers import CrossTransformerModuleBox
from .level_fn import FusionLevelFunction, LevelConfig
from .lm_layers import get_lm_layer
from .utils import PositionalEmbeddingLayer, RegroupPatchesLayerBox, get_visual_token_mask

MAX_NUM_PATCHES = 8192  # cross_f_box_wrapper.py:21
LEVEL_STREAMS = bool(int(_os.environ.get("XF_LEVEL_STREAMS", "1")))   # run independent FPN levels on side streams
SM_PARTITION = bool(int(_os.environ.get("XF_SM_PARTITION", "0")))     # give each concurrent level a fixed share of the SMs


def _default_pooling_factory(narr_embed_args, cross_layer_args):
    """The language-context producer (SBert/MiniLM etc., narr_pooling_layers.py) is outside this path.
    Inside the reference tree its own factory is used; elsewhere the caller injects a module."""
    try:
        from modeling.narration_embeds.narr_pooling_layers import get_narr_pooling_layer  # type: ignore
    except Exception as e:  # not running inside the reference tree
        raise RuntimeError(
            "narr_pooling_layer: pass `narr_pooling_layer=<module>` to CrossFusionBoxWrapper when the "
            "reference's modeling.narration_embeds package is not importable") from e
    return get_narr_pooling_layer(narr_embed_args["text_pooling"])(narr_embed_args, cross_layer_args["narr_out_mode"])


class CrossFusionBoxWrapper(nn.Module):
    def __init__(self, rcnn_model, cross_layer_args=None, narr_embed_args=None, criterion=None, narr_pooling_layer=None, precision=None):
        """precision: None / "bf16" = the tensor-core path (bf16 operands, fp32 accumulate; training and inference);
        "fp32" (or XF_PRECISION=fp32) = the forward-only fp32-tolerance mode of cross_fusion/level_fp32.py (3-way bf16 split
        GEMMs, ~1e-5 relative to the fp32 reference; the reference's Ego4Dv2 config runs precision 32)."""
        super().__init__()
        if cross_layer_args is None or narr_embed_args is None:
            # the reference's keyword defaults (cross_f_wrapper.py:16-54: type "asymmetric", narr_out_mode "embedding", no
            # fpn_features) select a variant its own box wrapper cannot construct (KeyError at :60 there);
            # same positional / keyword signature here, but an explicit message instead
            raise ValueError("CrossFusionBoxWrapper: pass cross_layer_args (the merged fusion YAML, run_experiment.py:75-77,100) and "
                             "narr_embed_args; the reference's defaults describe the unconstructible asymmetric variant")
        self.precision = precision or _os.environ.get("XF_PRECISION", "bf16")
        if self.precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.rcnn_model = rcnn_model
        self.narr_embed_args = narr_embed_args
        if "final_ln" in cross_layer_args["args"]:  # compatibility shim, reference :49-52
            w_ln = cross_layer_args["args"].pop("final_ln")
            cross_layer_args["args"]["final_norm"] = "ln" if w_ln else False
        self.cross_encoder_args = cross_layer_args
        self.forward_language_f = self.cross_encoder_args.get("forward_language_f", False)
        self.vis_mask_type = self.cross_encoder_args.get("vis_mask_type", "global")
        if self.cross_encoder_args.get("type", "cross_transformer") != "cross_transformer":
            raise NotImplementedError("only type: cross_transformer (the shipped, constructible variant) is implemented")
        if self.cross_encoder_args.get("narr_out_mode", "tokens") != "tokens":
            raise NotImplementedError("narr_out_mode 'embedding' is dead code in the reference box model")
        if self.cross_encoder_args.get("pos_embedding", "sin1d") != "sin1d":
            raise NotImplementedError("only the shipped sin1d positional embedding is implemented")
        pn = self.cross_encoder_args.get("patch_norm", {}) or {}
        if pn.get("visual") or pn.get("language"):
            raise NotImplementedError("patch_norm is null in the shipped config")

        self.dsampled_shapes = rcnn_model.get_dsampled_shapes()
        self.in_rgb_channels = rcnn_model.get_features_out_channels()
        self.fpn_features_idx = self.cross_encoder_args["fpn_features"][: len(self.dsampled_shapes)]
        self.vis_input_key = "image"
        self.token_dim = self.cross_encoder_args["args"]["input_f_size"]

        if narr_pooling_layer is not None:
            self.narr_pooling_layer = narr_pooling_layer
        else:
            self.narr_pooling_layer = _default_pooling_factory(narr_embed_args, cross_layer_args)

        self.cross_fusion_encoders = nn.ModuleList(self.setup_cross_fusion_encoders(self.cross_encoder_args))
        self.patches_to_token = nn.ModuleList(self.setup_patches_to_token())
        self.tokens_to_features = nn.ModuleList(self.setup_token_to_features_layers())

        self.criterion = criterion if criterion is not None else {}
        if self.criterion.get("lm", None):
            self.lm_layer = get_lm_layer(self)
        self.lm_on = self.criterion.get("lm", False)
        self.use_lm_f = self.cross_encoder_args["lm_args"].get("use_lm_f", False)
        self.multi_lm = self.cross_encoder_args["lm_args"].get("multi", False) and self.lm_on and not self.use_lm_f
        self._step = 0

    # ---- construction (reference :83-163, 266-294) ------------------------------------------
    def setup_cross_fusion_encoders(self, cross_encoder_args):
        encoders = []
        all_num_layers = cross_encoder_args["args"].pop("num_layers")
        if not isinstance(all_num_layers, list):
            all_num_layers = [all_num_layers] * len(self.dsampled_shapes)
        for i in range(len(self.dsampled_shapes)):
            pos = PositionalEmbeddingLayer(cross_encoder_args["pos_embedding"], MAX_NUM_PATCHES, self.token_dim)
            encoders.append(CrossTransformerModuleBox(no_patches=MAX_NUM_PATCHES, pos_embedding_layer=pos,
                                                      lang_pos_embedding=None, num_layers=all_num_layers[i],
                                                      **cross_encoder_args["args"]))
        return encoders

    def setup_token_to_features_layers(self):
        layers = []
        for i, shape in enumerate(self.dsampled_shapes):
            layers.append(RegroupPatchesLayerBox(self.token_dim, shape[0], shape[1], self.cross_encoder_args["patch_h"][i],
                                                 self.cross_encoder_args["patch_w"][i], self.in_rgb_channels[i],
                                                 self.cross_encoder_args["backproj_dropout"],
                                                 self.cross_encoder_args.get("backproj_activ_f", None)))
        return layers

    def setup_patches_to_token(self):
        mods = []
        for i, _ in enumerate(self.dsampled_shapes):
            ph, pw = self.cross_encoder_args["patch_h"][i], self.cross_encoder_args["patch_w"][i]
            mods.append(self.setup_patch_to_token(self.cross_encoder_args["patch_norm"], self.in_rgb_channels[i] * ph * pw,
                                                  self.token_dim, in_channels=self.in_rgb_channels[i], patch_h=ph, patch_w=pw))
        return mods

    def setup_patch_to_token(self, patch_norm, patch_dim, token_dim, in_channels=None, patch_h=None, patch_w=None):
        if in_channels is None:
            raise NotImplementedError("linear patch embedding without channels is not used by the box model")
        return nn.Conv2d(in_channels=in_channels, out_channels=token_dim, kernel_size=(patch_h, patch_w),
                         stride=(patch_h, patch_w), bias=False)

    # ---- FPN laterals folded into the back-projection ---------------------------------
    def fuse_fpn_inner(self, fpn):
        """`fpn`: the torchvision FeaturePyramidNetwork the fused maps feed (rcnn_to_wrap.backbone.fpn,
        faster_rcnn_wrapper.py:419-421).  After this call every level computes its FPN lateral (inner 1x1 conv) inside
        the back-projection GEMM, the [B, C, h, w] fused maps are never materialised, and `forward` runs the rest of the
        FPN (top-down additions, 3x3 output convs, extra blocks) on the laterals instead of calling
        `rcnn_model.apply_fpn`.  The FPN keeps owning its parameters (not re-registered here: state_dict unchanged).
        Pass None to undo."""
        if fpn is not None:
            if len(fpn.inner_blocks) != len(self.fpn_features_idx):
                raise ValueError("fuse_fpn_inner: the FPN must have one inner block per fused level")
            for blk in fpn.inner_blocks:
                conv = blk[0] if isinstance(blk, nn.Sequential) else blk
                if (isinstance(blk, nn.Sequential) and len(blk) != 1) or not isinstance(conv, nn.Conv2d) or \
                        conv.kernel_size != (1, 1) or conv.stride != (1, 1) or conv.groups != 1 or conv.bias is None:
                    raise NotImplementedError("fuse_fpn_inner: inner blocks must be plain 1x1 convolutions with bias")
        self.__dict__["_xf_fpn"] = fpn

    # ---- cached bf16 weight copies (cross_fusion/level_fn.py) ---------------------------------
    def invalidate_weight_cache(self):
        """Drops the cached bf16 weight copies; call after modifying parameters behind autograd's back (`p.data.<op>_()`)
        between two inference calls.  train() / eval() switches do it automatically."""
        from ..weight_cache import invalidate
        invalidate(self)

    def train(self, mode: bool = True):
        if mode != self.training:
            self.invalidate_weight_cache()
        return super().train(mode)

    # ---- the hot path -----------------------------------------------------------------------
    def run_level(self, i: int, feat: torch.Tensor, language_f: torch.Tensor, lang_pad_mask, need_lang_out=False, out_stream=None,
                  gemm_ctas: int = 0, lateral_conv=None):
        """One FPN level: reference :180-212.  Returns (fused [B,C,h,w], fused language tokens or None)."""
        enc: CrossTransformerModuleBox = self.cross_fusion_encoders[i]
        t2f: RegroupPatchesLayerBox = self.tokens_to_features[i]
        pe = self.patches_to_token[i]
        get_visual_token_mask(None, self.vis_mask_type)
        p = t2f.patch_h
        n = (feat.shape[2] // p) * (feat.shape[3] // p)
        if n > MAX_NUM_PATCHES:
            raise ValueError(f"{n} visual tokens exceed MAX_NUM_PATCHES={MAX_NUM_PATCHES}")
        if feat.shape[2] % p or feat.shape[3] % p:
            raise ValueError("feature map size must be divisible by the patch size")
        seed = 0
        if self.training:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        # the dropout seed of the last forward per level: with the site ids of LevelConfig.stream() it reproduces every
        # keep mask of the step (xf_debug_dropout_mask; tests/test_gpu_dropout_parity.py)
        self.__dict__.setdefault("_xf_last_seeds", {})[i] = seed
        cfg = LevelConfig(level=i, patch=p, num_heads=enc.num_heads, num_layers=enc.num_layers, training=self.training,
                          patch_dropout=float(enc.patch_dropout), token_dropout=float(enc.token_dropout),
                          backproj_dropout=float(t2f.back_dropout.p), seed=seed, need_lang_out=need_lang_out,
                          out_stream=out_stream, gemm_ctas=gemm_ctas, lateral=lateral_conv is not None)
        params = [pe.weight, enc.image_kind_embedding, enc.lang_kind_embedding, enc.pos_embedding_layer.table(),
                  *enc.level_params(), enc.final_norm_layer.weight, enc.final_norm_layer.bias, t2f.linear.weight,
                  t2f.linear.bias]
        if self.precision == "fp32":
            if lateral_conv is not None:
                raise NotImplementedError("precision='fp32' does not combine with fuse_fpn_inner")
            from .level_fp32 import fusion_level_forward_fp32
            fused, lang_out = fusion_level_forward_fp32(cfg, feat, language_f, lang_pad_mask, *params)
            return fused.to(feat.dtype) if feat.dtype in (torch.float32, torch.bfloat16) else fused, (lang_out if need_lang_out else None)
        if lateral_conv is not None:
            params += [lateral_conv.weight, lateral_conv.bias]
        fused, lang_out = FusionLevelFunction.apply(cfg, feat, language_f, lang_pad_mask, *params)
        return fused, (lang_out if need_lang_out else None)

    def _sm_shares(self, feats, language_f):
        """Spatial partition of the SMs between the concurrently running FPN levels: each level's persistent GEMM grids get a
        share proportional to the level's algorithmic FLOPs (SURVEY 8d formula), so the four levels advance side by side and
        finish together instead of queueing full-machine grids behind one another (the coarse levels fill only 2.3 - 4.6
        waves of a 148-SM grid).  XF_SM_SHARES="a,b,c,d" overrides (CTAs per level, 0 = whole machine); "off" disables."""
        env = _os.environ.get("XF_SM_SHARES", "")
        n_lv = len(self.fpn_features_idx)
        if env == "off":
            return None
        if env:
            v = [int(x) for x in env.split(",")]
            return (v + [0] * n_lv)[:n_lv]
        if not SM_PARTITION:
            return None
        D, L = self.token_dim, language_f.shape[1]
        work = []
        for i, key in enumerate(self.fpn_features_idx):
            f = feats[str(key)]
            p = self.tokens_to_features[i].patch_h
            n = (f.shape[2] // p) * (f.shape[3] // p)
            S = n + L
            nl = self.cross_fusion_encoders[i].num_layers
            work.append(4.0 * n * f.shape[1] * p * p * D + nl * (16.0 * S * D * D + 4.0 * S * S * D))
        total_sms = torch.cuda.get_device_properties(language_f.device).multi_processor_count
        tot = sum(work)
        shares = [max(2, int(round(total_sms * w / tot / 2.0)) * 2) for w in work]   # even: the GEMM runs on CTA pairs
        return shares

    def _level_streams(self, ref):
        """One side stream per FPN level on `ref`'s device (None on CPU tensors: the kernels will raise anyway)."""
        if not (isinstance(ref, torch.Tensor) and ref.is_cuda):
            return None
        cache = self.__dict__.setdefault("_xf_streams", {})
        dev = ref.device.index
        if dev not in cache:
            # XF_STREAM_PRIO="a,b,c,d": CUDA stream priority per level (-1 = high: its thread blocks are placed first when
            # SMs free up); default: all equal
            prio = [int(x) for x in _os.environ.get("XF_STREAM_PRIO", "").split(",") if x.strip()]
            n_lv = len(self.fpn_features_idx)
            prio = (prio + [0] * n_lv)[:n_lv]
            cache[dev] = [torch.cuda.Stream(device=ref.device, priority=prio[i]) for i in range(n_lv)]
            # parameter gradients are produced on the side streams and accumulated on the caller's stream: the
            # engine inserts the synchronisation; the mismatch it warns about is intended
            quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
            if quiet is not None:
                quiet(False)
        return cache[dev]

    def forward(self, x, targets=None):
        visual_data = x[self.vis_input_key]
        features_dict = self.rcnn_model.forward_features(visual_data, targets)
        language_f, att_w, att_mask = self.narr_pooling_layer(x["language_f"], pad_mask=True)
        lang_pad = None if att_mask is None else ~(att_mask.type(torch.bool))  # reference :196
        lm_from_fused = bool(self.lm_on and not self.use_lm_f)
        need_all_lang = bool(self.multi_lm or self.forward_language_f)
        mscale_l_features = []
        fused_l_features = None
        lm_tokens = None   # what the reference's `fused_l_features` holds after its loop: the LAST level's tokens (:177-222)
        last_level = len(self.fpn_features_idx) - 1
        # Levels are independent given the shared language input (forward_language_f: False, fusion yml :27),
        # so they are visited coarsest-first: autograd then runs the largest level's backward FIRST and its
        # gradient all-reduce overlaps the remaining levels.  With forward_language_f the reference order holds.
        level_order = list(enumerate(self.fpn_features_idx))
        if not self.forward_language_f and not self.multi_lm:
            level_order = level_order[::-1]
        # Independent levels run on their own CUDA streams: the persistent GEMM / attention kernels of one level
        # fill the tail waves of another (the three coarse levels have 2.3 - 4.6 waves per GEMM), and autograd
        # replays each level's backward on the stream its forward ran on.
        side = self._level_streams(language_f) if (LEVEL_STREAMS and len(level_order) > 1 and not self.forward_language_f
                                                   and not self.multi_lm) else None
        cur = torch.cuda.current_stream() if side is not None else None
        keep_alive = []
        if side is not None:
            # Host run-ahead limiter.  Tensors that cross streams (the fused maps) return to the caching allocator
            # only when the consuming stream has passed the point of the free, so how many blocks the allocator needs
            # depends on how far the host runs ahead of the device; an unthrottled host kept provoking cudaMalloc
            # calls in steady state, and cudaMalloc synchronises the device (observed: single steps of 60-120 ms).
            # Training: wait for the previous forward to have finished on the device (its backward is still queued,
            # so the device never idles); inference: stay at most two forwards ahead.
            pending = self.__dict__.setdefault("_xf_fwd_events", [])
            depth = 1 if (self.training and torch.is_grad_enabled()) else 2
            while len(pending) >= depth:
                pending.pop(0).synchronize()
        shares = self._sm_shares(features_dict["features"], language_f) if side is not None else None
        fpn = self.__dict__.get("_xf_fpn")

        def lat_conv(i):
            if fpn is None:
                return None
            blk = fpn.inner_blocks[i]
            return blk[0] if isinstance(blk, nn.Sequential) else blk

        for i, key in level_order:
            key = str(key)
            feat = features_dict["features"][key]
            need_lang_out = need_all_lang or (lm_from_fused and i == last_level)
            self.tokens_to_features[i].init_h = feat.shape[2]
            self.tokens_to_features[i].init_w = feat.shape[3]
            if side is not None:
                st = side[i % len(side)]
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    fused, fused_l_features = self.run_level(i, feat, language_f, lang_pad, need_lang_out, out_stream=cur,
                                                             gemm_ctas=shares[i] if shares else 0, lateral_conv=lat_conv(i))
                # Memory safety across streams without record_stream (whose deferred frees made the allocator's
                # demand depend on host run-ahead): the outputs come from the caller stream's pool (out_stream) and
                # are written on `st`, which waited for everything the caller had enqueued; the inputs, allocated on
                # the caller's stream and read on `st`, are kept alive until the caller's stream has joined `st`.
                keep_alive.append((feat, language_f, lang_pad))
            else:
                fused, fused_l_features = self.run_level(i, feat, language_f, lang_pad, need_lang_out, lateral_conv=lat_conv(i))
            if i == last_level:
                lm_tokens = fused_l_features   # independent of the visiting order
            if self.multi_lm:
                mscale_l_features.append(fused_l_features)
            if self.forward_language_f:
                if self.forward_language_f == "direct":
                    language_f = fused_l_features
                elif self.forward_language_f == "sum":
                    language_f = language_f + fused_l_features
                else:
                    raise NotImplementedError()
            features_dict["features"][key] = fused
        if side is not None:
            for st in side:
                cur.wait_stream(st)
            ev = torch.cuda.Event()
            ev.record(cur)
            self.__dict__["_xf_fwd_events"].append(ev)
            keep_alive.clear()   # the join is enqueued: later frees are ordered after the side streams' reads
        if fpn is not None:
            # features_dict["features"][key] now holds the LATERALS: finish the FPN here (reference: rcnn_model.apply_fpn ->
            # backbone.fpn(features), faster_rcnn_wrapper.py:171-174,419-421)
            from ..obj_detection.fpn import fpn_from_laterals
            features_dict["features"] = fpn_from_laterals(fpn, features_dict["features"])
        else:
            features_dict = self.rcnn_model.apply_fpn(features_dict)
        if "hand_boxes" in x:
            features_dict["hand_boxes"] = x["hand_boxes"]
        if "hand_poses" in x:
            features_dict["hand_poses"] = x["hand_poses"]
        rcnn_outs = self.rcnn_model.apply_rpn_roi_on_features(features_dict)
        if self.lm_on:
            rcnn_outs["lm"] = self.lm_layer(
                mscale_l_features if self.multi_lm else lm_tokens if not self.use_lm_f else language_f,
                None if att_mask is None else att_mask.type(torch.bool))
        return rcnn_outs

    # ---- delegations (reference :232-264) ---------------------------------------------------
    def call_model_epoch_triggers(self, epoch):
        if epoch >= self.narr_embed_args["train_ep"] and self.narr_embed_args["train_ep"] != -1:
            self.narr_pooling_layer.unfreeze_embeddings()
        self.rcnn_model.call_model_epoch_triggers(epoch)

    def dets_from_outs(self, outs, orig_img_shapes=None, targets=None, hand_poses=None, hand_boxes=None):
        return self.rcnn_model.dets_from_outs(outs, orig_img_shapes, targets=targets, hand_poses=hand_poses,
                                              hand_boxes=hand_boxes)

    def forward_w_dets(self, x, targets=None):
        outs = self(x, targets)
        original_image_shapes = [tuple(img.shape[1:]) for img in x["image"]]
        return self.dets_from_outs(outs, orig_img_shapes=original_image_shapes, targets=targets,
                                   hand_poses=x.get("hand_poses"), hand_boxes=x.get("hand_boxes"))

    def postprocess_detections(self, detections, proposals, image_sizes, original_image_shapes):
        return self.rcnn_model.postprocess_detections(detections, proposals, image_sizes, original_image_shapes)

    def compute_rpn_loss(self, objectness, pred_bbox_deltas, labels, regression_targets):
        return self.rcnn_model.compute_rpn_loss(objectness, pred_bbox_deltas, labels, regression_targets)
