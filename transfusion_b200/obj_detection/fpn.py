"""SURVEY 8f N1 — the FPN that consumes the fused feature maps, with its lateral (inner 1x1) convolutions already applied.

The reference runs torchvision's ``FeaturePyramidNetwork`` on the fused maps (``rcnn_apply_fpn``,
modeling/obj_detection/faster_rcnn_wrapper.py:419-421).  When ``CrossFusionBoxWrapper.fuse_fpn_inner(fpn)`` is active the
fusion levels emit the laterals themselves (the 1x1 conv is folded into the back-projection GEMM,
cross_fusion/level_fn.py), and this function performs the remainder of ``FeaturePyramidNetwork.forward``: top-down nearest
upsampling + addition, the 3x3 ``layer_blocks`` and the ``extra_blocks`` (cuDNN convolutions through torch: they are the
detector's own layers, not part of the fusion path)."""
from __future__ import annotations

from collections import OrderedDict

import torch.nn.functional as F


def fpn_from_laterals(fpn, laterals):
    """laterals: ordered mapping level name -> [B, Co, h, w] (finest first), i.e. ``inner_blocks[i](x[i])`` for every i."""
    names = list(laterals.keys())
    x = [laterals[k] for k in names]
    x = [t.float() if t.dtype != fpn.layer_blocks[0][0].weight.dtype else t for t in x]
    last_inner = x[-1]
    results = [fpn.get_result_from_layer_blocks(last_inner, -1)]
    for idx in range(len(x) - 2, -1, -1):
        inner_lateral = x[idx]
        inner_top_down = F.interpolate(last_inner, size=inner_lateral.shape[-2:], mode="nearest")
        last_inner = inner_lateral + inner_top_down
        results.insert(0, fpn.get_result_from_layer_blocks(last_inner, idx))
    if fpn.extra_blocks is not None:
        results, names = fpn.extra_blocks(results, x, names)
    return OrderedDict((k, v) for k, v in zip(names, results))
