"""Downstream consumers of the fused feature maps that run on the same C-ABI kernels."""
from .box_head import FusedBoxHead  # noqa: F401
