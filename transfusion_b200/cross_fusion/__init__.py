"""Host-side mirror of the reference package ``modeling/cross_fusion`` (same class names,
constructor / forward signatures and state_dict keys) whose compute is the C-ABI CUDA library."""
from .cross_f_box_layers import CrossTransformerModuleBox
from .cross_f_box_wrapper import CrossFusionBoxWrapper
from .utils import PositionalEmbeddingLayer, RegroupPatchesLayerBox

__all__ = ["CrossFusionBoxWrapper", "CrossTransformerModuleBox", "PositionalEmbeddingLayer", "RegroupPatchesLayerBox"]
