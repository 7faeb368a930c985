"""transfusion_b200 — B200-native (sm_100a) implementation of TransFusion's cross_fusion hot path
behind the reference's own nn.Module interface.  See DESIGN.md."""
__all__ = ["ops"]
