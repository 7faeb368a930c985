"""CPU, build container only: the oracle restatement against the UNMODIFIED reference module
imported from /root/reference (skipped where the reference tree is absent, e.g. the GPU box)."""
import pytest
import torch

from oracle import ref_loader, ref_math
from tests.golden_utils import rel_fro

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


def test_config1_c5_forward_matches_reference():
    """BASELINE.json configs[0] shape family (C5 level, D=896, L=64) at B=2 to stay fast."""
    D, B, L = 896, 2, 64
    cfg = ref_loader.build_fusion_cfg(D, n_levels=1, num_layers=[4], patch=[1], dropout=0.0)
    m = ref_loader.build_reference_module(cfg, [(24, 32)], [2048], seed=0)
    m.train()
    torch.manual_seed(0)
    feats = {"0": torch.relu(torch.randn(B, 2048, 24, 32))}
    lang = torch.randn(B, L, D)
    mask = torch.zeros(B, L, dtype=torch.int64)
    mask[0, :] = 1
    mask[1, :48] = 1
    with torch.no_grad():
        ref, _ = ref_loader.run_reference(m, feats, lang, mask)
        sd = {k: v for k, v in m.state_dict().items()}
        ours, _ = ref_math.cross_fusion_forward(feats, lang, mask, sd, [1], 4, [4])
    assert ref["0"].shape == (B, 2048, 24, 32)
    assert rel_fro(ours["0"], ref["0"]) < 1e-5


def test_state_dict_keys_match_survey_appendix_c():
    cfg = ref_loader.build_fusion_cfg(64, n_levels=2, num_layers=[1, 1], patch=[2, 1], dropout=0.0)
    m = ref_loader.build_reference_module(cfg, [(8, 8), (4, 4)], [8, 16], seed=0)
    keys = set(m.state_dict().keys())
    assert "patches_to_token.0.weight" in keys
    assert "cross_fusion_encoders.1.pos_embedding_layer.pos_embedding" in keys
    assert "cross_fusion_encoders.0.t_encoder.layers.0.self_attn.in_proj_weight" in keys
    assert "tokens_to_features.1.linear.bias" in keys
