"""CPU: host-side mirror of the reference interface — constructor contract, state_dict keys
(SURVEY Appendix C), config compatibility shims, loud failure without CUDA."""
import copy

import pytest
import torch

from oracle import ref_math
from oracle.ref_loader import FakeRCNN, PassThroughPooling
from transfusion_b200.configs import WORKLOADS, default_fusion_cfg, level_shapes
from transfusion_b200.cross_fusion import CrossFusionBoxWrapper
from transfusion_b200.parallel import level_buckets, shard_range
from tests.golden_utils import load_golden


def _build(D=64, shapes=((8, 8), (4, 4)), channels=(8, 16), patch=(2, 1), layers=(1, 2), lm=False, **kw):
    cfg = default_fusion_cfg(D, n_levels=len(shapes), num_layers=list(layers), patch=list(patch), **kw)
    return CrossFusionBoxWrapper(FakeRCNN(list(shapes), list(channels), 9, 6), cfg, {"text_pooling": "x", "train_ep": -1},
                                 criterion={"lm": 1 if lm else 0}, narr_pooling_layer=PassThroughPooling())


def test_state_dict_keys_match_reference_contract():
    m = _build(lm=True)
    keys = set(m.state_dict().keys())
    for i, nl in enumerate((1, 2)):
        assert f"patches_to_token.{i}.weight" in keys
        for k in ("image_kind_embedding", "lang_kind_embedding", "heatmap_token", "padding_mask",
                  "pos_embedding_layer.pos_embedding", "final_norm_layer.weight", "final_norm_layer.bias"):
            assert f"cross_fusion_encoders.{i}.{k}" in keys
        for l in range(nl):
            pre = f"cross_fusion_encoders.{i}.t_encoder.layers.{l}."
            for k in ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
                      "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
                      "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias"):
                assert pre + k in keys
        assert f"tokens_to_features.{i}.linear.weight" in keys and f"tokens_to_features.{i}.linear.bias" in keys
    for k in ("lm_layer.ln.weight", "lm_layer.mlp_noun.weight", "lm_layer.mlp_verb.bias"):
        assert k in keys
    sd = m.state_dict()
    assert sd["patches_to_token.0.weight"].shape == (64, 8, 2, 2)
    assert sd["tokens_to_features.0.linear.weight"].shape == (8 * 4, 64)
    assert sd["cross_fusion_encoders.0.pos_embedding_layer.pos_embedding"].shape == (1, 8192, 64)


@pytest.mark.parametrize("name", ["fusion4_d32", "c5_d64_lm"])
def test_golden_reference_state_dict_loads(name):
    g = load_golden(name)
    feats = g["features"]
    keys = sorted(feats, key=int)
    m = _build(D=g["lang"].shape[-1], shapes=[tuple(feats[k].shape[2:]) for k in keys],
               channels=[feats[k].shape[1] for k in keys], patch=g["patch"], layers=g["layers"], lm=g["lm_on"])
    missing, unexpected = m.load_state_dict(g["params"], strict=False)
    assert not unexpected
    assert all(("pos_embedding" in k or "padding_mask" in k) for k in missing)


def test_pos_embedding_buffer_is_the_sin1d_table():
    m = _build()
    buf = m.cross_fusion_encoders[0].pos_embedding_layer.pos_embedding
    assert torch.allclose(buf[0, :50], ref_math.sin1d_table(50, 64), atol=1e-6)


def test_ctor_pops_num_layers_and_final_ln_like_the_reference():
    cfg = default_fusion_cfg(64, n_levels=1, num_layers=[1], patch=[1])
    cfg["args"].pop("final_norm")
    cfg["args"]["final_ln"] = True
    m = CrossFusionBoxWrapper(FakeRCNN([(4, 4)], [8]), cfg, {"text_pooling": "x", "train_ep": -1}, criterion={},
                              narr_pooling_layer=PassThroughPooling())
    assert "num_layers" not in cfg["args"] and "final_ln" not in cfg["args"] and cfg["args"]["final_norm"] == "ln"
    assert m.cross_fusion_encoders[0].num_layers == 1


def test_unsupported_dead_variants_raise():
    cfg = default_fusion_cfg(64, n_levels=1, num_layers=[1], patch=[1])
    cfg["type"] = "asymmetric"
    with pytest.raises(NotImplementedError):
        CrossFusionBoxWrapper(FakeRCNN([(4, 4)], [8]), cfg, {"text_pooling": "x", "train_ep": -1}, criterion={},
                              narr_pooling_layer=PassThroughPooling())


def test_cpu_tensors_fail_loudly_no_fallback():
    m = _build()
    m.rcnn_model.features = {"0": torch.randn(1, 8, 8, 8), "1": torch.randn(1, 16, 4, 4)}
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        m({"image": None, "language_f": (torch.randn(1, 4, 64), torch.ones(1, 4, dtype=torch.int64))})


def test_lm_head_fails_loudly_on_cpu_and_keeps_reference_parameter_names():
    """PoolPredictor (lm_layers.py:30-81): same submodule / parameter names as the reference, no CPU path."""
    from transfusion_b200._lib import XfError
    from transfusion_b200.cross_fusion.lm_layers import PoolPredictor
    head = PoolPredictor({"type": "mean", "ln": True, "repr_size": 16}, 32, 5, 3)
    assert sorted(k for k, _ in head.named_parameters()) == sorted([
        "ln.weight", "ln.bias", "repr_mlp.1.weight", "repr_mlp.1.bias", "mlp_noun.weight", "mlp_noun.bias",
        "mlp_verb.weight", "mlp_verb.bias"])
    with pytest.raises(XfError):
        head(torch.randn(2, 4, 32), torch.ones(2, 4, dtype=torch.bool))


def test_workload_shapes_match_survey_appendix_b():
    v2, v1 = WORKLOADS["ego4dv2"], WORKLOADS["ego4dv1"]
    n2 = [(h // p) * (w // p) for (h, w), p in zip(level_shapes(v2), v2["patch"])]
    n1 = [(h // p) * (w // p) for (h, w), p in zip(level_shapes(v1), v1["patch"])]
    assert n2 == [3072, 768, 768, 768] and n1 == [4000, 1000, 1000, 1000]
    m = _build()
    assert len(level_buckets(m)) == 2
    assert shard_range(40, 1, 3) == (13, 26)
