"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

For each case the reference ``CrossFusionBoxWrapper`` (cross_f_box_wrapper.py:41) is
constructed under a fixed seed with all dropout probabilities 0 and ``train()`` mode
(keeps torch >= 1.12 off the nested-tensor fast path), run forward and
backward (loss = sum(out * fixed random cotangent)), and inputs, parameters, outputs and
gradients are frozen to a compressed npz.  The sin1d ``pos_embedding`` buffers are
deterministic functions of (pos, D) and are not stored.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import ref_loader

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # four levels, the shipped patch sizes 4,4,2,1 and 4 layers per level, tiny widths
    "fusion4_d32": dict(D=32, heads=4, image=(64, 96), channels=[8, 16, 32, 64], patch=[4, 4, 2, 1],
                        layers=[4, 4, 4, 4], B=2, L=8, lens=[8, 5], lm=False, seed=11),
    # single C5-like level with the LM head on, ragged language lengths incl. length 1
    "c5_d64_lm": dict(D=64, heads=4, image=(128, 160), channels=[48], patch=[1], strides=[32],
                      layers=[2], B=3, L=12, lens=[12, 1, 7], lm=True, seed=23),
    # head_dim not a multiple of 8 (like Ego4Dv1's 178): D=40, 4 heads -> d=10
    "c4_d40_oddhead": dict(D=40, heads=4, image=(64, 64), channels=[24], patch=[2], strides=[16],
                           layers=[2], B=2, L=6, lens=[3, 6], lm=False, seed=37),
}


def make_inputs(case):
    g = torch.Generator().manual_seed(case["seed"] + 1000)
    H, W = case["image"]
    strides = case.get("strides", [4, 8, 16, 32][: len(case["channels"])])
    feats = {}
    for i, (C, s) in enumerate(zip(case["channels"], strides)):
        feats[str(i)] = torch.relu(torch.randn(case["B"], C, H // s, W // s, generator=g))
    lang = 0.5 * torch.randn(case["B"], case["L"], case["D"], generator=g)
    mask = torch.zeros(case["B"], case["L"], dtype=torch.int64)
    for b, n in enumerate(case["lens"]):
        mask[b, :n] = 1
    cot = {k: torch.randn(v.shape, generator=g) for k, v in feats.items()}
    return feats, lang, mask, cot, strides


def run_case(name, case):
    feats, lang, mask, cot, strides = make_inputs(case)
    H, W = case["image"]
    shapes = [(H // s, W // s) for s in strides]
    cfg = ref_loader.build_fusion_cfg(case["D"], n_levels=len(shapes), num_layers=case["layers"],
                                      num_heads=case["heads"], patch=case["patch"], dropout=0.0)
    m = ref_loader.build_reference_module(cfg, shapes, case["channels"], lm=case["lm"], seed=case["seed"],
                                          noun_classes=9, verb_classes=6)
    m.train()
    feats_in = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    lang_in = lang.clone().requires_grad_(True)
    out, lm = ref_loader.run_reference(m, feats_in, lang_in, mask)
    loss = sum((out[k] * cot[k]).sum() for k in out)
    if lm is not None:
        loss = loss + lm["noun_logits"].sum() * 0.5 + (lm["verb_logits"] ** 2).sum() * 0.25
    loss.backward()

    blob = {}
    for k, v in feats.items():
        blob[f"in.features.{k}"] = v.numpy()
        blob[f"in.cotangent.{k}"] = cot[k].numpy()
        blob[f"out.features.{k}"] = out[k].detach().numpy()
        blob[f"grad.features.{k}"] = feats_in[k].grad.numpy()
    blob["in.language_f"] = lang.numpy()
    blob["in.att_mask"] = mask.numpy()
    blob["grad.language_f"] = lang_in.grad.numpy()
    if lm is not None:
        blob["out.lm.noun_logits"] = lm["noun_logits"].detach().numpy()
        blob["out.lm.verb_logits"] = lm["verb_logits"].detach().numpy()
    for k, p in m.named_parameters():
        if k.startswith("rcnn_model") or k.startswith("narr_pooling_layer"):
            continue
        blob[f"param.{k}"] = p.detach().numpy()
        if p.grad is not None:
            blob[f"pgrad.{k}"] = p.grad.numpy()
    blob["meta.patch"] = np.array(case["patch"])
    blob["meta.layers"] = np.array(case["layers"])
    blob["meta.heads"] = np.array(case["heads"])
    blob["meta.lm"] = np.array(int(case["lm"]))
    blob["meta.torch_version"] = np.array(torch.__version__)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in blob.items()})
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB), loss={float(loss):.6f}")


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, case in CASES.items():
        run_case(name, case)


if __name__ == "__main__":
    main()
