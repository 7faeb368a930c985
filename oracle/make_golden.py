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
    # LM head fed by the LAST level's fused language tokens (lm_args.use_lm_f: False, cross_f_box_wrapper.py:224-227)
    "lmfused2_d32": dict(D=32, heads=4, image=(64, 64), channels=[16, 24], patch=[2, 1], strides=[16, 32],
                         layers=[2, 1], B=2, L=6, lens=[6, 4], lm=True, use_lm_f=False, seed=41),
    # fused language tokens chained into the next level (forward_language_f, :203-209), LM head on the chained tokens
    "fwdlang_sum_d32": dict(D=32, heads=4, image=(64, 64), channels=[16, 24], patch=[2, 1], strides=[16, 32],
                            layers=[1, 2], B=2, L=5, lens=[2, 5], lm=True, fwd_lang="sum", seed=43),
    "fwdlang_direct_d32": dict(D=32, heads=4, image=(64, 64), channels=[16, 24], patch=[2, 1], strides=[16, 32],
                               layers=[1, 1], B=2, L=5, lens=[5, 3], lm=True, use_lm_f=False, fwd_lang="direct", seed=47),
    # train-mode dropout (yml probabilities 0.1 / 0.15 / 0.1) with the keep masks recorded from the reference run
    "dropout2_d32": dict(D=32, heads=4, image=(64, 64), channels=[16, 24], patch=[2, 1], strides=[16, 32],
                         layers=[2, 2], B=2, L=6, lens=[6, 3], lm=False, dropout=True, seed=53),
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
                                      num_heads=case["heads"], patch=case["patch"],
                                      dropout=1.0 if case.get("dropout") else 0.0,
                                      use_lm_f=case.get("use_lm_f"), forward_language_f=case.get("fwd_lang"))
    m = ref_loader.build_reference_module(cfg, shapes, case["channels"], lm=case["lm"], seed=case["seed"],
                                          noun_classes=9, verb_classes=6)
    m.train()
    feats_in = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    lang_in = lang.clone().requires_grad_(True)
    # forward_language_f == "sum" adds in place (:206): hand the module a non-leaf alias of the leaf
    lang_arg = lang_in * 1.0 if case.get("fwd_lang") else lang_in
    rec = None
    if case.get("dropout"):
        with ref_loader.recorded_dropout(case["seed"] + 2000) as rec:
            out, lm = ref_loader.run_reference(m, feats_in, lang_arg, mask)
    else:
        out, lm = ref_loader.run_reference(m, feats_in, lang_arg, mask)
    loss = sum((out[k] * cot[k]).sum() for k in out)
    if lm is not None:
        loss = loss + lm["noun_logits"].sum() * 0.5 + (lm["verb_logits"] ** 2).sum() * 0.25
    loss.backward()

    # The reference's OWN bf16 error on these inputs (torch.autocast on CPU, same weights / masks): the anchor of the
    # GPU tolerances ("no worse than 2x the reference's own autocast-bf16 error").
    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    fp32_pg = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    fp32_out = {k: v.detach().clone() for k, v in out.items()}
    fp32_gl = lang_in.grad.clone()
    m.zero_grad(set_to_none=True)
    feats_b = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    lang_b = lang.clone().requires_grad_(True)
    lang_barg = lang_b * 1.0 if case.get("fwd_lang") else lang_b
    with torch.autocast("cpu", dtype=torch.bfloat16):
        if case.get("dropout"):
            with ref_loader.recorded_dropout(case["seed"] + 2000):
                out_b, lm_b = ref_loader.run_reference(m, feats_b, lang_barg, mask)
        else:
            out_b, lm_b = ref_loader.run_reference(m, feats_b, lang_barg, mask)
    loss_b = sum((out_b[k].float() * cot[k]).sum() for k in out_b)
    if lm_b is not None:
        loss_b = loss_b + lm_b["noun_logits"].float().sum() * 0.5 + (lm_b["verb_logits"].float() ** 2).sum() * 0.25
    loss_b.backward()
    ac_err = {f"out.{k}": rel(out_b[k].detach().float(), fp32_out[k]) for k in out_b}
    ac_err["glang"] = rel(lang_b.grad, fp32_gl)
    for k, p in m.named_parameters():
        if p.grad is not None and k in fp32_pg:
            ac_err[f"pgrad.{k}"] = rel(p.grad, fp32_pg[k])
    for k, p in m.named_parameters():   # restore the fp32 gradients for the blob below
        p.grad = fp32_pg.get(k)

    blob = {}
    for k, v in ac_err.items():
        blob[f"refbf16err.{k}"] = np.array(v, dtype=np.float64)
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
    blob["meta.use_lm_f"] = np.array(int(case.get("use_lm_f", True)))
    blob["meta.fwd_lang"] = np.array(case.get("fwd_lang") or "")
    if rec is not None:
        blob["meta.drop"] = np.array([cfg["args"]["patch_dropout"], cfg["args"]["token_dropout"], cfg["backproj_dropout"]])
        for lvl, d in ref_loader.masks_by_site(rec.masks, len(shapes), case["layers"]).items():
            for site, mk in d.items():
                blob[f"mask.{lvl}.{site}"] = np.packbits(mk.numpy().astype(np.uint8).reshape(-1))
                blob[f"maskshape.{lvl}.{site}"] = np.array(mk.shape)
    blob["meta.torch_version"] = np.array(torch.__version__)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **{k: (v.astype(np.float32) if v.dtype == np.float64 else v) for k, v in blob.items()})
    worst = max(ac_err.items(), key=lambda kv: kv[1])
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB), loss={float(loss):.6f}; reference's own bf16-autocast "
          f"error: worst {worst[0]} = {worst[1]:.3e}, out {max(v for k, v in ac_err.items() if k.startswith('out.')):.3e}")


def main():
    import sys
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    only = set(sys.argv[1:])
    for name, case in CASES.items():
        if not only or name in only:
            run_case(name, case)


if __name__ == "__main__":
    main()


# ---------------------------------------------------------------------------------------------------------------
# Optimizer golden: the UNMODIFIED reference RAdam (runner/metrics_losses/radam_optim.py) preceded by
# torch.nn.utils.clip_grad_norm_ (what Lightning's gradient_clip_val does), 8 steps so that the rectification threshold
# N_sma >= 5 is crossed (steps 1-5 leave the parameters untouched, radam_optim.py:86-87,98), ragged tensor sizes.
def run_radam_golden(name="radam8"):
    import sys
    if ref_loader.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    from runner.metrics_losses.radam_optim import RAdam
    g = torch.Generator().manual_seed(91)
    shapes = [(37, 16), (64,), (5, 3, 7), (1, 1, 33)]
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    extra = torch.nn.Parameter(torch.randn(50, generator=g))   # clipped together with `params`, stepped by another optimizer
    p0 = [p.detach().clone() for p in params]
    lr, wd, max_norm, steps = 2e-4, 1e-4, 4.0, 8                # ego_nao_res50_ego4dv2.yml:129,159,161
    opt = RAdam(params, lr=lr, weight_decay=wd)
    blob = {"meta.lr": np.array(lr), "meta.wd": np.array(wd), "meta.max_norm": np.array(max_norm), "meta.steps": np.array(steps)}
    for t in range(steps):
        scale = 6.0 if t % 2 == 0 else 0.3   # some steps clip (norm > 4), some do not
        for i, p in enumerate(params):
            p.grad = scale * torch.randn(p.shape, generator=g)
            blob[f"grad.{t}.{i}"] = p.grad.numpy().copy()
        extra.grad = scale * torch.randn(extra.shape, generator=g)
        blob[f"extra_sqnorm.{t}"] = np.array(float(extra.grad.double().pow(2).sum()))
        torch.nn.utils.clip_grad_norm_(params + [extra], max_norm)
        opt.step()
        if t in (4, 5):   # around the N_sma threshold
            for i, p in enumerate(params):
                blob[f"p_after{t + 1}.{i}"] = p.detach().numpy().copy()
    for i, p in enumerate(params):
        blob[f"p0.{i}"] = p0[i].numpy()
        blob[f"p.{i}"] = p.detach().numpy().copy()
        blob[f"m.{i}"] = opt.state[p]["exp_avg"].numpy().copy()
        blob[f"v.{i}"] = opt.state[p]["exp_avg_sq"].numpy().copy()
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e3:.1f} kB)")


if __name__ == "__main__":
    import sys as _sys
    if not _sys.argv[1:] or "radam8" in _sys.argv[1:]:
        run_radam_golden()
