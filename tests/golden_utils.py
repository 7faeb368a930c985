"""Helpers shared by the oracle / GPU parity tests: load a golden npz (made by
oracle/make_golden.py from the unmodified reference) into torch tensors."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["fusion4_d32", "c5_d64_lm", "c4_d40_oddhead", "lmfused2_d32", "fwdlang_sum_d32", "fwdlang_direct_d32"]
DROPOUT_CASES = ["dropout2_d32"]   # reference run in train mode with recorded keep masks (oracle/ref_loader.recorded_dropout)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {"params": {}, "pgrads": {}, "features": {}, "cot": {}, "out": {}, "gfeat": {}, "lm": {}}
    for k in z.files:
        v = z[k]
        if k.startswith("param."):
            g["params"][k[6:]] = torch.from_numpy(v.copy())
        elif k.startswith("pgrad."):
            g["pgrads"][k[6:]] = torch.from_numpy(v.copy())
        elif k.startswith("in.features."):
            g["features"][k[12:]] = torch.from_numpy(v.copy())
        elif k.startswith("in.cotangent."):
            g["cot"][k[13:]] = torch.from_numpy(v.copy())
        elif k.startswith("out.features."):
            g["out"][k[13:]] = torch.from_numpy(v.copy())
        elif k.startswith("grad.features."):
            g["gfeat"][k[14:]] = torch.from_numpy(v.copy())
        elif k.startswith("out.lm."):
            g["lm"][k[7:]] = torch.from_numpy(v.copy())
    g["lang"] = torch.from_numpy(z["in.language_f"].copy())
    g["att_mask"] = torch.from_numpy(z["in.att_mask"].copy())
    g["glang"] = torch.from_numpy(z["grad.language_f"].copy())
    g["patch"] = [int(x) for x in z["meta.patch"]]
    g["layers"] = [int(x) for x in z["meta.layers"]]
    g["heads"] = int(z["meta.heads"])
    g["lm_on"] = bool(int(z["meta.lm"]))
    g["use_lm_f"] = bool(int(z["meta.use_lm_f"])) if "meta.use_lm_f" in z.files else True
    g["fwd_lang"] = (str(z["meta.fwd_lang"]) or False) if "meta.fwd_lang" in z.files else False
    g["drop"] = tuple(float(x) for x in z["meta.drop"]) if "meta.drop" in z.files else (0.0, 0.0, 0.0)
    # the reference's own bf16-autocast error per tensor on these inputs (oracle/make_golden.py): anchors the GPU bounds
    g["ref_bf16_err"] = {k[len("refbf16err."):]: float(z[k]) for k in z.files if k.startswith("refbf16err.")}
    g["masks"] = {}
    for k in z.files:
        if k.startswith("mask."):
            _, lvl, site = k.split(".", 2)
            shape = tuple(int(x) for x in z["maskshape." + k[5:]])
            n = int(np.prod(shape))
            bits = np.unpackbits(z[k])[:n].reshape(shape)
            g["masks"].setdefault(lvl, {})[site] = torch.from_numpy(bits.astype(np.bool_))
    return g


def rel_fro(a, b):
    a = a.detach().double()
    b = b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def grad_bound(g, key, floor=1e-2, cap=2e-2):
    """"no worse than 2x the reference's own autocast-bf16 error on the same inputs", never tighter than
    `floor` (the bound used at the shipped widths) and never looser than `cap`."""
    return min(cap, max(floor, 2.0 * g["ref_bf16_err"].get(key, 0.0)))
