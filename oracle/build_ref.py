"""TEST INFRASTRUCTURE ONLY — recipe that stages the UNMODIFIED reference hot path under ``oracle/_ref/``.

    python -m oracle.build_ref            # needs /root/reference (the build container)

The reference is pure Python (no setup.py, nothing to compile), so "building" it means copying, byte for byte, exactly the
modules its ``CrossFusionBoxWrapper`` imports (traced from sys.modules after importing it, see FILES) plus the fusion /
run YAMLs, keeping the package layout.  ``oracle/_ref/`` is git-ignored (no reference source enters the history) but NOT
gpurun-ignored: it travels to the GPU box with the snapshot, where ``bench.py --impl reference`` /
``--impl reference-gpu`` and tests/test_oracle_vs_reference.py run the reference's own code (``cpu_baseline.kind =
"reference"``).  MANIFEST.json records the sha256 of every staged file so a stale or edited copy is detected.
Nothing in the product imports this directory.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC = os.environ.get("XF_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

FILES = [
    "modeling/__init__.py",
    "modeling/cross_fusion/cross_f_layers.py",
    "modeling/cross_fusion/cross_f_wrapper.py",
    "modeling/cross_fusion/cross_qkv_layers.py",
    "modeling/cross_fusion/utils.py",
    "modeling/cross_fusion/ego_fusion/cross_f_box_asymm.py",
    "modeling/cross_fusion/ego_fusion/cross_f_box_layers.py",
    "modeling/cross_fusion/ego_fusion/cross_f_box_wrapper.py",
    "modeling/cross_fusion/ego_fusion/lm_layers.py",
    "modeling/cross_fusion/ego_fusion/torch18_adapters.py",
    "modeling/cross_fusion/ego_fusion/cross_fusion_config_sym_ego_res50.yml",
    "modeling/obj_detection/wrapper_utils.py",
    "runner/metrics_losses/radam_optim.py",   # the optimizer of SURVEY 8f N4 (oracle for the fused RAdam kernel)
]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose=True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"oracle/_ref: {SRC} not present; keeping whatever is staged")
        return os.path.isfile(os.path.join(DST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"oracle/_ref: staged {len(manifest)} unmodified reference files from {SRC}")
    return True


def verify() -> bool:
    """True when every staged file still matches its recorded hash."""
    mpath = os.path.join(DST, "MANIFEST.json")
    if not os.path.isfile(mpath):
        return False
    with open(mpath) as f:
        man = json.load(f)["files"]
    return all(os.path.isfile(os.path.join(DST, r)) and _sha(os.path.join(DST, r)) == h for r, h in man.items())


if __name__ == "__main__":
    build()
