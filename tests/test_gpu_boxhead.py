"""GPU: SURVEY 8f N1 -- the RoI box head (TwoMLPHead fc6 / fc7 + ReLU, box_regressor, noun / verb classifiers;
modeling/obj_detection/roi_wrappers.py:198-214) on xf_gemm against the stock torch modules in fp32, at the shipped widths
(12544-wide fc6, R = 1280, 129 nouns, 82 verbs, 128 boxes per image).

Tolerance ("no worse than 2x the reference's own autocast-bf16 error on the same inputs"): every output and
gradient must be within max(1e-2, 2 x E) rel-Frobenius of the fp32 modules, where E is the error of the SAME torch modules
under torch.autocast(bf16).  The anchor matters here: with random weights a ReLU net is ill-conditioned for gradients --
any bf16 rounding flips the sign of the ~0.3 % of pre-activations that sit within rounding distance of zero, and each flip
switches a unit's whole gradient contribution (E is 3-6 % on the input gradient for torch's own bf16 kernels)."""
import types

import pytest
import torch
from torch import nn

from tests.golden_utils import rel_fro
from transfusion_b200 import _lib
from transfusion_b200.obj_detection import FusedBoxHead

pytestmark = pytest.mark.gpu
DEV = "cuda"


class TwoMLPHead(nn.Module):
    """torchvision.models.detection.faster_rcnn.TwoMLPHead (fc6, fc7 with ReLU), restated so the test has no torchvision import."""

    def __init__(self, in_channels, representation_size):
        super().__init__()
        self.fc6 = nn.Linear(in_channels, representation_size)
        self.fc7 = nn.Linear(representation_size, representation_size)

    def forward(self, x):
        x = x.flatten(start_dim=1)
        return torch.relu(self.fc7(torch.relu(self.fc6(x))))


def _reference_roi(in_f, R, nouns, verbs):
    torch.manual_seed(0)
    roi = types.SimpleNamespace()
    roi.roi_head_wrap = types.SimpleNamespace(box_head=TwoMLPHead(in_f, R).to(DEV))
    roi.dropout_1 = nn.Identity()
    roi.classif_dropout = nn.Identity()
    roi.box_regressor = nn.Sequential(nn.Identity(), nn.Linear(R, 4 * nouns)).to(DEV)
    roi.noun_classifier = nn.Linear(R, nouns).to(DEV)
    roi.verb_classifier = nn.Linear(R, verbs).to(DEV) if verbs else None
    return roi


def _params(roi, verbs):
    params = {"fc6.w": roi.roi_head_wrap.box_head.fc6.weight, "fc6.b": roi.roi_head_wrap.box_head.fc6.bias,
              "fc7.w": roi.roi_head_wrap.box_head.fc7.weight, "fc7.b": roi.roi_head_wrap.box_head.fc7.bias,
              "reg.w": roi.box_regressor[1].weight, "reg.b": roi.box_regressor[1].bias,
              "noun.w": roi.noun_classifier.weight, "noun.b": roi.noun_classifier.bias}
    if verbs:
        params.update({"verb.w": roi.verb_classifier.weight, "verb.b": roi.verb_classifier.bias})
    return params


def _torch_forward(roi, x):
    f = roi.roi_head_wrap.box_head(x)
    return roi.box_regressor(f), roi.noun_classifier(f), (roi.verb_classifier(f) if roi.verb_classifier is not None else None)


@pytest.mark.parametrize("rows,in_f,R,nouns,verbs", [(256, 256 * 7 * 7, 1280, 129, 82), (150, 12544, 1024, 88, 75), (37, 96, 64, 9, None)])
def test_box_head_matches_torch_modules(rows, in_f, R, nouns, verbs):
    roi = _reference_roi(in_f, R, nouns, verbs)
    head = FusedBoxHead.from_roi_heads(roi)
    head.train()
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.relu(torch.randn(rows, in_f, device=DEV, generator=g)) * 0.5
    shape4 = (rows, 256, 7, 7) if in_f == 12544 else (rows, in_f)
    x_ref = x.clone().reshape(shape4).requires_grad_(True)
    x_xf = x.clone().reshape(shape4).requires_grad_(True)
    cot = [torch.randn(rows, 4 * nouns, device=DEV, generator=g), torch.randn(rows, nouns, device=DEV, generator=g),
           torch.randn(rows, verbs, device=DEV, generator=g) if verbs else None]

    ref = _torch_forward(roi, x_ref)
    sum((o * c).sum() for o, c in zip(ref, cot) if o is not None).backward()
    params = _params(roi, verbs)
    ref_grads = {k: p.grad.clone() for k, p in params.items()}
    for p in params.values():
        p.grad = None
    # the reference's own bf16 error (torch.autocast on the same modules): the anchor of the bounds
    x_ac = x.clone().reshape(shape4).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out_ac = _torch_forward(roi, x_ac)
    sum((o.float() * c).sum() for o, c in zip(out_ac, cot) if o is not None).backward()
    err_ac = {"x": rel_fro(x_ac.grad, x_ref.grad)}
    err_ac.update({k: rel_fro(p.grad, ref_grads[k]) for k, p in params.items()})
    err_ac.update({f"out{i}": rel_fro(o.float(), r) for i, (o, r) in enumerate(zip(out_ac, ref)) if r is not None})
    for p in params.values():
        p.grad = None

    def bound(key):
        return max(1e-2, 2.0 * err_ac[key])

    n0 = _lib.lib().xf_launch_count()
    out = head(x_xf)
    assert _lib.lib().xf_launch_count() - n0 >= 4   # the CUDA library ran (cast + 3 GEMMs)
    for i, (o, r, name) in enumerate(zip(out, ref, ("box_regression", "class_logits", "verb_logits"))):
        if r is None:
            assert o is None
            continue
        assert o.dtype == torch.float32 and o.shape == r.shape
        assert rel_fro(o, r) < bound(f"out{i}"), (name, rel_fro(o, r), err_ac[f"out{i}"])
    sum((o * c).sum() for o, c in zip(out, cot) if o is not None).backward()
    assert rel_fro(x_xf.grad, x_ref.grad) < bound("x"), (rel_fro(x_xf.grad, x_ref.grad), err_ac["x"])
    for k, p in params.items():
        assert p.grad is not None, k
        assert rel_fro(p.grad, ref_grads[k]) < bound(k), (k, rel_fro(p.grad, ref_grads[k]), err_ac[k])


def test_box_head_rejects_unfused_dropout_and_cpu():
    roi = _reference_roi(96, 64, 9, 6)
    roi.classif_dropout = nn.Dropout(0.2)
    with pytest.raises(NotImplementedError):
        FusedBoxHead.from_roi_heads(roi)
    head = FusedBoxHead(96, 64, 9, 6)
    with pytest.raises(RuntimeError):
        head(torch.randn(4, 96))
