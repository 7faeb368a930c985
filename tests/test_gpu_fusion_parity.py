"""GPU parity of the CUDA-backed CrossFusionBoxWrapper (through the C ABI) against
 (1) the golden vectors frozen from the unmodified reference (tests/golden/*.npz) and
 (2) the CPU oracle (oracle/ref_math.py) on the shipped token width D = 896 / 712.

Tolerances (bf16 compute, fp32 accumulate; SURVEY §8c "observed error budget"): the reference's own
autocast-bf16 run deviates from its fp32 run by rel-Frobenius 4.2e-3 on fused features and 5.0e-3 on
weight grads; we require rel-Frobenius <= 1e-2 on fused features / LM logits and on gradients (2x the reference's own
bf16 error), and
max-abs <= 5e-2 * max(1, rms(ref)) on fused features."""
import pytest
import torch

from oracle import ref_math
from tests.fusion_testlib import build_module, param_dict, run_module
from tests.golden_utils import CASES, grad_bound, load_golden, rel_fro

pytestmark = pytest.mark.gpu

REL_OUT, REL_GRAD, MAXABS = 1e-2, 1e-2, 5e-2


def _check_out(got, ref, name):
    got, ref = got.float().cpu(), ref.float()
    assert got.shape == ref.shape, name
    assert torch.isfinite(got).all(), name
    r = rel_fro(got, ref)
    rms = float(ref.pow(2).mean().sqrt())
    ma = float((got - ref).abs().max())
    assert r < REL_OUT, f"{name}: rel-Frobenius {r:.3e}"
    assert ma < MAXABS * max(1.0, rms), f"{name}: max-abs {ma:.3e} (rms {rms:.3e})"


@pytest.mark.parametrize("name", CASES)
def test_golden_forward_backward(name):
    g = load_golden(name)
    feats = g["features"]
    shapes = [tuple(feats[k].shape[2:]) for k in sorted(feats, key=int)]
    channels = [feats[k].shape[1] for k in sorted(feats, key=int)]
    D = g["lang"].shape[-1]
    m = build_module(D, shapes, channels, g["patch"], g["layers"], g["heads"], lm=g["lm_on"], use_lm_f=g["use_lm_f"],
                     forward_language_f=g["fwd_lang"])
    missing, unexpected = m.load_state_dict(g["params"], strict=False)
    assert not unexpected
    assert all(("pos_embedding" in k or "padding_mask" in k) for k in missing), missing
    m.train()
    f_in = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
    lang = g["lang"].cuda().requires_grad_(True)
    # forward_language_f == "sum" adds in place in the reference (:206); hand the module a non-leaf alias like make_golden does
    out, lm = run_module(m, f_in, lang * 1.0 if g["fwd_lang"] else lang, g["att_mask"].cuda())
    for k in out:
        _check_out(out[k], g["out"][k], f"{name}/features.{k}")
    loss = sum((out[k].float() * g["cot"][k].cuda()).sum() for k in out)
    if g["lm_on"]:
        _check_out(lm["noun_logits"], g["lm"]["noun_logits"], "noun_logits")
        _check_out(lm["verb_logits"], g["lm"]["verb_logits"], "verb_logits")
        loss = loss + lm["noun_logits"].sum() * 0.5 + (lm["verb_logits"] ** 2).sum() * 0.25
    loss.backward()
    torch.cuda.synchronize()
    # Gradient bounds: 1e-2 (the bound at the shipped widths), relaxed per tensor to 2x the reference's OWN bf16-autocast
    # error on these inputs where that is larger (capped at 2e-2): the golden cases are D = 32 .. 64 wide, where a
    # contraction averages 14 .. 28x fewer rounding errors than at D = 896 (the reference itself is 5e-3 .. 8e-3 off).
    for k in f_in:
        assert rel_fro(f_in[k].grad.cpu(), g["gfeat"][k]) < 2e-2, f"grad features.{k}"
    assert rel_fro(lang.grad.cpu(), g["glang"]) < grad_bound(g, "glang")
    pd = param_dict(m)
    n_checked = 0
    for k, gr in g["pgrads"].items():
        assert pd[k].grad is not None, k
        r = rel_fro(pd[k].grad.cpu(), gr)
        assert r < grad_bound(g, f"pgrad.{k}"), f"pgrad {k}: {r:.3e} (bound {grad_bound(g, 'pgrad.' + k):.2e})"
        n_checked += 1
    assert n_checked == len(g["pgrads"])
    for k, p in pd.items():
        if k.endswith("heatmap_token"):
            assert p.grad is None  # unused parameter, as in the reference


def _oracle_case(D, heads, shapes, channels, patch, layers, B, L, lens, seed):
    m = build_module(D, shapes, channels, patch, layers, heads, seed=seed)
    m.train()
    g = torch.Generator().manual_seed(seed + 1)
    feats = {str(i): torch.relu(torch.randn(B, c, h, w, generator=g)) for i, ((h, w), c) in enumerate(zip(shapes, channels))}
    lang = 0.5 * torch.randn(B, L, D, generator=g)
    mask = torch.zeros(B, L, dtype=torch.int64)
    for b, n in enumerate(lens):
        mask[b, :n] = 1
    cot = {k: torch.randn(v.shape, generator=g) for k, v in feats.items()}
    # oracle (CPU fp32)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in param_dict(m).items()}
    f_cpu = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    l_cpu = lang.clone().requires_grad_(True)
    ref, _ = ref_math.cross_fusion_forward(f_cpu, l_cpu, mask, sd, patch, heads, layers)
    sum((ref[k] * cot[k]).sum() for k in ref).backward()
    # CUDA path
    f_gpu = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
    l_gpu = lang.cuda().requires_grad_(True)
    out, _ = run_module(m, f_gpu, l_gpu, mask.cuda())
    sum((out[k].float() * cot[k].cuda()).sum() for k in out).backward()
    torch.cuda.synchronize()
    for k in out:
        _check_out(out[k], ref[k].detach(), f"features.{k}")
        assert rel_fro(f_gpu[k].grad.cpu(), f_cpu[k].grad) < REL_GRAD, f"grad features.{k}"
    assert rel_fro(l_gpu.grad.cpu(), l_cpu.grad) < REL_GRAD
    worst = ("", 0.0)
    for k, p in param_dict(m).items():
        if k.endswith("heatmap_token"):
            continue
        r = rel_fro(p.grad.cpu(), sd[k].grad)
        if r > worst[1]:
            worst = (k, r)
    assert worst[1] < REL_GRAD, f"worst param grad {worst}"


def test_oracle_parity_ego4dv2_width_c5_and_c4():
    """D = 896 (head_dim 224), two levels with the C4/C5 patch sizes (2, 1), ragged language lengths."""
    _oracle_case(896, 4, [(16, 24), (8, 12)], [64, 128], [2, 1], [2, 2], B=2, L=24, lens=[24, 9], seed=5)


def test_oracle_parity_ego4dv1_width():
    """D = 712 (head_dim 178 -> padded to 192), patch 4 level, ragged language."""
    _oracle_case(712, 4, [(16, 24)], [32], [4], [2], B=2, L=16, lens=[5, 16], seed=6)


def test_oracle_parity_multi_tile_sequence():
    """S > 128 so attention spans several query / key tiles, 4 layers."""
    _oracle_case(256, 4, [(20, 24)], [48], [1], [4], B=2, L=40, lens=[40, 17], seed=7)


def test_oracle_parity_sample_without_language_and_odd_grid():
    """A sample whose language context is entirely padding (all L keys masked; the reference still attends over the
    n visual keys) next to a full one; 15 x 19 token grid (the 480 x 608 image of the sweep,
    SURVEY 8d config 5): n = 285 is no multiple of any tile size."""
    _oracle_case(256, 4, [(30, 38)], [16], [2], [2], B=2, L=16, lens=[0, 16], seed=8)


def test_eval_mode_matches_oracle_and_is_deterministic():
    """Inference (config 4): eval + no_grad uses the same kernels without dropout; repeated calls are bit-identical."""
    m = build_module(256, [(16, 24)], [32], [2], [2], 4, dropout=True, seed=9)
    g = torch.Generator().manual_seed(10)
    feats = {"0": torch.relu(torch.randn(2, 32, 16, 24, generator=g))}
    lang = 0.5 * torch.randn(2, 12, 256, generator=g)
    mask = torch.ones(2, 12, dtype=torch.int64)
    mask[1, 7:] = 0
    sd = {k: v.detach().cpu() for k, v in param_dict(m).items()}
    ref, _ = ref_math.cross_fusion_forward({k: v.clone() for k, v in feats.items()}, lang.clone(), mask, sd, [2], 4, [2])
    m.eval()
    with torch.no_grad():
        o1, _ = run_module(m, {k: v.cuda() for k, v in feats.items()}, lang.cuda(), mask.cuda())
        o2, _ = run_module(m, {k: v.cuda() for k, v in feats.items()}, lang.cuda(), mask.cuda())
    _check_out(o1["0"], ref["0"].detach(), "eval features.0")
    assert torch.equal(o1["0"], o2["0"])


def test_train_mode_dropout_is_reproducible_per_seed():
    """Counter-based dropout keyed by a per-step seed drawn from torch's CPU generator (as the reference's dropout
    follows torch.manual_seed): the same seed reproduces outputs and gradients bit for bit; the next differs."""
    m = build_module(256, [(16, 24)], [32], [2], [2], 4, dropout=True, seed=11)
    m.train()
    g = torch.Generator().manual_seed(12)
    feats = {"0": torch.relu(torch.randn(2, 32, 16, 24, generator=g)).cuda()}
    lang = (0.5 * torch.randn(2, 12, 256, generator=g)).cuda()
    mask = torch.ones(2, 12, dtype=torch.int64).cuda()

    def run(seed):
        torch.manual_seed(seed)
        m.zero_grad(set_to_none=True)
        out, _ = run_module(m, {k: v.clone() for k, v in feats.items()}, lang.clone(), mask)
        out["0"].float().sum().backward()
        torch.cuda.synchronize()
        w1 = next(p for k, p in param_dict(m).items() if k.endswith("layers.0.linear1.weight"))
        return out["0"].detach().clone(), w1.grad.clone()

    a, ga = run(123)
    b, gb = run(123)
    c, _ = run(124)
    assert torch.equal(a, b)
    # weight gradients accumulate with fp32 atomics (split-K): equal up to summation order
    assert rel_fro(ga, gb) < 1e-5
    assert not torch.equal(a, c)


def test_channels_last_feature_maps_in_and_out():
    """channels_last maps from a channels_last backbone go through the module without a layout conversion:
    same numbers as with NCHW maps, the fused maps and the input gradients come back channels_last."""
    m = build_module(256, [(16, 24), (8, 12)], [32, 64], [2, 1], [1, 1], 4, seed=61)
    m.train()
    g = torch.Generator().manual_seed(62)
    feats = {"0": torch.relu(torch.randn(2, 32, 16, 24, generator=g)), "1": torch.relu(torch.randn(2, 64, 8, 12, generator=g))}
    lang = 0.5 * torch.randn(2, 10, 256, generator=g)
    mask = torch.ones(2, 10, dtype=torch.int64)
    outs, grads = [], []
    for fmt in (torch.contiguous_format, torch.channels_last):
        f_in = {k: v.cuda().contiguous(memory_format=fmt).requires_grad_(True) for k, v in feats.items()}
        out, _ = run_module(m, f_in, lang.cuda(), mask.cuda())
        for k in out:
            assert out[k].is_contiguous(memory_format=fmt)
        m.zero_grad(set_to_none=True)
        sum(o.float().pow(2).sum() for o in out.values()).backward()
        outs.append({k: v.detach().float().cpu() for k, v in out.items()})
        grads.append({k: v.grad.float().cpu() for k, v in f_in.items()})
        if fmt == torch.channels_last:
            for k, v in f_in.items():
                assert v.grad.is_contiguous(memory_format=torch.channels_last)
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k])
        assert rel_fro(grads[1][k], grads[0][k]) < 1e-5   # split-K atomics reorder fp32 sums upstream of this gradient
