"""GPU: the fused clip + RAdam step (xf_grad_sqnorm / xf_radam_step through transfusion_b200.optim.FusedRAdam) against the
golden trajectory of the UNMODIFIED reference optimizer (tests/golden/radam8.npz) and the CPU oracle, and the coherence of
the bf16 weight copies the step emits for the next forward (cross_fusion/level_fn.py cache).  fp32: 2e-6 relative."""
import pytest
import torch

from oracle import ref_math, ref_optim
from tests.fusion_testlib import build_module, param_dict, run_module
from tests.golden_utils import rel_fro
from tests.test_oracle_optim import load_radam_golden
from transfusion_b200 import _lib, optim

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_fused_radam_matches_reference_trajectory():
    g = load_radam_golden()
    params = [torch.nn.Parameter(p.clone().to(DEV)) for p in g["p0"]]
    opt = optim.FusedRAdam(params, lr=g["lr"], weight_decay=g["wd"], max_grad_norm=g["max_norm"])
    for t in range(g["steps"]):
        for p, gr in zip(params, g["grads"][t]):
            p.grad = gr.to(DEV)
        opt.step(extra_sqnorm=torch.tensor([g["extra"][t]], device=DEV))
        if t == 4:
            assert all(torch.equal(p.detach().cpu(), p0) for p, p0 in zip(params, g["p0"]))   # N_sma < 5: no movement
    for i, p in enumerate(params):
        assert rel_fro(p.detach().cpu(), g["p"][i]) < 2e-6, i
        assert rel_fro(opt.state[p]["exp_avg"].cpu(), g["m"][i]) < 2e-6, i
        assert rel_fro(opt.state[p]["exp_avg_sq"].cpu(), g["v"][i]) < 2e-6, i
        assert opt.state[p]["step"] == g["steps"]


def test_fused_radam_many_tensors_degenerated_sgd_and_sqnorm():
    torch.manual_seed(5)
    shapes = [(3 + i, 5 + (i % 7)) for i in range(70)] + [(1,), (2,), (897,)]   # > 2 launches of 32 jobs, odd sizes / tails
    p0 = [torch.randn(s) for s in shapes]
    grads = [[torch.randn(s) * (3.0 if t % 2 else 0.05) for s in shapes] for t in range(7)]
    params = [torch.nn.Parameter(p.clone().to(DEV)) for p in p0]
    opt = optim.FusedRAdam(params, lr=1e-3, weight_decay=1e-2, degenerated_to_sgd=True, max_grad_norm=1.0)
    for t in range(7):
        for p, gr in zip(params, grads[t]):
            p.grad = gr.to(DEV)
        if t == 0:
            sq = optim.grad_sqnorm([p.grad for p in params])
            ref = sum(float(gr.double().pow(2).sum()) for gr in grads[0])
            assert abs(float(sq) - ref) / ref < 1e-5
        opt.step()
    ps, ms, vs = ref_optim.run_steps(p0, grads, 1e-3, 1e-2, 1.0, None, degenerated_to_sgd=True)
    for i, p in enumerate(params):
        assert rel_fro(p.detach().cpu(), ps[i]) < 2e-6, i
        assert rel_fro(opt.state[p]["exp_avg_sq"].cpu(), vs[i]) < 2e-6, i


def _small_case():
    D, shapes, channels, patch, layers, B, L = 256, [(16, 24)], [32], [2], [2], 2, 12
    m = build_module(D, shapes, channels, patch, layers, 4, seed=31)
    g = torch.Generator().manual_seed(32)
    feats = {"0": torch.relu(torch.randn(B, 32, 16, 24, generator=g))}
    lang = 0.5 * torch.randn(B, L, D, generator=g)
    mask = torch.ones(B, L, dtype=torch.int64)
    mask[1, 7:] = 0
    return m, feats, lang, mask, (patch, 4, layers)


def test_training_with_fused_radam_keeps_bf16_weight_copies_coherent():
    """Two optimizer steps (past the rectification threshold via degenerated_to_sgd): the forward after each step must see
    the UPDATED weights although it no longer casts them (the step wrote the bf16 copies); checked against the oracle run
    on the updated fp32 parameters.  The steady-state forward launches no weight-cast kernel."""
    m, feats, lang, mask, (patch, H, layers) = _small_case()
    m.train()
    params = [p for k, p in param_dict(m).items() if not k.endswith("heatmap_token")]
    opt = optim.FusedRAdam(params, lr=5e-3, degenerated_to_sgd=True, max_grad_norm=4.0)
    f_gpu = {k: v.cuda() for k, v in feats.items()}
    launches = []
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        n0 = _lib.lib().xf_launch_count()
        out, _ = run_module(m, dict(f_gpu), lang.cuda(), mask.cuda())
        launches.append(_lib.lib().xf_launch_count() - n0)
        sd = {k: v.detach().cpu() for k, v in param_dict(m).items()}
        ref, _ = ref_math.cross_fusion_forward({k: v.clone() for k, v in feats.items()}, lang, mask, sd, patch, H, layers)
        assert rel_fro(out["0"].detach().float().cpu(), ref["0"]) < 1e-2, f"forward {it} does not see the current weights"
        out["0"].float().pow(2).sum().backward()
        w_before = params[0].detach().clone()
        opt.step()
        assert not torch.equal(params[0].detach(), w_before)
        for p in params:   # every flat cached copy is exactly the rounded updated parameter
            c = getattr(p, "_xf_bf16", None)
            if c is not None and c[2]:
                assert c[3] == "opt" and torch.equal(c[1].reshape(-1), p.detach().bfloat16().reshape(-1))
    assert launches[1] == launches[2] == launches[0] - 1   # the weight-cast launch is gone after the first step


def test_foreign_optimizer_writing_through_data_is_never_served_a_stale_copy():
    """The reference's RAdam updates parameters through `p.data` (radam_optim.py:96), which does not touch the version
    counter: in training the cached copies must not be trusted, in inference train()/eval() switches drop them."""
    m, feats, lang, mask, (patch, H, layers) = _small_case()
    f_gpu = {k: v.cuda() for k, v in feats.items()}

    def fwd():
        out, _ = run_module(m, dict(f_gpu), lang.cuda(), mask.cuda())
        return out["0"].detach().float().cpu()

    m.train()
    a = fwd()
    with torch.no_grad():
        for k, p in param_dict(m).items():
            if k.endswith("linear1.weight"):
                p.data.mul_(1.5)                    # version counter unchanged
    b = fwd()
    assert rel_fro(b, a) > 1e-3                    # training forward re-cast the weights
    m.eval()
    with torch.no_grad():
        c1 = fwd()
        n0 = _lib.lib().xf_launch_count()
        c2 = fwd()
        n1 = _lib.lib().xf_launch_count()
        c3 = fwd()
        n2 = _lib.lib().xf_launch_count()
    assert torch.equal(c1, c2) and torch.equal(c2, c3) and (n2 - n1) == (n1 - n0)
    m.train(); m.eval()                             # mode switch drops the cache
    with torch.no_grad():
        for k, p in param_dict(m).items():
            if k.endswith("linear1.weight"):
                p.data.mul_(0.5)
        d = fwd()
    sd = {k: v.detach().cpu() for k, v in param_dict(m).items()}
    ref, _ = ref_math.cross_fusion_forward({k: v.clone() for k, v in feats.items()}, lang, mask, sd, patch, H, layers)
    assert rel_fro(d, ref["0"]) < 1e-2
