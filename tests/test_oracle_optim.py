"""CPU: the optimizer oracle (oracle/ref_optim.py: clip_grad_norm_ + the reference's RAdam) against the golden
trajectory frozen from the UNMODIFIED reference class (tests/golden/radam8.npz, oracle/make_golden.py) and, where the
reference tree / oracle/_ref exists, against the class itself.  fp32 tolerance 2e-6 relative."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader, ref_optim
from tests.golden_utils import GOLDEN_DIR, rel_fro


def load_radam_golden():
    z = np.load(os.path.join(GOLDEN_DIR, "radam8.npz"))
    steps = int(z["meta.steps"])
    n = len([k for k in z.files if k.startswith("p0.")])
    g = {"lr": float(z["meta.lr"]), "wd": float(z["meta.wd"]), "max_norm": float(z["meta.max_norm"]), "steps": steps, "n": n}
    g["p0"] = [torch.from_numpy(z[f"p0.{i}"].copy()) for i in range(n)]
    g["grads"] = [[torch.from_numpy(z[f"grad.{t}.{i}"].copy()) for i in range(n)] for t in range(steps)]
    g["extra"] = [float(z[f"extra_sqnorm.{t}"]) for t in range(steps)]
    for key in ("p", "m", "v", "p_after5", "p_after6"):
        g[key] = [torch.from_numpy(z[f"{key}.{i}"].copy()) for i in range(n)]
    return g


def test_oracle_radam_matches_golden_trajectory():
    g = load_radam_golden()
    ps, ms, vs = ref_optim.run_steps(g["p0"], g["grads"], g["lr"], g["wd"], g["max_norm"], g["extra"])
    for i in range(g["n"]):
        assert rel_fro(ps[i], g["p"][i]) < 2e-6
        assert rel_fro(ms[i], g["m"][i]) < 2e-6
        assert rel_fro(vs[i], g["v"][i]) < 2e-6
    # the rectification threshold: parameters do not move during the first 5 steps (N_sma < 5, radam_optim.py:86-87)
    for i in range(g["n"]):
        assert torch.equal(g["p_after5"][i], g["p0"][i])
        assert not torch.equal(g["p_after6"][i], g["p0"][i])
    ps5, _, _ = ref_optim.run_steps(g["p0"], g["grads"][:5], g["lr"], g["wd"], g["max_norm"], g["extra"])
    assert all(torch.equal(a, b) for a, b in zip(ps5, g["p0"]))


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")
def test_oracle_radam_matches_reference_class_with_degenerated_sgd():
    if ref_loader.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    from runner.metrics_losses.radam_optim import RAdam
    torch.manual_seed(3)
    params = [torch.nn.Parameter(torch.randn(19, 5)), torch.nn.Parameter(torch.randn(7))]
    p0 = [p.detach().clone() for p in params]
    opt = RAdam(params, lr=1e-3, weight_decay=0.0, degenerated_to_sgd=True)
    grads = []
    for t in range(7):
        gs = [torch.randn(p.shape) for p in params]
        grads.append(gs)
        for p, g_ in zip(params, gs):
            p.grad = g_.clone()
        opt.step()
    ps, ms, vs = ref_optim.run_steps(p0, grads, 1e-3, 0.0, 0.0, None, degenerated_to_sgd=True)
    for i, p in enumerate(params):
        assert rel_fro(ps[i], p.detach()) < 2e-6
        assert rel_fro(vs[i], opt.state[p]["exp_avg_sq"]) < 2e-6
