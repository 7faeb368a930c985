"""TEST INFRASTRUCTURE ONLY — CPU restatement of the optimizer step that follows the fusion path's backward
(SURVEY 8f N4): Lightning's global-norm clip (runner/run_experiment.py:445-446 -> torch.nn.utils.clip_grad_norm_)
followed by the reference's RAdam (runner/metrics_losses/radam_optim.py:30-104).  Pinned against the unmodified
reference class by tests/test_oracle_optim.py (where the reference tree or oracle/_ref exists) and by
tests/golden/radam8.npz (oracle/make_golden.py) everywhere."""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch


def clip_coef(grads: Sequence[torch.Tensor], max_norm: float, extra_sqnorm: float = 0.0) -> float:
    """torch.nn.utils.clip_grad_norm_(norm_type=2): min(1, max_norm / (total_norm + 1e-6)); `extra_sqnorm` is the squared
    norm of gradients that are clipped together with these but stepped elsewhere."""
    if max_norm <= 0:
        return 1.0
    total = math.sqrt(sum(float(g.double().pow(2).sum()) for g in grads) + extra_sqnorm)
    return min(1.0, max_norm / (total + 1e-6))


def radam_step(p, g, m, v, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
               degenerated_to_sgd: bool = False):
    """One RAdam update of one tensor, fp32 like the reference (radam_optim.py:57-102).  `step` is the 1-based count AFTER
    the increment (:65).  Returns (p, m, v)."""
    beta1, beta2 = betas
    v = v * beta2 + (1 - beta2) * g * g                     # :62
    m = m * beta1 + (1 - beta1) * g                         # :63
    beta2_t = beta2 ** step                                 # :71
    n_max = 2 / (1 - beta2) - 1                             # :72
    n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)      # :73
    if n_sma >= 5:                                          # :77-85
        step_size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) / (1 - beta1 ** step)
    elif degenerated_to_sgd:
        step_size = 1.0 / (1 - beta1 ** step)
    else:
        step_size = -1
    if n_sma >= 5:                                          # :92-97
        if weight_decay != 0:
            p = p + (-weight_decay * lr) * p
        p = p + (-step_size * lr) * (m / (v.sqrt() + eps))
    elif step_size > 0:                                     # :98-102
        if weight_decay != 0:
            p = p + (-weight_decay * lr) * p
        p = p + (-step_size * lr) * m
    return p, m, v


def run_steps(params: List[torch.Tensor], grads_per_step: List[List[torch.Tensor]], lr, weight_decay=0.0, max_norm=0.0,
              extra_sqnorm: Optional[Sequence[float]] = None, degenerated_to_sgd=False, betas=(0.9, 0.999), eps=1e-8):
    """Clip + RAdam over several steps.  Returns (params, exp_avg, exp_avg_sq) lists."""
    ps = [p.clone().float() for p in params]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    for t, grads in enumerate(grads_per_step, start=1):
        c = clip_coef(grads, max_norm, extra_sqnorm[t - 1] if extra_sqnorm is not None else 0.0)
        for i, g in enumerate(grads):
            ps[i], ms[i], vs[i] = radam_step(ps[i], g.float() * c, ms[i], vs[i], t, lr, betas, eps, weight_decay, degenerated_to_sgd)
    return ps, ms, vs
