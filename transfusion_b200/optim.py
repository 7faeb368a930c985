"""Fused RAdam + global-norm gradient clipping for the fusion path's parameters.

Drop-in for the reference optimizer ``runner/metrics_losses/radam_optim.py:RAdam`` (same constructor arguments,
``param_groups`` and per-parameter ``state`` keys ``step`` / ``exp_avg`` / ``exp_avg_sq``, so optimizer state dicts
interchange) restricted to CUDA fp32 parameters, plus the clipping Lightning applies before the step
(``gradient_clip_val`` / ``gradient_clip_algorithm="norm"``, runner/run_experiment.py:445-446).  One
``xf_radam_step`` launch covers up to 32 tensors: a single HBM pass over (param, grad, exp_avg, exp_avg_sq) that
also emits the bf16 weight copy the next forward's tensor-core GEMMs read (``cross_fusion/level_fn.py`` caches those
copies per parameter version, so the per-step weight-cast launch disappears).  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch

from . import _lib
from ._lib import XfRAdam, XfRAdamJob, check, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _jobs(entries):
    arr = (XfRAdamJob * len(entries))()
    for j, (p, g, m, v, pb) in zip(arr, entries):
        j.param = p.data_ptr() if p is not None else None
        j.grad = g.data_ptr()
        j.exp_avg = m.data_ptr() if m is not None else None
        j.exp_avg_sq = v.data_ptr() if v is not None else None
        j.param_bf16 = pb.data_ptr() if pb is not None else None
        j.n = g.numel()
    return arr


def grad_sqnorm(grads: Iterable[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out (device fp32 scalar, created zeroed when None) += sum of squares of the given fp32 CUDA gradients."""
    grads = [g for g in grads if g is not None]
    if out is None:
        dev = grads[0].device if grads else "cuda"
        out = torch.zeros(1, device=dev, dtype=torch.float32)
    for g in grads:
        if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()):
            raise _lib.XfError("grad_sqnorm: contiguous fp32 CUDA gradients required")
    for i in range(0, len(grads), _lib.XF_OPT_MAX_JOBS):
        chunk = [(None, g, None, None, None) for g in grads[i:i + _lib.XF_OPT_MAX_JOBS]]
        check(lib().xf_grad_sqnorm(_jobs(chunk), len(chunk), C.c_void_p(out.data_ptr()), _stream()), "xf_grad_sqnorm")
    return out


class FusedRAdam(torch.optim.Optimizer):
    """radam_optim.py:6-104 (RAdam) over the C ABI.  ``max_grad_norm`` > 0 adds Lightning's global-norm clip: the squared
    norm of THIS optimizer's gradients is measured on the device and added to ``extra_sqnorm`` (a device scalar with the
    squared norm of every gradient clipped together with them but stepped elsewhere), no host synchronisation."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, degenerated_to_sgd=False,
                 max_grad_norm: float = 0.0):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        self.degenerated_to_sgd = degenerated_to_sgd
        self.max_grad_norm = float(max_grad_norm)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None, extra_sqnorm: Optional[torch.Tensor] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        sq = None
        if self.max_grad_norm > 0:
            sq = grad_sqnorm([p.grad for g in self.param_groups for p in g["params"] if p.grad is not None])
            if extra_sqnorm is not None:
                sq += extra_sqnorm.to(sq.dtype).reshape(1)
        for group in self.param_groups:
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.dtype == torch.float32
                        and p.grad.is_contiguous()):
                    raise _lib.XfError("FusedRAdam: contiguous fp32 CUDA parameters / gradients only (no CPU path)")
                st = self.state[p]
                if len(st) == 0:   # radam_optim.py:49-52
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1    # :65
                by_step.setdefault(st["step"], []).append(p)
            beta1, beta2 = group["betas"]
            for step, plist in by_step.items():
                # the parameter (and its bf16 copy) only moves when N_sma >= 5 or degenerated_to_sgd (radam_optim.py:89-102)
                b2t = beta2 ** step
                n_sma = (2.0 / (1.0 - beta2) - 1.0) - 2.0 * step * b2t / (1.0 - b2t)
                h_mode_moves = n_sma >= 5 or self.degenerated_to_sgd
                h = XfRAdam()
                h.lr_d, h.beta1_d, h.beta2_d, h.eps, h.weight_decay_d = group["lr"], beta1, beta2, group["eps"], group["weight_decay"]
                h.degenerated_to_sgd = int(self.degenerated_to_sgd)
                h.step = int(step)
                h.max_grad_norm = self.max_grad_norm
                h.grad_sqnorm = sq.data_ptr() if sq is not None else None
                for i in range(0, len(plist), _lib.XF_OPT_MAX_JOBS):
                    entries = []
                    for p in plist[i:i + _lib.XF_OPT_MAX_JOBS]:
                        st = self.state[p]
                        entries.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"], _bf16_sink(p)))
                    check(lib().xf_radam_step(_jobs(entries), len(entries), C.byref(h), _stream()), "xf_radam_step")
                for p in plist:   # the kernel wrote p in place behind autograd's back: bump the version, keep the bf16 copy current
                    torch.autograd.graph.increment_version(p)
                    cache = getattr(p, "_xf_bf16", None)
                    if cache is not None and cache[2] and _bf16_sink(p) is not None and h_mode_moves:
                        p._xf_bf16 = (p._version, cache[1], True, "opt")
        return loss


def _bf16_sink(p):
    """The cached bf16 copy of `p` kept by cross_fusion/level_fn.py when it has p's own flat layout (no head padding)."""
    cache = getattr(p, "_xf_bf16", None)
    if cache is None or not cache[2]:
        return None
    buf = cache[1]
    return buf if (buf.numel() == p.numel() and buf.device == p.device) else None
