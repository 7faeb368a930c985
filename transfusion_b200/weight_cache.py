"""bf16 copies of fp32 master weights for the tensor-core GEMMs, cached on the parameter.

A copy is stored as ``prm._xf_bf16 = (version, buffer, flat layout?, tag)`` and reused
  * in training only when ``transfusion_b200.optim.FusedRAdam`` produced it in its update pass (tag "opt": the optimizer
    owns parameter AND copy; one forward consumes it), so steady-state training casts nothing; any other optimizer may
    write through ``p.data`` without touching the version counter (the reference's RAdam does, radam_optim.py:96), so
    nothing else is trusted while training;
  * in inference while the parameter's version counter is unchanged (module.train() / eval() switches drop the cache;
    ``invalidate(module)`` does it on demand)."""
from __future__ import annotations

import torch


def bf16_weight(prm, rows: int, cols: int, casts: list, trust_version: bool, dev, pad=None):
    """Returns the bf16 [rows, cols] (or head-padded, pad = (rin, rout, cin, cout, out_rows, out_cols)) copy of `prm`,
    appending a cast job (ops.cast_pad_multi tuple) to `casts` when the cached copy cannot be used."""
    bf = torch.bfloat16
    cache = getattr(prm, "_xf_bf16", None)
    flat = pad is None
    ok = cache is not None and cache[0] == prm._version and cache[1].device == dev and cache[2] == flat
    if ok and (trust_version or cache[3] == "opt"):
        if cache[3] == "opt":
            prm._xf_bf16 = (cache[0], cache[1], cache[2], "used")
        return cache[1]
    if flat:
        if cache is not None and cache[2] and cache[1].device == dev and cache[1].numel() == rows * cols:
            buf = cache[1].view(rows, cols)
        else:
            buf = torch.empty(rows, cols, device=dev, dtype=bf)
        casts.append((prm.reshape(rows, cols), buf, rows, cols, 0, 0, 0, 0))
    else:
        rin, rout, cin, cout, orows, ocols = pad
        buf = torch.zeros(orows, ocols, device=dev, dtype=bf)
        casts.append((prm, buf, rows, cols, rin, rout, cin, cout))
    prm._xf_bf16 = (prm._version, buf, flat, "cast")
    return buf


def invalidate(module: torch.nn.Module):
    for p in module.parameters():
        if hasattr(p, "_xf_bf16"):
            del p._xf_bf16
