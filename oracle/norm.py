"""Oracle: RMSNorm and fused residual-add + RMSNorm (TEST INFRASTRUCTURE ONLY).

Call sites: ``python/minisgl/layers/norm.py:16-21`` (``rmsnorm(x, w, eps[, out=x])``, also
the per-head q/k norm ``python/minisgl/layers/attention.py:50-53`` on ``[nnz, H, D]``) and
``norm.py:32-38`` (``fused_add_rmsnorm(x, residual, w, eps)``, both in place).
Arithmetic: FlashInfer ``norm.cuh`` (flashinfer-python 0.6.11.post2)
``RMSNormKernel`` / ``FusedAddRMSNormKernel``:

* ``rcp = rsqrt(mean(float(x)**2) + eps)``; ``y = float(x) * rcp * float(w)`` -> one rounding.
* fused: ``s = float(x) + float(residual)`` (fp32); ``residual <- round(s)``;
  ``rcp`` from the *unrounded* ``s``; ``x <- round(s * rcp * float(w))``.
"""

from __future__ import annotations

from typing import Tuple

import torch


def ref_rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """Normalises over the last dim (works for 2-D rows and the 3-D per-head variant)."""
    xf = x.float()
    rcp = torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (xf * rcp * w.float()).to(x.dtype)


def ref_fused_add_rmsnorm(
    x: torch.Tensor, residual: torch.Tensor, w: torch.Tensor, eps: float
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns ``(new_x, new_residual)`` (the reference mutates both in place)."""
    s = x.float() + residual.float()
    new_res = s.to(residual.dtype)
    rcp = torch.rsqrt(s.pow(2).mean(dim=-1, keepdim=True) + eps)
    return (s * rcp * w.float()).to(x.dtype), new_res
