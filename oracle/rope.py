"""Oracle: rotary position embedding (TEST INFRASTRUCTURE ONLY).

* cache build: ``python/minisgl/layers/rotary.py:12-32`` -- ``inv_freq = base**(-2i/D)``
  (optionally post-processed: llama3 ``rotary.py:69-91``, yarn ``rotary.py:93-112``),
  ``cache[p] = [cos(p*f) | sin(p*f)]`` fp32 ``[max_pos, D]``.
* apply: FlashInfer ``apply_rope_with_cos_sin_cache_inplace`` (call site
  ``rotary.py:45-51``; arithmetic ``flashinfer/pos_enc.cuh`` ``vec_apply_llama_rope_cos_sin``,
  flashinfer-python 0.6.11.post2): neox (non-interleaved) layout, fp32
  ``x' = x*cos + rot(x)*sin`` with ``rot(x)[i] = -x[i+D/2]`` for ``i < D/2`` else ``x[i-D/2]``,
  cos/sin index ``i mod D/2``; one rounding to the storage dtype; in place on q and k.
"""

from __future__ import annotations

import math
from typing import Any, Callable, Dict, Optional

import torch


def ref_inv_freq(rotary_dim: int, base: float) -> torch.Tensor:
    return 1.0 / (base ** (torch.arange(0, rotary_dim, 2, dtype=torch.float) / rotary_dim))


def ref_llama3_post(
    inv_freq: torch.Tensor,
    factor: float,
    low_freq_factor: float,
    high_freq_factor: float,
    original_max_position: int,
) -> torch.Tensor:
    """rotary.py:69-91."""
    wave_len = 2 * math.pi / inv_freq
    if low_freq_factor == high_freq_factor:
        return torch.where(
            wave_len < original_max_position / high_freq_factor, inv_freq, inv_freq / factor
        )
    delta = high_freq_factor - low_freq_factor
    smooth = (original_max_position / wave_len - low_freq_factor) / delta
    smooth = torch.clamp(smooth, 0, 1)
    return ((1 - smooth) / factor + smooth) * inv_freq


def ref_yarn_post(
    inv_freq: torch.Tensor,
    rotary_dim: int,
    base: float,
    factor: float,
    orig_max_pos: int,
    beta_fast: float = 32.0,
    beta_slow: float = 1.0,
) -> torch.Tensor:
    """rotary.py:93-112."""

    def corr(num_rot: float) -> float:
        return rotary_dim * math.log(orig_max_pos / (num_rot * 2 * math.pi)) / (2 * math.log(base))

    low = max(math.floor(corr(beta_fast)), 0)
    high = min(math.ceil(corr(beta_slow)), rotary_dim // 2 - 1)
    ramp = torch.clamp(
        (torch.arange(rotary_dim // 2, dtype=torch.float32) - low) / max(high - low, 1), 0, 1
    )
    return (inv_freq / factor) * ramp + inv_freq * (1 - ramp)


def ref_cos_sin_cache(
    rotary_dim: int,
    max_position: int,
    base: float,
    rope_scaling: Optional[Dict[str, Any]] = None,
) -> torch.Tensor:
    """fp32 ``[max_position, rotary_dim]`` = cos | sin (rotary.py:24-32)."""
    inv_freq = ref_inv_freq(rotary_dim, base)
    if rope_scaling is not None:
        kind = rope_scaling["rope_type"]
        if kind == "llama3":
            inv_freq = ref_llama3_post(
                inv_freq,
                rope_scaling["factor"],
                rope_scaling["low_freq_factor"],
                rope_scaling["high_freq_factor"],
                rope_scaling["original_max_position_embeddings"],
            )
        elif kind == "yarn":
            inv_freq = ref_yarn_post(
                inv_freq,
                rotary_dim,
                base,
                rope_scaling["factor"],
                rope_scaling["original_max_position_embeddings"],
                rope_scaling.get("beta_fast", 32.0),
                rope_scaling.get("beta_slow", 1.0),
            )
        elif kind != "default":
            raise ValueError(f"Unsupported rope_scaling = {rope_scaling}")
    t = torch.arange(max_position, dtype=torch.float)
    freqs = torch.einsum("i,j -> ij", t, inv_freq)
    return torch.cat((freqs.cos(), freqs.sin()), dim=-1)


def ref_apply_rope_neox(
    positions: torch.Tensor,  # int [nnz]
    x: torch.Tensor,  # [nnz, H*D] or [nnz, H, D]; returns a new tensor, same dtype
    head_size: int,
    cos_sin_cache: torch.Tensor,
) -> torch.Tensor:
    nnz = positions.numel()
    xs = x.reshape(nnz, -1, head_size).float()
    half = head_size // 2
    cs = cos_sin_cache[positions.to(torch.int64)]  # [nnz, D]
    cos = cs[:, :half].unsqueeze(1)  # [nnz, 1, D/2]
    sin = cs[:, half:].unsqueeze(1)
    x1, x2 = xs[..., :half], xs[..., half:]
    o1 = x1 * cos + (-x2) * sin
    o2 = x2 * cos + x1 * sin
    return torch.cat((o1, o2), dim=-1).to(x.dtype).reshape(x.shape)
