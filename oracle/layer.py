"""Oracle: the fixed pre-attention sequence of ``AttentionLayer.forward``
(``python/minisgl/layers/attention.py:47-57``)  (TEST INFRASTRUCTURE ONLY).

``qkv [nnz, (Hq+2Hkv)*D]`` -> split -> optional per-head q/k RMSNorm (in place) -> neox RoPE
(in place on q, k) -> ``backend.forward(q[nnz,Hq,D], k, v, layer_id, batch)`` -> ``[nnz, Hq*D]``.
"""

from __future__ import annotations

from typing import Optional, Sequence

import torch

from .attention import ref_backend_forward
from .norm import ref_rmsnorm
from .rope import ref_apply_rope_neox


def ref_pre_attention(
    qkv: torch.Tensor,
    positions: torch.Tensor,
    hq: int,
    hkv: int,
    d: int,
    cos_sin_cache: torch.Tensor,
    q_norm_w: Optional[torch.Tensor] = None,
    k_norm_w: Optional[torch.Tensor] = None,
    eps: float = 1e-6,
):
    """Returns new ``(q [nnz,Hq,D], k [nnz,Hkv*D], v [nnz,Hkv*D])`` after norm + rope."""
    nnz = qkv.shape[0]
    q, k, v = qkv.split([hq * d, hkv * d, hkv * d], dim=-1)
    q = q.reshape(nnz, hq, d)
    k = k.reshape(nnz, hkv, d)
    if q_norm_w is not None:
        q = ref_rmsnorm(q, q_norm_w, eps)
    if k_norm_w is not None:
        k = ref_rmsnorm(k, k_norm_w, eps)
    q = ref_apply_rope_neox(positions, q, d, cos_sin_cache)
    k = ref_apply_rope_neox(positions, k, d, cos_sin_cache)
    return q, k.reshape(nnz, hkv * d), v.contiguous()


def ref_attention_layer(
    qkv: torch.Tensor,
    positions: torch.Tensor,
    hq: int,
    hkv: int,
    d: int,
    cos_sin_cache: torch.Tensor,
    k_cache: torch.Tensor,
    v_cache: torch.Tensor,
    out_loc: torch.Tensor,
    slot_rows: Sequence[torch.Tensor],
    q_lens: Sequence[int],
    q_norm_w: Optional[torch.Tensor] = None,
    k_norm_w: Optional[torch.Tensor] = None,
    eps: float = 1e-6,
) -> torch.Tensor:
    q, k, v = ref_pre_attention(qkv, positions, hq, hkv, d, cos_sin_cache, q_norm_w, k_norm_w, eps)
    o = ref_backend_forward(q, k, v, k_cache, v_cache, out_loc, slot_rows, q_lens)
    return o.reshape(qkv.shape[0], hq * d)
