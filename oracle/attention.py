"""Oracle: paged attention on CPU with torch SDPA (TEST INFRASTRUCTURE ONLY).

Restates what every reference backend's ``forward`` computes
(``python/minisgl/attention/fi.py:176-188``, ``fa.py:49-65``, ``trtllm.py:49-89``):

1. append: ``k_cache[out_loc[t]] = k[t]``, ``v_cache[out_loc[t]] = v[t]``
   (``python/minisgl/kvcache/mha_pool.py:45-56``);
2. for each request ``r`` with ``q_len = extend_len`` query rows and ``kv_len = device_len``
   cached tokens (slots ``page_table[table_idx, :kv_len]``):
   ``O = softmax(Q K^T * D**-0.5 + causal) V`` with the causal mask bottom-right aligned
   (query row ``i`` sees kv positions ``<= kv_len - q_len + i``; FlashInfer ``causal=True``
   ``fi.py:164``, FA ``causal=True`` ``fa.py:154``), GQA: q head ``h`` uses kv head
   ``h // (Hq // Hkv)``; no window, no softcap, ``pos_encoding_mode="NONE"`` (``fi.py:218``).

Math is fp32 over the stored (bf16/fp16) values; the result is rounded once to the
input dtype -- the tolerance the product is held to is 1e-3 relative (north_star).
"""

from __future__ import annotations

from typing import Sequence

import torch
import torch.nn.functional as F

from .store import ref_store_kv


def ref_attention_one(
    q: torch.Tensor,  # [q_len, Hq, D]
    k: torch.Tensor,  # [kv_len, Hkv, D]
    v: torch.Tensor,  # [kv_len, Hkv, D]
    scale: float,
) -> torch.Tensor:
    """Single request, dense K/V already gathered. Returns fp32 ``[q_len, Hq, D]``."""
    q_len, hq, d = q.shape
    kv_len, hkv, _ = k.shape
    g = hq // hkv
    # GQA without materialising repeated K/V: fold the group into the query-row axis,
    # q -> [Hkv, g*q_len, D] against k, v -> [Hkv, kv_len, D]
    qf = q.float().reshape(q_len, hkv, g, d).permute(1, 2, 0, 3).reshape(hkv, g * q_len, d)
    kf = k.float().permute(1, 0, 2)  # [Hkv, kv, D]
    vf = v.float().permute(1, 0, 2)
    if q_len == 1:
        mask = None  # the single (last) query row sees every key
    else:
        # bottom-right aligned causal mask, repeated for the g heads of a group
        qi = torch.arange(q_len).unsqueeze(1) + (kv_len - q_len)
        ki = torch.arange(kv_len).unsqueeze(0)
        mask = (ki <= qi).repeat(g, 1)  # [g*q, kv] True = attend
    o = F.scaled_dot_product_attention(qf, kf, vf, attn_mask=mask, scale=scale)  # [Hkv, g*q, D]
    return o.reshape(hkv, g, q_len, d).permute(2, 0, 1, 3).reshape(q_len, hq, d).contiguous()


def ref_paged_attention(
    q: torch.Tensor,  # [nnz, Hq, D]
    k_cache: torch.Tensor,  # [slots, Hkv, D]  (one layer, token-granular view)
    v_cache: torch.Tensor,
    slot_rows: Sequence[torch.Tensor],  # per request: int tensor [kv_len] of slots
    q_lens: Sequence[int],
    scale: float | None = None,
    exact: bool = False,
) -> torch.Tensor:
    """Ragged batch. Query rows are the concatenation over requests (scheduler.py:236-259).
    `exact`: return the unrounded fp32 result (what oracle/tolerance.vs_exact_oracle expects) instead of
    rounding it once to q's dtype like a backend does."""
    nnz, hq, d = q.shape
    scale = d**-0.5 if scale is None else scale
    out = torch.empty((nnz, hq, d), dtype=torch.float32)
    off = 0
    for rows, ql in zip(slot_rows, q_lens):
        idx = rows.to(torch.int64)
        o = ref_attention_one(q[off : off + ql], k_cache[idx], v_cache[idx], scale)
        out[off : off + ql] = o
        off += ql
    assert off == nnz
    return out if exact else out.to(q.dtype)


def ref_backend_forward(
    q: torch.Tensor,  # [nnz, Hq, D]
    k: torch.Tensor,  # [nnz, Hkv*D] (may be a strided view)
    v: torch.Tensor,
    k_cache: torch.Tensor,  # [slots, Hkv, D], modified in place
    v_cache: torch.Tensor,
    out_loc: torch.Tensor,  # [nnz]
    slot_rows: Sequence[torch.Tensor],
    q_lens: Sequence[int],
    scale: float | None = None,
) -> torch.Tensor:
    """``store_kv`` then attention, exactly the order of fi.py:185-188."""
    ref_store_kv(k_cache, v_cache, out_loc, k, v)
    return ref_paged_attention(q, k_cache, v_cache, slot_rows, q_lens, scale)


def attention_flops(q_lens: Sequence[int], kv_lens: Sequence[int], hq: int, d: int) -> int:
    """Exact causal flop count, SURVEY.md section 8(d): 4*Hq*D*sum(q*cached + q(q+1)/2)."""
    tot = 0
    for ql, kl in zip(q_lens, kv_lens):
        cached = kl - ql
        tot += ql * cached + ql * (ql + 1) // 2
    return 4 * hq * d * tot
