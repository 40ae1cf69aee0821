"""Oracle: KV-append row scatter (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Restates ``store_kv_cache`` (``python/minisgl/kernel/csrc/jit/store.cu:41-50``):
for every token ``t``: ``k_cache[indices[t]] = k[t]`` and ``v_cache[indices[t]] = v[t]``,
whole rows, byte for byte.  Called by ``MHAKVCache.store_kv``
(``python/minisgl/kvcache/mha_pool.py:45-56``).  Duplicate indices (padded dummy
requests, ``python/minisgl/engine/engine.py:98``) all carry the same slot; like the
reference the result for a duplicated slot is "one of the writers".
"""

from __future__ import annotations

import numpy as np
import torch


def _as_bytes_2d(t: torch.Tensor) -> np.ndarray:
    """View a 2-D row-strided tensor as uint8 rows (no copy for contiguous rows)."""
    assert t.dim() == 2 and t.stride(1) == 1
    rows, cols = t.shape
    flat = t.contiguous().view(torch.uint8) if t.dtype != torch.uint8 else t.contiguous()
    return flat.numpy().reshape(rows, cols * t.element_size())


def ref_store_kv(
    k_cache: torch.Tensor,
    v_cache: torch.Tensor,
    indices: torch.Tensor,
    k: torch.Tensor,
    v: torch.Tensor,
) -> None:
    """In-place scatter on CPU tensors. ``k_cache``/``v_cache``: ``[slots, ...]``;
    ``k``/``v``: ``[n, row]`` (may be row-strided views); ``indices``: int32/int64 ``[n]``."""
    assert k_cache.device.type == "cpu"
    n = indices.numel()
    slots = k_cache.shape[0]
    kc = k_cache.view(slots, -1)
    vc = v_cache.view(slots, -1)
    idx = indices.to(torch.int64)
    assert k.shape[0] == n and v.shape[0] == n
    assert kc.shape[1] == k.reshape(n, -1).shape[1]
    # byte-exact: plain assignment of same-dtype rows never changes a bit
    kc[idx] = k.reshape(n, -1)
    vc[idx] = v.reshape(n, -1)


def ref_store_kv_bytes(
    cache_bytes: np.ndarray, indices: np.ndarray, rows_bytes: np.ndarray
) -> np.ndarray:
    """Pure-numpy byte-level variant used to cross-check ``ref_store_kv`` itself."""
    out = cache_bytes.copy()
    for t in range(indices.shape[0]):
        out[int(indices[t])] = rows_bytes[t]
    return out
