"""Oracle: embedding / row gather (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Restates ``index_kernel`` / ``masked_index_kernel`` (``python/minisgl/kernel/csrc/jit/index.cu:34-96``),
reached through ``indexing(weights, indices, output=, vocab_range=)``
(``python/minisgl/kernel/index.py:32-53``) from ``VocabParallelEmbedding.forward``
(``python/minisgl/layers/embedding.py:31-41``):

* plain:  ``out[t] = weights[indices[t]]`` whole rows, byte for byte;
* masked: ``pos = indices[t] - start`` compared as an UNSIGNED value with ``length``
  (``index.cu:82-91``: ``pos`` is ``size_t``, so negative differences wrap and fail the test):
  ``out[t] = weights[pos]`` if ``pos < length`` else a zero row.

Pinned against the reference's own known-answer function ``ref_indexing``
(``tests/kernel/test_index.py:13-29`` = ``F.embedding`` + mask) in ``tests/test_oracle_golden.py``.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def ref_indexing(
    weights: torch.Tensor, indices: torch.Tensor, vocab_range: Optional[Tuple[int, int]] = None
) -> torch.Tensor:
    """CPU tensors in, fresh ``[n, dim]`` tensor of the weights' dtype out (bit-exact rows)."""
    assert weights.device.type == "cpu" and weights.dim() == 2
    n = indices.numel()
    if n == 0:
        return weights.new_empty(0, weights.shape[1])
    es = weights.element_size()
    w_bytes = weights.contiguous().view(torch.uint8).numpy().reshape(weights.shape[0], weights.shape[1] * es)
    idx = indices.to(torch.int64).numpy()
    out = np.zeros((n, w_bytes.shape[1]), dtype=np.uint8)
    if vocab_range is None:
        out[:] = w_bytes[idx]
    else:
        start, length = int(vocab_range[0]), int(vocab_range[1])
        pos = (idx - start).astype(np.uint64)  # size_t arithmetic of the kernel
        keep = pos < np.uint64(length)
        out[keep] = w_bytes[pos[keep].astype(np.int64)]
    return torch.from_numpy(out).view(weights.dtype).reshape(n, weights.shape[1])
