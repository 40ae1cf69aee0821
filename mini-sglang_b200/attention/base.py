"""The reference's attention-backend interface (``python/minisgl/attention/base.py:12-63``).

Inside a mini-sglang process the reference's own ABCs are re-exported, so ``B200AttnBackend`` is a
genuine ``BaseAttnBackend`` subclass there; elsewhere an interface mirror with the same five
abstract methods is defined.
"""

from __future__ import annotations

import abc
from typing import TYPE_CHECKING, List

if TYPE_CHECKING:  # pragma: no cover
    import torch

try:  # pragma: no cover - only when the reference is on sys.path
    from minisgl.attention.base import BaseAttnBackend, BaseAttnMetadata, HybridBackend

    USING_REFERENCE_ABCS = True
except ImportError:
    USING_REFERENCE_ABCS = False

    class BaseAttnMetadata(abc.ABC):
        """Per-batch metadata; ``get_last_indices`` feeds the LM head in prefill
        (reference layers/embedding.py:92-94)."""

        @abc.abstractmethod
        def get_last_indices(self, bs: int) -> "torch.Tensor": ...

    class BaseAttnBackend(abc.ABC):
        @abc.abstractmethod
        def forward(self, q, k, v, layer_id: int, batch) -> "torch.Tensor":
            """Append k, v at ``batch.out_loc`` and return causal attention ``[nnz, Hq, D]``."""

        @abc.abstractmethod
        def prepare_metadata(self, batch) -> None:
            """Set ``batch.attn_metadata`` (runs on the scheduler stream, one step ahead)."""

        @abc.abstractmethod
        def init_capture_graph(self, max_seq_len: int, bs_list: List[int]) -> None: ...

        @abc.abstractmethod
        def prepare_for_capture(self, batch) -> None: ...

        @abc.abstractmethod
        def prepare_for_replay(self, batch) -> None: ...

    class HybridBackend(BaseAttnBackend):
        """``--attn p,d``: one backend for prefill, another for decode; the CUDA-graph hooks go to
        the decode backend only (reference base.py:37-63)."""

        def __init__(self, prefill_backend: BaseAttnBackend, decode_backend: BaseAttnBackend):
            self.prefill_backend = prefill_backend
            self.decode_backend = decode_backend

        def _pick(self, batch) -> BaseAttnBackend:
            return self.prefill_backend if batch.is_prefill else self.decode_backend

        def forward(self, q, k, v, layer_id, batch):
            return self._pick(batch).forward(q, k, v, layer_id, batch)

        def prepare_metadata(self, batch) -> None:
            self._pick(batch).prepare_metadata(batch)

        def init_capture_graph(self, max_seq_len, bs_list) -> None:
            self.decode_backend.init_capture_graph(max_seq_len, bs_list)

        def prepare_for_capture(self, batch) -> None:
            self.decode_backend.prepare_for_capture(batch)

        def prepare_for_replay(self, batch) -> None:
            self.decode_backend.prepare_for_replay(batch)


__all__ = ["BaseAttnBackend", "BaseAttnMetadata", "HybridBackend", "USING_REFERENCE_ABCS"]
