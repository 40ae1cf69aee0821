"""Interface mirror of the data the attention boundary receives.

When the real ``minisgl`` package is importable its own ``Req`` / ``Batch`` / ``Context`` are used
untouched (the backend only duck-types the fields listed below).  On a box without the reference
(the GPU box, the tests, ``bench.py``) these stand-ins carry the same field names and invariants
(reference ``python/minisgl/core.py:29-136``):

* ``Req``:   ``table_idx``, ``cached_len``, ``device_len``, ``extend_len``, ``complete_one()``
* ``Batch``: ``reqs``, ``padded_reqs``, ``phase``, ``positions``, ``out_loc``, ``attn_metadata``
* ``Context``: ``page_size``, ``page_table`` (token-granular slots), ``kv_cache``, ``attn_backend``
"""

from __future__ import annotations

import contextlib
from typing import Any, List, Optional

import torch


class Req:
    """One running request as the scheduler hands it over (reference core.py:29-73)."""

    def __init__(
        self,
        *,
        table_idx: int,
        cached_len: int,
        device_len: int,
        max_device_len: Optional[int] = None,
        uid: int = 0,
    ) -> None:
        if not (0 <= cached_len < device_len):
            raise ValueError(f"need 0 <= cached_len < device_len, got {cached_len}, {device_len}")
        self.table_idx = table_idx
        self.cached_len = cached_len
        self.device_len = device_len
        self.max_device_len = device_len if max_device_len is None else max_device_len
        self.uid = uid

    @property
    def extend_len(self) -> int:
        return self.device_len - self.cached_len

    @property
    def remain_len(self) -> int:
        return self.max_device_len - self.device_len

    @property
    def can_decode(self) -> bool:
        return self.remain_len > 0

    def complete_one(self) -> None:
        """After a forward: everything on device is cached, one more slot is needed."""
        self.cached_len, self.device_len = self.device_len, self.device_len + 1

    def __repr__(self) -> str:  # pragma: no cover
        return f"Req(table_idx={self.table_idx}, cached={self.cached_len}, device={self.device_len})"


class Batch:
    """A forward's worth of requests (reference core.py:76-97)."""

    def __init__(self, reqs: List[Req], phase: str) -> None:
        if phase not in ("prefill", "decode"):
            raise ValueError(f"bad phase {phase!r}")
        self.reqs = reqs
        self.phase = phase
        self.padded_reqs: List[Req] = reqs
        self.input_ids: torch.Tensor
        self.positions: torch.Tensor
        self.out_loc: torch.Tensor
        self.attn_metadata: Any = None

    @property
    def is_prefill(self) -> bool:
        return self.phase == "prefill"

    @property
    def is_decode(self) -> bool:
        return self.phase == "decode"

    @property
    def size(self) -> int:
        return len(self.reqs)

    @property
    def padded_size(self) -> int:
        return len(self.padded_reqs)


class Context:
    """Process-wide state the layers reach through ``get_global_ctx()`` (reference core.py:100-122)."""

    def __init__(self, page_size: int) -> None:
        self.page_size = page_size
        self.page_table: torch.Tensor  # int32 [max_running_req + 1, align32(max_seq_len)], token slots
        self.kv_cache: Any = None
        self.attn_backend: Any = None
        self._batch: Optional[Batch] = None

    @property
    def batch(self) -> Batch:
        if self._batch is None:
            raise AssertionError("No active batch in context")
        return self._batch

    @contextlib.contextmanager
    def forward_batch(self, batch: Batch):
        if self._batch is not None:
            raise AssertionError("Nested forward_batch is not allowed")
        self._batch = batch
        try:
            yield
        finally:
            self._batch = None


_CTX: Optional[Context] = None


def set_global_ctx(ctx: Optional[Context]) -> None:
    """Unlike the reference (core.py:128-131) ``None`` is accepted so tests can reset the state."""
    global _CTX
    if ctx is not None and _CTX is not None:
        raise AssertionError("Global context is already set")
    _CTX = ctx


def get_global_ctx() -> Context:
    """Prefer the reference's live context when running inside mini-sglang."""
    if _CTX is not None:
        return _CTX
    try:  # pragma: no cover - only inside a real mini-sglang process
        from minisgl.core import get_global_ctx as _ref_get
    except ImportError:
        raise AssertionError("Global context is not set") from None
    return _ref_get()  # the reference's own assertion fires when its context is missing
