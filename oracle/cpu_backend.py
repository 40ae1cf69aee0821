"""Oracle: a torch-SDPA ``BaseAttnBackend`` for CPU (TEST INFRASTRUCTURE ONLY).

BASELINE configs[0] ("torch-SDPA CPU backend, plumbing, no GPU").  The reference registers no
such backend (python/minisgl/attention/__init__.py:22-40); this one follows the interface
(attention/base.py:18-34) and the semantics of the fa backend (fa.py:49-105) with the oracle
functions, so the same driver can run the product backend on a GPU and this one on CPU.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from . import metadata as o_meta
from .attention import ref_paged_attention
from .store import ref_store_kv


@dataclass
class CpuMetadata:
    cu_seqlens_q: torch.Tensor
    cu_seqlens_k: torch.Tensor
    cache_seqlens: torch.Tensor
    max_seqlen_q: int
    max_seqlen_k: int
    page_table: torch.Tensor  # token-granular slots [bs, max_k]

    def get_last_indices(self, bs: int) -> torch.Tensor:
        return self.cu_seqlens_q[1 : 1 + bs] - 1


class SDPACpuBackend:
    def __init__(self, ctx, hq: int, hkv: int, head_dim: int) -> None:
        self.ctx, self.hq, self.hkv, self.d = ctx, hq, hkv, head_dim

    def prepare_metadata(self, batch) -> None:
        triples = [(r.table_idx, r.cached_len, r.device_len) for r in batch.padded_reqs]
        md = o_meta.ref_prepare_metadata(self.ctx.page_table.numpy(), triples, self.ctx.page_size)
        batch.attn_metadata = CpuMetadata(
            torch.from_numpy(md.cu_seqlens_q), torch.from_numpy(md.cu_seqlens_k),
            torch.from_numpy(md.cache_seqlens), md.max_seqlen_q, md.max_seqlen_k,
            torch.from_numpy(md.slot_table),
        )

    def forward(self, q, k, v, layer_id: int, batch) -> torch.Tensor:
        md = batch.attn_metadata
        pool = self.ctx.kv_cache
        kc = pool.k_cache(layer_id).view(-1, self.hkv, self.d)
        vc = pool.v_cache(layer_id).view(-1, self.hkv, self.d)
        nnz = q.shape[0]
        ref_store_kv(kc, vc, batch.out_loc, k.reshape(nnz, -1), v.reshape(nnz, -1))
        lens = md.cache_seqlens.tolist()
        rows = [md.page_table[i, :n] for i, n in enumerate(lens)]
        q_lens = np.diff(md.cu_seqlens_q.numpy()).tolist()
        return ref_paged_attention(q.reshape(nnz, self.hq, self.d), kc, vc, rows, q_lens)

    def init_capture_graph(self, max_seq_len: int, bs_list: List[int]) -> None:
        pass

    def prepare_for_capture(self, batch) -> None:
        self.prepare_metadata(batch)

    def prepare_for_replay(self, batch) -> None:
        pass


class CpuPool:
    """[2, L, pages, page, Hkv, D] on CPU -- layout of python/minisgl/kvcache/mha_pool.py:28-37."""

    def __init__(self, hkv, layers, d, num_pages, page_size, dtype):
        self._buf = torch.zeros((2, layers, num_pages, page_size, hkv, d), dtype=dtype)
        self.device, self.dtype, self.num_layers = torch.device("cpu"), dtype, layers

    def k_cache(self, i):
        return self._buf[0, i]

    def v_cache(self, i):
        return self._buf[1, i]
