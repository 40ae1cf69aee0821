"""A tiny deterministic "model" around the attention boundary, used by the plumbing tests
(BASELINE configs[0]): L layers of  x -> qkv = x @ Wqkv -> AttentionLayer-sequence -> x += o @ Wo.
Weights are seeded randn (the reference's dummy-weight mode, engine/engine.py:37,140-144);
greedy "sampling" = argmax over a seeded lm-head.  Everything outside the attention boundary is
plain torch on the backend's device -- it is scaffolding, not the product."""
from __future__ import annotations

from typing import Callable, List

import numpy as np
import torch

from oracle import metadata as o_meta
from oracle import norm as o_norm
from oracle import rope as o_rope


class TinyModel:
    def __init__(self, hq, hkv, d, layers, hidden, vocab, device, dtype=torch.bfloat16, seed=42):
        g = torch.Generator().manual_seed(seed)
        mk = lambda *s: (torch.randn(*s, generator=g) / (s[0] ** 0.5)).to(dtype).to(device)  # noqa: E731
        self.hq, self.hkv, self.d, self.layers = hq, hkv, d, layers
        self.embed = (torch.randn(vocab, hidden, generator=g)).to(dtype).to(device)
        self.wqkv = [mk(hidden, (hq + 2 * hkv) * d) for _ in range(layers)]
        self.wo = [mk(hq * d, hidden) for _ in range(layers)]
        self.qw = [(torch.rand(d, generator=g) + 0.5).to(dtype).to(device) for _ in range(layers)]
        self.kw = [(torch.rand(d, generator=g) + 0.5).to(dtype).to(device) for _ in range(layers)]
        self.lm_head = mk(hidden, vocab)
        self.cos_sin = o_rope.ref_cos_sin_cache(d, 4096, 1e6).to(device)
        self.device, self.dtype = device, dtype


def run_generation(model: TinyModel, ctx, backend, make_req: Callable, make_batch: Callable,
                   prompts: List[List[int]], out_len: int, page_size: int, pre_attention: Callable,
                   free_pages: List[int], forced=None, hidden_trace=None):
    """Prefill all prompts in one batch, then `out_len - 1` decode steps. Returns generated ids
    [n_seqs, out_len] and the last-step hidden states (fp32) for numeric comparison."""
    dev = model.device
    n = len(prompts)
    table = ctx.page_table
    reqs = [make_req(table_idx=i, cached_len=0, device_len=len(p)) for i, p in enumerate(prompts)]
    tokens = [list(p) for p in prompts]
    generated = [[] for _ in range(n)]
    hidden_last = None
    for step in range(out_len):
        phase = "prefill" if step == 0 else "decode"
        batch = make_batch(reqs, phase)
        triples = [(r.table_idx, r.cached_len, r.device_len) for r in reqs]
        pt = table.cpu().numpy()
        o_meta.ref_allocate_paged(pt, free_pages, triples, page_size)
        table.copy_(torch.from_numpy(pt).to(table.device))
        batch.positions = torch.from_numpy(o_meta.ref_positions(triples)).to(dev)
        batch.out_loc = torch.from_numpy(o_meta.ref_out_loc(pt, triples)).to(dev)
        ids = torch.tensor([t for r, toks in zip(reqs, tokens) for t in toks[r.cached_len:r.device_len]],
                           dtype=torch.long, device=dev)
        backend.prepare_metadata(batch)
        x = model.embed[ids]
        with ctx.forward_batch(batch):
            for l in range(model.layers):
                qkv = (x @ model.wqkv[l]).contiguous()
                o = pre_attention(model, l, qkv, batch)  # norm + rope + backend.forward
                x = x + (o.reshape(x.shape[0], -1) @ model.wo[l])
        last = batch.attn_metadata.get_last_indices(n).long()
        h = x[last]
        logits = (h @ model.lm_head).float()
        nxt = logits.argmax(-1).tolist()
        hidden_last = h.float().cpu()
        if hidden_trace is not None:
            hidden_trace.append(hidden_last)
        if forced is not None:  # teacher forcing: follow another run's tokens, compare hidden states
            nxt = [int(t) for t in forced[:, step]]
        for i, r in enumerate(reqs):
            generated[i].append(nxt[i])
            tokens[i].append(nxt[i])
            r.complete_one()
    return np.array(generated), hidden_last


def oracle_pre_attention(model: TinyModel, l: int, qkv, batch, backend):
    from oracle.layer import ref_pre_attention

    q, k, v = ref_pre_attention(qkv, batch.positions, model.hq, model.hkv, model.d, model.cos_sin,
                                model.qw[l], model.kw[l], 1e-6)
    return backend.forward(q, k, v, l, batch)
