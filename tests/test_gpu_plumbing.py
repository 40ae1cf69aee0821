"""GPU end-to-end plumbing (BASELINE configs[0] driven through the product path): the tiny model of
tests/plumbing.py generates tokens through `AttentionLayer` (fused qk-norm + RoPE launch) and
`B200AttnBackend` (prefill, then decode steps with page allocation, appends, growing KV) and must
reproduce the CPU oracle backend token for token."""
import random

import numpy as np
import pytest
import torch

from oracle.cpu_backend import CpuPool, SDPACpuBackend
from plumbing import TinyModel, oracle_pre_attention, run_generation

pytestmark = pytest.mark.gpu


def _setup(b200, page_size, device, layers, n_seqs, max_seq):
    hq, hkv, d = 16, 8, 128
    num_pages = n_seqs * max_seq // page_size + 4
    ctx = b200.Context(page_size)
    b200.core.set_global_ctx(None)
    b200.set_global_ctx(ctx)
    ctx.page_table = torch.zeros((n_seqs + 1, max_seq), dtype=torch.int32, device=device)
    free = [p * page_size for p in range(num_pages)]
    random.Random(page_size).shuffle(free)
    return ctx, num_pages, free, (hq, hkv, d)


@pytest.mark.parametrize("page_size", [1, 64])
def test_generation_matches_cpu_oracle_backend(b200, native_lib, page_size):
    layers, n_seqs, in_len, out_len, hidden, vocab, max_seq = 2, 4, 128, 12, 256, 512, 256
    rnd = random.Random(0)
    prompts = [[rnd.randrange(vocab) for _ in range(in_len - 7 * i)] for i in range(n_seqs)]

    # ---- CPU oracle backend
    ctx, num_pages, free, (hq, hkv, d) = _setup(b200, page_size, torch.device("cpu"), layers, n_seqs, max_seq)
    ctx.kv_cache = CpuPool(hkv, layers, d, num_pages + 1, page_size, torch.bfloat16)
    cpu_backend = SDPACpuBackend(ctx, hq, hkv, d)
    ctx.attn_backend = cpu_backend
    model = TinyModel(hq, hkv, d, layers, hidden, vocab, torch.device("cpu"))
    ref_trace = []
    ref_ids, ref_h = run_generation(model, ctx, cpu_backend, b200.Req, b200.Batch, prompts, out_len, page_size,
                                    lambda m, l, qkv, batch: oracle_pre_attention(m, l, qkv, batch, cpu_backend), free,
                                    hidden_trace=ref_trace)

    # ---- product path on the GPU
    dev = torch.device("cuda")
    ctx, num_pages, free, _ = _setup(b200, page_size, dev, layers, n_seqs, max_seq)
    ctx.kv_cache = b200.MHAKVCache(hkv, layers, d, num_pages + 1, page_size, torch.bfloat16, dev)
    ctx.kv_cache._kv_buffer.fill_(float("nan"))  # a real pool is torch.empty
    from types import SimpleNamespace

    backend = b200.create_attention_backend("b200", SimpleNamespace(num_qo_heads=hq, num_kv_heads=hkv, head_dim=d))
    ctx.attn_backend = backend
    gmodel = TinyModel(hq, hkv, d, layers, hidden, vocab, dev)
    rotary = b200.layers.RotaryEmbedding(d, d, 4096, 1e6, device=dev)
    attn_layers = []
    for l in range(layers):
        qn, kn = b200.layers.RMSNorm(d, 1e-6), b200.layers.RMSNorm(d, 1e-6)
        qn.weight, kn.weight = gmodel.qw[l], gmodel.kw[l]
        attn_layers.append(b200.layers.AttentionLayer(l, hq, hkv, d, rotary, qn, kn, fuse_pre_attention=True))

    def pre(m, l, qkv, batch):
        return attn_layers[l].forward(qkv)

    # teacher forcing with the CPU run's tokens: identical trajectories, so the hidden state of every
    # step (prefill + 11 decode steps, pages allocated as it grows) must agree to bf16 noise
    trace = []
    ids, h = run_generation(gmodel, ctx, backend, b200.Req, b200.Batch, prompts, out_len, page_size, pre, free,
                            forced=ref_ids, hidden_trace=trace)
    for step, (a, b) in enumerate(zip(trace, ref_trace)):
        err = (a - b).abs().max().item() / b.abs().max().item()
        assert err < 3e-2, f"step {step}: hidden states differ by {err:.3e}"
    assert (ids == ref_ids).mean() >= 0.9, (ids, ref_ids)  # greedy picks agree except on near ties
