"""BASELINE configs[0]: Qwen3-0.6B head shapes, 4 seqs, in=128, out=32, CPU oracle backend, tp=1.
Pass = runs end to end through real Req / Batch / page-table objects and matches itself for
page_size in {1, 16, 64}."""
import random

import numpy as np
import pytest
import torch

from oracle.cpu_backend import CpuPool, SDPACpuBackend
from plumbing import TinyModel, oracle_pre_attention, run_generation


def _run(b200, page_size, layers=2, n_seqs=4, in_len=128, out_len=8, shuffle=True):
    hq, hkv, d, hidden, vocab = 16, 8, 128, 256, 512
    rnd = random.Random(0)
    prompts = [[rnd.randrange(vocab) for _ in range(in_len)] for _ in range(n_seqs)]
    max_seq = 256
    num_pages = n_seqs * max_seq // page_size + 4
    ctx = b200.Context(page_size)
    ctx.page_table = torch.zeros((n_seqs + 1, max_seq), dtype=torch.int32)
    ctx.kv_cache = CpuPool(hkv, layers, d, num_pages + 1, page_size, torch.bfloat16)
    backend = SDPACpuBackend(ctx, hq, hkv, d)
    ctx.attn_backend = backend
    free = [p * page_size for p in range(num_pages)]
    if shuffle:
        random.Random(page_size).shuffle(free)
    model = TinyModel(hq, hkv, d, layers, hidden, vocab, torch.device("cpu"))
    pre = lambda m, l, qkv, batch: oracle_pre_attention(m, l, qkv, batch, backend)  # noqa: E731
    return run_generation(model, ctx, backend, b200.Req, b200.Batch, prompts, out_len, page_size, pre, free)


def test_cfg0_plumbing_self_consistent_across_page_sizes(b200):
    base_ids, base_h = _run(b200, 1)
    assert base_ids.shape == (4, 8)
    for ps in (16, 64):
        ids, h = _run(b200, ps)
        assert np.array_equal(ids, base_ids), f"page_size {ps} changed the generated tokens"
        assert torch.equal(h, base_h), "paging must not change a single bit of the CPU path"


def test_cfg0_page_permutation_invariance(b200):
    a_ids, a_h = _run(b200, 16, shuffle=True)
    b_ids, b_h = _run(b200, 16, shuffle=False)
    assert np.array_equal(a_ids, b_ids) and torch.equal(a_h, b_h)


@pytest.mark.slow
def test_cfg0_full_shape(b200):
    """The literal cfg0 numbers: 4 seqs, in=128, out=32."""
    ids, _ = _run(b200, 16, layers=2, out_len=32)
    assert ids.shape == (4, 32)
