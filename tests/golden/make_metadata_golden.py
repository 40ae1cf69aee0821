"""Generate golden vectors for the INTEGER side of the attention boundary by importing the
Python reference in the build container (it cannot travel to the GPU box).

    PYTHONPATH=/root/reference/python python tests/golden/make_metadata_golden.py

What runs is the reference's own code, unmodified:
  * CacheManager.allocate_paged            (python/minisgl/scheduler/cache.py:42-53,127-146)
  * _make_positions / _make_input_tuple    (python/minisgl/scheduler/scheduler.py:236-259)
  * FlashAttentionBackend.prepare_metadata (python/minisgl/attention/fa.py:67-105)
  * FlashInferBackend.prepare_metadata     (python/minisgl/attention/fi.py:190-225)
  * get_last_indices                       (fa.py:32-33)
Only environment shims are applied, none of them touches arithmetic: `pin_memory=True` is dropped
(no CUDA driver here), the backends are built with object.__new__ (their __init__ needs a GPU /
flashinfer workspaces) and FIMetadata.__post_init__'s is_cuda assertions are skipped.
Output: tests/golden/metadata_golden.json (committed).
"""
import json
import random
import sys
from pathlib import Path
from types import SimpleNamespace

import torch

_real_tensor, _real_arange, _real_empty, _real_ones = torch.tensor, torch.arange, torch.empty, torch.ones


def _strip(fn):
    def wrapped(*a, **kw):
        kw.pop("pin_memory", None)
        return fn(*a, **kw)
    return wrapped


torch.tensor, torch.arange, torch.empty, torch.ones = map(_strip, (_real_tensor, _real_arange, _real_empty, _real_ones))

from minisgl import core  # noqa: E402
from minisgl.attention import fa, fi  # noqa: E402
from minisgl.scheduler import cache as sched_cache  # noqa: E402
from minisgl.scheduler import scheduler as sched  # noqa: E402

fi.FIMetadata.__post_init__ = lambda self: None


def build_case(name, page_size, lens, phase, pad_to=None, seed=0, shuffle=True):
    """lens: list of (cached_len, device_len). Pages for [0, cached) are allocated by a first
    allocate_paged call on Req(cached_len=0, device_len=cached) objects, the rest by the real call."""
    rnd = random.Random(seed)
    max_reqs = max(len(lens), pad_to or 0)
    max_seq = (max(d for _, d in lens) + 63) // 64 * 64  # whole pages are written
    num_pages = sum((d + page_size - 1) // page_size for _, d in lens) + 4
    page_table = torch.zeros((max_reqs + 1, max_seq), dtype=torch.int32)
    core._GLOBAL_CTX = None
    ctx = core.Context(page_size)
    ctx.page_table = page_table
    ctx.kv_cache = SimpleNamespace(device=torch.device("cpu"), dtype=torch.bfloat16)
    core.set_global_ctx(ctx)
    cm = sched_cache.CacheManager(num_pages, page_size, page_table, "naive")
    if shuffle:
        perm = list(range(num_pages))
        rnd.shuffle(perm)
        cm.free_slots = torch.tensor(perm, dtype=torch.int32) * page_size
    free0 = cm.free_slots.tolist()
    dummy_slot = num_pages * page_size
    page_table[max_reqs].fill_(dummy_slot)

    def mk(t, c, d):
        return core.Req(input_ids=torch.zeros(d, dtype=torch.int32), table_idx=t, cached_len=c,
                        output_len=8, uid=t, sampling_params=None, cache_handle=None)

    # earlier forwards: allocate the cached part
    pre = [mk(t, 0, c) for t, (c, d) in enumerate(lens) if c > 0]
    if pre:
        cm.allocate_paged(pre)
    reqs = [mk(t, c, d) for t, (c, d) in enumerate(lens)]
    cm.allocate_paged(reqs)
    batch = core.Batch(reqs=reqs, phase=phase)
    dummy = core.Req(input_ids=torch.tensor([0], dtype=torch.int32), table_idx=max_reqs, cached_len=0,
                     output_len=1, uid=-1, sampling_params=None, cache_handle=None)
    batch.padded_reqs = reqs + [dummy] * ((pad_to or len(reqs)) - len(reqs))
    dev = torch.device("cpu")
    batch.positions = sched._make_positions(batch, dev)
    input_mapping = sched._make_input_tuple(batch, dev)
    batch.out_loc = page_table[input_mapping]

    fa_b = object.__new__(fa.FlashAttentionBackend)
    fa_b.kvcache, fa_b.page_size = ctx.kv_cache, page_size
    fa_b.prepare_metadata(batch)
    m = batch.attn_metadata
    bs = len(batch.padded_reqs)
    out = {
        "name": name, "page_size": page_size, "phase": phase,
        "reqs": [[r.table_idx, r.cached_len, r.device_len] for r in batch.padded_reqs],
        "free_slots_before": free0,
        "page_table": page_table.tolist(),
        "positions": batch.positions.tolist(),
        "out_loc": batch.out_loc.tolist(),
        "fa_cu_seqlens_q": m.cu_seqlens_q.tolist(),
        "fa_cu_seqlens_k": m.cu_seqlens_k.tolist(),
        "fa_cache_seqlens": m.cache_seqlens.tolist(),
        "fa_max_seqlen_q": m.max_seqlen_q, "fa_max_seqlen_k": m.max_seqlen_k,
        "fa_page_table": m.page_table.tolist(),
        "last_indices": m.get_last_indices(bs).tolist(),
    }
    fi_b = object.__new__(fi.FlashInferBackend)
    fi_b.device, fi_b.kvcache = dev, ctx.kv_cache
    fi_b.qo_head_local, fi_b.kv_head_local = 16, 8
    fi_b.config = SimpleNamespace(head_dim=128)
    fi_b.cached_ones_cpu = torch.tensor([], dtype=torch.int32)
    fi_b.decode_wrappers = fi_b.prefill_wrapper = None
    fi_b.prepare_metadata(batch)
    m = batch.attn_metadata
    out.update({
        "fi_indices": m.indices.tolist(),
        "fi_cu_seqlens_q": m.cu_seqlens_q_cpu.tolist(),
        "fi_cu_seqlens_k": m.cu_seqlens_k_cpu.tolist(),
        "fi_seq_lens": m.seq_lens_cpu.tolist(),
        "fi_last_indices": m.get_last_indices(bs).tolist(),
    })
    return out


def main():
    rnd = random.Random(1)
    cases = []
    for ps in (1, 16, 64):
        cases.append(build_case(f"decode_ps{ps}", ps, [(n - 1, n) for n in (1, 2, 63, 64, 65, 200, 513)], "decode"))
        cases.append(build_case(f"decode_pad_ps{ps}", ps, [(99, 100), (300, 301), (40, 41)], "decode", pad_to=8))
        cases.append(build_case(f"prefill_nocache_ps{ps}", ps, [(0, n) for n in (1, 17, 64, 65, 333)], "prefill"))
        al = lambda c: (c // ps) * ps  # noqa: E731  radix hits are page aligned
        cases.append(build_case(f"extend_ps{ps}", ps, [(al(64), 200), (al(128), al(128) + 1), (0, 50), (al(320), 700)], "prefill"))
        lens = []
        for _ in range(12):
            d = rnd.randint(2, 400)
            lens.append((al(rnd.randint(0, d - 1)), d))
        cases.append(build_case(f"random_ps{ps}", ps, lens, "prefill", seed=ps))
    path = Path(__file__).with_name("metadata_golden.json")
    path.write_text(json.dumps({"generator": "tests/golden/make_metadata_golden.py",
                                "reference": "sgl-project/mini-sglang @ 20fcd7f", "cases": cases}))
    print(f"wrote {path} ({len(cases)} cases, {path.stat().st_size/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
