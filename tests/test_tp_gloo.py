"""N > 1 path on CPU: world_size-2 gloo.  The attention path shards by heads exactly like the
reference's tensor parallelism (python/minisgl/layers/attention.py:33-34): each rank owns
Hq/tp query heads and max(1, Hkv/tp) kv heads of every layer, sees identical metadata, needs no
collective inside attention; the only exchange is the all-reduce after o_proj
(python/minisgl/layers/linear.py:102-106).  Here: two ranks run the oracle attention on their head
shard, all-reduce the o_proj partial sums over gloo and must reproduce the single-rank result."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.attention import ref_paged_attention


def _inputs():
    g = torch.Generator().manual_seed(0)
    hq, hkv, d, hidden = 8, 4, 128, 64
    lens = [5, 33, 70]
    slots = sum(lens)
    kc = torch.randn(slots, hkv, d, generator=g).to(torch.bfloat16)
    vc = torch.randn(slots, hkv, d, generator=g).to(torch.bfloat16)
    q = torch.randn(len(lens), hq, d, generator=g).to(torch.bfloat16)
    wo = (torch.randn(hq * d, hidden, generator=g) / 32).to(torch.bfloat16)
    rows, off = [], 0
    for n in lens:
        rows.append(torch.arange(off, off + n))
        off += n
    return hq, hkv, d, kc, vc, q, wo, rows


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib

    b200 = importlib.import_module("mini-sglang_b200")
    b200.utils.set_tp_info(rank, world)
    hq, hkv, d, kc, vc, q, wo, rows = _inputs()
    hq_l = b200.utils.div_even(hq, world)
    hkv_l = b200.utils.div_even(hkv, world, allow_replicate=True)
    # head shard of this rank: contiguous slices, kv heads follow their q heads
    qs = q[:, rank * hq_l : (rank + 1) * hq_l]
    ks = kc[:, rank * hkv_l : (rank + 1) * hkv_l]
    vs = vc[:, rank * hkv_l : (rank + 1) * hkv_l]
    o = ref_paged_attention(qs, ks, vs, rows, [1] * len(rows)).float()
    partial = o.reshape(len(rows), -1) @ wo[rank * hq_l * d : (rank + 1) * hq_l * d].float()
    dist.all_reduce(partial)  # sum over TP ranks
    if rank == 0:
        out.put(partial)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_tp2_head_sharding_matches_single_rank():
    hq, hkv, d, kc, vc, q, wo, rows = _inputs()
    full = ref_paged_attention(q, kc, vc, rows, [1] * len(rows)).float().reshape(len(rows), -1) @ wo.float()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    # heads are independent: the only difference is the order of the fp32 sum over head shards
    assert torch.allclose(got, full, rtol=1e-4, atol=1e-3), (got - full).abs().max()


def test_kv_head_replication_when_tp_exceeds_kv_heads():
    import importlib

    b200 = importlib.import_module("mini-sglang_b200")
    assert b200.utils.div_even(8, 8, allow_replicate=True) == 1  # tp = Hkv: one kv head per rank
    assert b200.utils.div_even(4, 8, allow_replicate=True) == 1  # tp > Hkv: replicated
    with pytest.raises(AssertionError):
        b200.utils.div_even(3, 8, allow_replicate=True)


# ------------------------------------------------------------------ all-reduce plug-in: host-side routing
class _FakeComm:
    """Stands in for B200AllReduce (which needs GPUs + CUDA IPC): same `fits` rule, counts its calls."""

    def __init__(self, max_bytes):
        self.max_bytes, self.calls = max_bytes, 0

    def fits(self, x):
        return x.dtype in (torch.bfloat16, torch.float16) and x.shape[-1] % 8 == 0 and x.numel() * x.element_size() <= self.max_bytes

    def all_reduce(self, x):
        self.calls += 1
        dist.all_reduce(x)
        return x


def _plugin_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib

    d_mod = importlib.import_module("mini-sglang_b200.distributed")

    class Fallback:  # the plug-in that was active before (the reference's NCCL path; gloo here)
        calls = 0

        def all_reduce(self, x):
            Fallback.calls += 1
            dist.all_reduce(x)
            return x

        def all_gather(self, x):
            parts = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(parts, x)
            return torch.cat(parts)

    comm = _FakeComm(max_bytes=64 * 1024)
    impl = d_mod.B200DistributedImpl(comm, Fallback())
    small = torch.full((8, 1024), float(rank + 1), dtype=torch.bfloat16)         # 16 KB: decode-sized -> ours
    big = torch.full((128, 1024), float(rank + 1), dtype=torch.bfloat16)         # 256 KB: prefill-sized -> fallback
    odd = torch.full((8, 1023), float(rank + 1), dtype=torch.bfloat16)           # row not a multiple of 16 B -> fallback
    f32 = torch.full((8, 1024), float(rank + 1), dtype=torch.float32)            # dtype the kernel does not take -> fallback
    res = [impl.all_reduce(t) for t in (small, big, odd, f32)]
    gathered = impl.all_gather(torch.full((2, 4), float(rank)))
    if rank == 0:
        out.put((comm.calls, Fallback.calls, [float(r.float().mean()) for r in res], gathered.shape))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_allreduce_plugin_routes_by_message_size():
    """B200DistributedImpl (the reference's DistributedImpl interface, distributed/impl.py:16-21): decode-sized
    16-bit messages go to the push kernel's communicator, everything else to the previous plug-in."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29731
    procs = [ctx.Process(target=_plugin_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ours, fallback, means, gshape = out.get(timeout=100)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert ours == 1 and fallback == 3
    assert means == [3.0, 3.0, 3.0, 3.0]  # 1 + 2 on every element, whichever path carried it
    assert tuple(gshape) == (4, 4)


def test_enable_b200_allreduce_is_a_noop_for_tp1():
    import importlib

    d_mod = importlib.import_module("mini-sglang_b200.distributed")
    assert d_mod.enable_b200_allreduce(0, 1, None, "cpu") is None
