#!/usr/bin/env python
"""Multi-GPU check + timing of the one-shot NVLink all-reduce (csrc/allreduce.cu), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/ar_check.py [--out gpurun_out/ar_check_tp2.json]

Checks (every rank; rank 0 prints one JSON line):
  * plain all-reduce == round(fp32 sum of the ranks' inputs in rank order), bit exact, for decode and
    small-prefill shapes of the BASELINE models (hidden 1024 / 5120 / 8192), bf16 and fp16, strided
    input rows; identical bits on every rank; at world 2 also bit-identical to NCCL's result;
  * fused all-reduce + residual add + RMSNorm == our all-reduce followed by b200_fused_add_rmsnorm,
    bit exact (residual and output);
  * 28 fused calls captured in ONE CUDA graph, replayed: same bits as the eager sequence;
  * 1500 back-to-back calls with changing row counts (buffer parity / epoch / flag reuse);
  * the reference's plug-in point, when baseline/_ref is importable: `DistributedCommunicator().all_reduce`
    routes through `B200DistributedImpl` (and large messages through the previous plug-in).
Timing: CUDA events, median of 200, per message size, ours vs torch.distributed (NCCL) eager and
NCCL-in-graph when capture works; device-side max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-timing", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import importlib

    pkg = importlib.import_module("mini-sglang_b200")
    pkg.build_native()
    from importlib import import_module

    d_mod = import_module("mini-sglang_b200.distributed")
    comm = d_mod.B200AllReduce(rank, world, dist.group.WORLD, dev, max_bytes=4 << 20)
    res = {"world": world, "checks": {}, "timing_us": {}}
    g = torch.Generator(device=dev).manual_seed(1000 + rank)

    def gathered_sum(x):
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous())
        acc = torch.zeros_like(x, dtype=torch.float32)
        for p in parts:  # rank order, fp32, one rounding at the end
            acc += p.float()
        return acc.to(x.dtype)

    def same_on_all_ranks(t):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        return all(torch.equal(parts[0].view(torch.int16), p.view(torch.int16)) for p in parts)

    # ---------------------------------------------------------------- 1. plain
    ok_plain = ok_ranks = ok_nccl = True
    for dtype in (torch.bfloat16, torch.float16):
        for rows, dim in ((1, 1024), (7, 1024), (130, 1024), (256, 1024), (64, 5120), (33, 8192), (256, 8192)):
            wide = torch.randn((rows, dim + 64), device=dev, generator=g, dtype=torch.float32).to(dtype)
            x = wide[:, :dim]  # strided rows
            want = gathered_sum(x)
            out = torch.empty((rows, dim), device=dev, dtype=dtype)
            comm.all_reduce(x, out=out)
            ok_plain &= torch.equal(out.view(torch.int16), want.view(torch.int16))
            ok_ranks &= same_on_all_ranks(out)
            if world == 2:
                y = x.contiguous().clone()
                dist.all_reduce(y)
                ok_nccl &= torch.equal(out.view(torch.int16), y.view(torch.int16))
            z = x.contiguous().clone()  # in place
            comm.all_reduce(z)
            ok_plain &= torch.equal(z.view(torch.int16), want.view(torch.int16))
    res["checks"]["plain_bit_exact_vs_fp32_rank_order_sum"] = bool(ok_plain)
    res["checks"]["identical_on_all_ranks"] = bool(ok_ranks)
    if world == 2:
        res["checks"]["bit_identical_to_nccl_world2"] = bool(ok_nccl)

    # ---------------------------------------------------------------- 2. fused
    ok_fused = True
    for rows, dim in ((1, 1024), (96, 1024), (256, 1024), (48, 5120), (64, 8192)):
        x = torch.randn((rows, dim), device=dev, generator=g, dtype=torch.float32).to(torch.bfloat16)
        resid = torch.randn((rows, dim), device=dev, dtype=torch.float32, generator=torch.Generator(device=dev).manual_seed(7)).to(torch.bfloat16)
        w = (torch.rand(dim, device=dev, generator=torch.Generator(device=dev).manual_seed(8)) + 0.5).to(torch.bfloat16)
        y = x.clone()
        comm.all_reduce(y)
        r_ref = resid.clone()
        pkg.ops.fused_add_rmsnorm(y, r_ref, w, 1e-6)
        r_got, o_got = resid.clone(), torch.empty_like(x)
        comm.all_reduce(x, out=o_got, residual=r_got, weight=w, eps=1e-6)
        ok_fused &= torch.equal(o_got.view(torch.int16), y.view(torch.int16))
        ok_fused &= torch.equal(r_got.view(torch.int16), r_ref.view(torch.int16))
    res["checks"]["fused_equals_allreduce_then_fused_add_rmsnorm"] = bool(ok_fused)

    # ---------------------------------------------------------------- 3. CUDA graph
    rows, dim, layers = 130, 1024, 28
    xs = torch.randn((layers, rows, dim), device=dev, generator=g, dtype=torch.float32).to(torch.bfloat16)
    w = torch.ones(dim, device=dev, dtype=torch.bfloat16)
    resid0 = torch.zeros((rows, dim), device=dev, dtype=torch.bfloat16)
    outs = torch.empty_like(xs)

    def sequence(resid):
        for l in range(layers):
            comm.all_reduce(xs[l], out=outs[l], residual=resid, weight=w, eps=1e-6)

    r_eager = resid0.clone()
    sequence(r_eager)
    o_eager = outs.clone()
    r_graph = resid0.clone()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        r_static = resid0.clone()
        sequence(r_static)  # warm-up on the capture stream
        r_static.copy_(resid0)
        st.synchronize()
        with torch.cuda.graph(graph, stream=st):
            sequence(r_static)
    ok_graph = True
    for _ in range(3):
        r_static.copy_(resid0)
        graph.replay()
        torch.cuda.synchronize()
        ok_graph &= torch.equal(r_static.view(torch.int16), r_eager.view(torch.int16))
        ok_graph &= torch.equal(outs.view(torch.int16), o_eager.view(torch.int16))
    res["checks"]["graph_replay_equals_eager"] = bool(ok_graph)

    # ---------------------------------------------------------------- 4. stress
    import random

    rnd = random.Random(5)
    acc_ref = torch.zeros((256, 1024), device=dev, dtype=torch.float32)
    buf = torch.empty((256, 1024), device=dev, dtype=torch.bfloat16)
    base = torch.randn((256, 1024), device=dev, generator=g, dtype=torch.float32).to(torch.bfloat16)
    total = gathered_sum(base).float()
    ok_stress = True
    for it in range(1500):
        n = rnd.choice([1, 2, 3, 8, 31, 64, 129, 200, 256])
        buf[:n].copy_(base[:n])
        comm.all_reduce(buf[:n])
        if it % 100 == 0:
            ok_stress &= torch.equal(buf[:n].float(), total[:n].to(torch.bfloat16).float())
    torch.cuda.synchronize()
    res["checks"]["stress_1500_calls"] = bool(ok_stress)

    # ---------------------------------------------------------------- 5. reference plug-in point
    ref = next((d for d in (ROOT / "oracle" / "_ref" / "minisgl_site", ROOT / "baseline" / "_ref")
                if (d / "minisgl" / "core.py").exists()), None)
    if ref is not None:
        sys.path.insert(0, str(ref))
        from minisgl.distributed import DistributedCommunicator

        before = list(DistributedCommunicator.plugins)
        DistributedCommunicator.plugins.append(d_mod.B200DistributedImpl(comm, DistributedCommunicator.plugins[-1]))
        c = DistributedCommunicator()
        x = torch.randn((64, 1024), device=dev, generator=g, dtype=torch.float32).to(torch.bfloat16)
        want = gathered_sum(x)
        launches0 = pkg._cabi.launch_count()
        got = c.all_reduce(x.clone())
        small_ok = torch.equal(got.view(torch.int16), want.view(torch.int16)) and pkg._cabi.launch_count() == launches0 + 1
        big = torch.randn((4096, 1024), device=dev, generator=g, dtype=torch.float32).to(torch.bfloat16)  # 8 MB > max_bytes
        want_big = gathered_sum(big).float()
        launches0 = pkg._cabi.launch_count()
        got_big = c.all_reduce(big.clone())
        big_ok = pkg._cabi.launch_count() == launches0 and (got_big.float() - want_big).abs().max().item() <= 0.07 * want_big.abs().max().item()
        DistributedCommunicator.plugins[:] = before
        res["checks"]["reference_plugin_point"] = bool(small_ok and big_ok)

    # ---------------------------------------------------------------- timing
    if not args.skip_timing:
        def timed(fn, reps=200):
            for _ in range(20):
                fn()
            ts = []
            for _ in range(reps):
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return round(float(t.item()), 2)

        def timed_graph(fn, n=28, reps=30):
            s2 = torch.cuda.Stream()
            s2.wait_stream(torch.cuda.current_stream())
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.stream(s2):
                fn()
                s2.synchronize()
                with torch.cuda.graph(gr, stream=s2):
                    for _ in range(n):
                        fn()
            ts = []
            for _ in range(reps):
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(s2):
                    e0.record()
                    gr.replay()
                    e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3 / n)
            t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return round(float(t.item()), 2)

        w1 = torch.ones(1024, device=dev, dtype=torch.bfloat16)
        for rows in (1, 8, 64, 128, 256):
            x = torch.randn((rows, 1024), device=dev, dtype=torch.bfloat16)
            rs = torch.zeros_like(x)
            o = torch.empty_like(x)
            key = f"{rows}x1024_bf16"
            res["timing_us"][key] = {
                "b200_eager": timed(lambda: comm.all_reduce(x, out=o)),
                "b200_fused_norm_eager": timed(lambda: comm.all_reduce(x, out=o, residual=rs, weight=w1, eps=1e-6)),
                "nccl_eager": timed(lambda: dist.all_reduce(x)),
                "b200_in_graph": timed_graph(lambda: comm.all_reduce(x, out=o)),
                "b200_fused_norm_in_graph": timed_graph(lambda: comm.all_reduce(x, out=o, residual=rs, weight=w1, eps=1e-6)),
            }
    ok = all(res["checks"].values())
    res["ok"] = bool(ok)
    flags = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    res["ok_all_ranks"] = bool(flags.item())
    comm.destroy()
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if args.out:
            Path(args.out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.out).write_text(json.dumps(res, indent=1))
    dist.destroy_process_group()
    if not res["ok_all_ranks"]:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
