#!/usr/bin/env python
"""Why is the host-buffer (e2e) leg slower than the device-resident one?  Measures, on the bench workload:
  A. one step's qkv upload alone (pinned host -> device, 30 MB), B. one decode graph replay alone,
  C. upload on a copy stream CONCURRENT with the replay, D. upload then replay on one stream (serial).
If C ~ max(A, B) the engines overlap and the e2e leg's pipelining is at fault; if C ~ A + B they do not."""
import importlib
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    pkg = importlib.import_module("mini-sglang_b200")
    pkg.build_native()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    sched = bench.Schedule()
    r = bench.AttentionPathRunner(pkg, sched, bench.HQ, bench.HKV, 64, dev, fuse_pre_attention=True)
    tr = sched.live(500)
    bs = r.pad_bs(len(tr))
    r.capture(bs)
    st = r.schedule_step(tr)
    for _ in range(3):
        r.decode_step(st)
    torch.cuda.synchronize()
    qkv_h = torch.empty(tuple(r.qkv.shape), dtype=torch.bfloat16).pin_memory()
    stg = torch.empty_like(r.qkv)
    copy_stream = torch.cuda.Stream(dev)
    g = r.graphs[bs]

    def timed(fn, reps=10):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(r.stream):
                e0.record()
            fn()
            with torch.cuda.stream(r.stream):
                e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return round(sorted(ts)[len(ts) // 2], 4)

    def upload_only():
        with torch.cuda.stream(r.stream):
            stg[:bs].copy_(qkv_h[:bs], non_blocking=True)

    def replay_only():
        with torch.cuda.stream(r.stream):
            g.replay()

    def concurrent():
        ev = torch.cuda.Event()
        copy_stream.wait_stream(r.stream)
        with torch.cuda.stream(copy_stream):
            stg[:bs].copy_(qkv_h[:bs], non_blocking=True)
            ev.record(copy_stream)
        with torch.cuda.stream(r.stream):
            g.replay()
            r.stream.wait_event(ev)

    def serial():
        with torch.cuda.stream(r.stream):
            stg[:bs].copy_(qkv_h[:bs], non_blocking=True)
            g.replay()

    res = {"bs": bs, "upload_MB": round(bs * bench.L * r.width * 2 / 1e6, 2), "A_upload_ms": timed(upload_only), "B_replay_ms": timed(replay_only),
           "C_concurrent_ms": timed(concurrent), "D_serial_ms": timed(serial)}
    res["upload_GBs"] = round(res["upload_MB"] / res["A_upload_ms"], 1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
