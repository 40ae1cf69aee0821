#!/bin/bash
# 1 GPU: are the intermittent 60-160 ms stalls of the short-step runs (one rank's shard of tp8) host-side?
# Same command eight times, cyclic GC parked (default) vs left on; per-step GPU time and host own time.
O=gpurun_out/r2c18; mkdir -p $O
for i in 1 2 3 4 5 6 7 8; do
  keep=$(( i % 2 ))
  B200_BENCH_KEEP_GC=$keep timeout 300 python bench.py --tp-shard 8 --steps 40 --warmup 4 --skip-cpu --skip-prefill --skip-ref-gpu > $O/shard8_$i.json 2> $O/shard8_$i.err
  echo "run $i keep_gc=$keep: $(grep '^{' $O/shard8_$i.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], 'gpu_step', d['gpu_step_ms'], 'host_own', d['host_step_own_ms'], d['host_us_per_step'])")"
done
