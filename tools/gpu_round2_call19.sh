#!/bin/bash
# 1 GPU: the default bench line after the last bench.py changes; ncu --set full of the decode kernel on the
# cfg2 batch (GQA 5, 256-token prefix shared by all requests): DRAM bytes vs algorithmic bytes, L2 hit rate.
O=gpurun_out/r2c19; mkdir -p $O
timeout 600 python bench.py > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"; tail -2 $O/bench_cfg1.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2c19/bench_cfg1.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "gpu_launches", "host_step_own_ms", "gpu_step_ms")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"])
print({k: v for k, v in d["ref_gpu"].items() if "oracle" in k or "p16" in k or k.endswith("_ok") or "parity_b200" in k or "parity_trt" in k})
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_decode_tc -s 2 -c 1 -o $O/decode_cfg2 python tools/microbench.py decode --config cfg2 --iter 127 --layers 1 --reps 3 > $O/ncu_decode_cfg2.log 2>&1; echo "ncu decode cfg2 rc=$?"; tail -3 $O/ncu_decode_cfg2.log
ls -la $O
for i in 1 2 3 4 5 6; do
  timeout 300 python bench.py --tp-shard 8 --steps 40 --warmup 4 --skip-cpu --skip-prefill --skip-ref-gpu > $O/shard8_$i.json 2> $O/shard8_$i.err
  echo "shard8 run $i: $(grep '^{' $O/shard8_$i.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], 'gpu_step', d['gpu_step_ms'], 'host_own', d['host_step_own_ms'])")"
done
