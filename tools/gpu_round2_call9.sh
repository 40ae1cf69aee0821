#!/bin/bash
# 1-GPU session: prefill v3 (decoupled) parity + A/B, combine / plan-policy effect at the tp8 shard shape.
O=gpurun_out/r2c9; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q --maxfail=10 -k "prefill or golden or decode_qwen3" > $O/pytest.log 2>&1; echo "pytest rc=$? : $(tail -1 $O/pytest.log)"; grep -h "^FAILED\|^ERROR" $O/pytest.log | head -20
for fr in 2 0; do
  timeout 300 python tools/microbench.py prefill --layers 4 --reps 3 --opt prefill_full_row=$fr > $O/prefill_cfg1_v$fr.log 2>&1; tail -4 $O/prefill_cfg1_v$fr.log
  timeout 300 python tools/microbench.py prefill --config cfg4 --layers 4 --reps 3 --batches 1 --opt prefill_full_row=$fr > $O/prefill_cfg4_v$fr.log 2>&1; tail -2 $O/prefill_cfg4_v$fr.log
  timeout 300 python tools/microbench.py prefill --config cfg2 --layers 4 --reps 3 --batches 2 --opt prefill_full_row=$fr > $O/prefill_cfg2_v$fr.log 2>&1; tail -4 $O/prefill_cfg2_v$fr.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_v3 -s 2 -c 1 -o $O/prefill_v3_cfg4 python tools/microbench.py prefill --config cfg4 --layers 1 --reps 3 --batches 1 --opt prefill_full_row=2 > $O/ncu_prefill_v3_cfg4.log 2>&1; echo "ncu v3 cfg4 rc=$?"
for n in 2 4 8; do
  timeout 600 python bench.py --tp-shard $n --steps 40 --warmup 4 --skip-prefill --skip-cpu --skip-ref-gpu > $O/bench_shard$n.json 2> $O/bench_shard$n.err; python - <<PY
import json
d = json.loads([l for l in open("$O/bench_shard$n.json") if l.startswith("{")][-1])
r = d["roofline"]
print("shard $n value", d["value"], "ms/step", d["ms_per_step"], "us/layer", round(d["ms_per_step"] * 1e3 / 28, 1), "| attention-only us/launch", r["us_per_launch"], "frac", r["frac"], "e2e", d["e2e"]["value"], d["e2e"]["frac_of_value"])
PY
done
