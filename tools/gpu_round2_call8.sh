#!/bin/bash
# 1-GPU session: prefill v3 parity + A/B timings + ncu, bench lines.
O=gpurun_out/r2c8; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_attention.py -m gpu -q --maxfail=10 -k "prefill or golden" > $O/pytest_prefill.log 2>&1; echo "pytest_prefill rc=$? : $(tail -1 $O/pytest_prefill.log)"; grep -h "^FAILED\|^ERROR" $O/pytest_prefill.log | head -20
for fr in 2 0; do
  timeout 300 python tools/microbench.py prefill --layers 4 --reps 3 --opt prefill_full_row=$fr > $O/prefill_cfg1_v$fr.log 2>&1; tail -4 $O/prefill_cfg1_v$fr.log
  timeout 300 python tools/microbench.py prefill --config cfg4 --layers 4 --reps 3 --batches 1 --opt prefill_full_row=$fr > $O/prefill_cfg4_v$fr.log 2>&1; tail -2 $O/prefill_cfg4_v$fr.log
  timeout 300 python tools/microbench.py prefill --config cfg2 --layers 4 --reps 3 --batches 2 --opt prefill_full_row=$fr > $O/prefill_cfg2_v$fr.log 2>&1; tail -4 $O/prefill_cfg2_v$fr.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_v3 -s 2 -c 1 -o $O/prefill_v3_cfg4 python tools/microbench.py prefill --config cfg4 --layers 1 --reps 3 --batches 1 --opt prefill_full_row=2 > $O/ncu_prefill_v3_cfg4.log 2>&1; echo "ncu v3 cfg4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_v3 -s 2 -c 1 -o $O/prefill_v3_cfg1 python tools/microbench.py prefill --layers 1 --reps 3 --batches 1 --opt prefill_full_row=2 > $O/ncu_prefill_v3_cfg1.log 2>&1; echo "ncu v3 cfg1 rc=$?"
timeout 900 python bench.py --steps 40 --warmup 4 --opt prefill_full_row=2 > $O/bench_cfg1_v3.json 2> $O/bench_cfg1_v3.err; echo "bench rc=$?"; python - <<PY
import json
d = json.loads([l for l in open("$O/bench_cfg1_v3.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"], "host idle", d.get("host_us_per_step_gpu_idle"))
print("prefill", d["prefill"])
print({k: v for k, v in d["ref_gpu"].items() if "prefill" in k})
PY
