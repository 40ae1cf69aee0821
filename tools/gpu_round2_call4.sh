#!/bin/bash
# One GPU-box session (1 GPU): parity tests in separate processes (a kernel trap poisons its CUDA context),
# the in-situ reference run, A/B timings, bench lines.
O=gpurun_out/r2c4; mkdir -p $O
run_pytest() { # name, extra args...
  local name=$1; shift
  timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_insitu.py "$@" > $O/pytest_$name.log 2>&1
  echo "pytest_$name rc=$? : $(tail -1 $O/pytest_$name.log)"
}
run_pytest fused -k "fused"
run_pytest prefill -k "prefill and not fused"
run_pytest rest -k "not fused and not prefill"
grep -h "^FAILED\|^ERROR" $O/pytest_*.log | head -30
timeout 1500 python tools/insitu.py run --out $O/insitu_summary.json > $O/insitu_run.log 2>&1; echo "insitu rc=$?"
for d in /tmp/insitu_*; do mkdir -p $O/insitu_logs; cp $d/*.log $O/insitu_logs/ 2>/dev/null; done
python - <<'PY'
import json
try:
    s = json.load(open("gpurun_out/r2c4/insitu_summary.json"))
    for k, v in s.items():
        if isinstance(v, dict) and "logits_rel_worst" in v:
            print(k, "logits_rel_worst=%.3e" % v["logits_rel_worst"], "kv_first_bit_exact", v["k_first_bit_exact"], v["v_first_bit_exact"], "argmax", v["argmax_agree"], "graph", v["graph_replays"], "ext", v["extend_reqs_with_cache_hit"], "chunk", v["chunked_reqs"], "launches", v["b200_launches"])
        elif not isinstance(v, dict):
            print(k, v)
    print("errors:", list(s.get("errors", {})))
except Exception as e:
    print("no insitu summary:", e)
PY
for fr in 1 0; do
  timeout 300 python tools/microbench.py prefill --layers 4 --reps 3 --opt prefill_full_row=$fr > $O/prefill_cfg1_fullrow$fr.log 2>&1; tail -4 $O/prefill_cfg1_fullrow$fr.log
  timeout 300 python tools/microbench.py prefill --config cfg4 --layers 4 --reps 3 --batches 1 --opt prefill_full_row=$fr > $O/prefill_cfg4_fullrow$fr.log 2>&1; tail -2 $O/prefill_cfg4_fullrow$fr.log
done
timeout 900 python bench.py --steps 40 --warmup 4 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?"; cut -c1-1200 $O/bench_cfg1.json; tail -3 $O/bench_cfg1.err
timeout 600 python bench.py --steps 40 --warmup 4 --unfused-pre-attention --skip-prefill --skip-cpu --skip-ref-gpu > $O/bench_cfg1_unfused.json 2> $O/bench_cfg1_unfused.err; cut -c1-300 $O/bench_cfg1_unfused.json
timeout 600 python bench.py --steps 40 --warmup 4 --opt decode_early_kv=0 --skip-prefill --skip-cpu --skip-ref-gpu > $O/bench_cfg1_noearly.json 2> $O/bench_cfg1_noearly.err; cut -c1-300 $O/bench_cfg1_noearly.json
