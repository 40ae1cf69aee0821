#!/bin/bash
# 1-GPU session: quick parity of the changed kernels, reference-GPU goldens, probes, ncu captures, sweeps, bench lines.
O=gpurun_out/r2c6; mkdir -p $O
run_pytest() { local name=$1; shift
  timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_insitu.py "$@" > $O/pytest_$name.log 2>&1
  echo "pytest_$name rc=$? : $(tail -1 $O/pytest_$name.log)"; }
run_pytest attention tests/test_gpu_attention.py
timeout 600 python tests/golden/make_reference_gpu_golden.py > $O/golden.log 2>&1; echo "golden rc=$?"; tail -12 $O/golden.log | cut -c1-250
cp gpurun_out/reference_gpu_golden.npz gpurun_out/reference_gpu_parity.txt $O/ 2>/dev/null
timeout 300 python tools/e2e_probe.py > $O/e2e_probe.json 2> $O/e2e_probe.err; cat $O/e2e_probe.json
timeout 300 python tools/elementwise_bench.py --out $O/elementwise.json > $O/elementwise.log 2>&1; cat $O/elementwise.log | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_elementwise.csv python tools/elementwise_bench.py --reps 2 > /dev/null 2>&1; echo "ncu elementwise rc=$?"
# ncu --set full: prefill (long prompts, default softmax variant + full-row), prefill cfg1 batch, fused decode
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_tc -s 2 -c 1 -o $O/prefill_cfg4 python tools/microbench.py prefill --config cfg4 --layers 1 --reps 3 --batches 1 > $O/ncu_prefill_cfg4.log 2>&1; echo "ncu prefill cfg4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_tc -s 2 -c 1 -o $O/prefill_cfg4_fullrow python tools/microbench.py prefill --config cfg4 --layers 1 --reps 3 --batches 1 --opt prefill_full_row=1 > $O/ncu_prefill_cfg4_fr.log 2>&1; echo "ncu prefill cfg4 fullrow rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_tc -s 2 -c 1 -o $O/prefill_cfg1 python tools/microbench.py prefill --layers 1 --reps 3 --batches 1 > $O/ncu_prefill_cfg1.log 2>&1; echo "ncu prefill cfg1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_decode_tc -s 2 -c 1 -o $O/decode_fused python tools/microbench.py decode --iter 500 --layers 1 --reps 3 > $O/ncu_decode.log 2>&1; echo "ncu decode rc=$?"
# launch list of the bench command (shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 8 --warmup 3 --skip-prefill --skip-cpu --skip-ref-gpu > /dev/null 2>&1; echo "ncu launches rc=$?"
# split-KV policy at the TP shard shapes (captured graphs of 28 launches)
for shape in "8 4" "4 2" "2 1"; do set -- $shape
  timeout 600 python tools/decode_sweep.py --graph --layers 28 --hq $1 --hkv $2 --iters 100,500,900 --targets 1,2,4 --nosplit 0,75 --fused 0,1 --out $O/decode_sweep_hq$1_hkv$2.json > $O/decode_sweep_hq$1_hkv$2.log 2>&1; grep BEST $O/decode_sweep_hq$1_hkv$2.log | cut -c1-250
done
timeout 900 python bench.py --steps 40 --warmup 4 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"; cut -c1-400 $O/bench_cfg1.json
timeout 900 python bench.py --config cfg2 --steps 20 --warmup 3 --prefill-batches 3 --prefill-layers 8 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"; cut -c1-400 $O/bench_cfg2.json; tail -2 $O/bench_cfg2.err
timeout 900 python bench.py --config cfg4 --steps 20 --warmup 3 --prefill-batches 4 --prefill-layers 8 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?"; cut -c1-400 $O/bench_cfg4.json; tail -2 $O/bench_cfg4.err
# one rank's shard of tp2 / tp4 / tp8 on this GPU (no all-reduce partner): where does the per-layer time go?
for n in 2 4 8; do
  timeout 600 python bench.py --tp-shard $n --steps 40 --warmup 4 --skip-prefill --skip-cpu --skip-ref-gpu > $O/bench_shard$n.json 2> $O/bench_shard$n.err; python - <<PY
import json
d = json.loads([l for l in open("$O/bench_shard$n.json") if l.startswith("{")][-1])
r = d["roofline"]
print("shard $n value", d["value"], "ms/step", d["ms_per_step"], "us/layer", round(d["ms_per_step"] * 1e3 / 28, 1), "| attention-only us/launch", r["us_per_launch"], "frac", r["frac"], "e2e", d["e2e"]["value"])
PY
done
timeout 600 python bench.py --tp-shard 8 --steps 40 --warmup 4 --skip-prefill --skip-cpu --skip-ref-gpu --opt decode_early_kv=0 > $O/bench_shard8_noearly.json 2>/dev/null; cut -c1-200 $O/bench_shard8_noearly.json
timeout 600 python bench.py --tp-shard 8 --steps 40 --warmup 4 --skip-prefill --skip-cpu --skip-ref-gpu --unfused-pre-attention > $O/bench_shard8_unfused.json 2>/dev/null; cut -c1-200 $O/bench_shard8_unfused.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 400 --csv --log-file $O/launches_shard8.csv python bench.py --tp-shard 8 --steps 8 --warmup 3 --skip-prefill --skip-cpu --skip-ref-gpu > /dev/null 2>&1; echo "ncu launches shard8 rc=$?"
ls -la $O | head -60
