#!/bin/bash
# Final 1-GPU validation: what the driver runs (pytest -m gpu, smoke, bench both arms) + the extra configs.
O=gpurun_out/r2final; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? : $(tail -1 $O/pytest_gpu.log)"; grep -h "^FAILED\|^ERROR" $O/pytest_gpu.log | head -20
cp gpurun_out/insitu_summary.json $O/ 2>/dev/null
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $O/smoke.log
timeout 900 python bench.py > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench cfg1 rc=$?"; cut -c1-300 $O/bench_cfg1.json
timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"; cut -c1-400 $O/bench_reference.json
timeout 900 python bench.py --config cfg2 --steps 20 --warmup 3 --prefill-batches 3 --prefill-layers 8 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 rc=$?"; cut -c1-200 $O/bench_cfg2.json
timeout 900 python bench.py --config cfg4 --steps 20 --warmup 3 --prefill-batches 4 --prefill-layers 8 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "bench cfg4 rc=$?"; cut -c1-200 $O/bench_cfg4.json
timeout 300 python tools/elementwise_bench.py --out $O/elementwise.json > $O/elementwise.log 2>&1; tail -8 $O/elementwise.log
for n in 2 4 8; do timeout 600 python bench.py --tp-shard $n --steps 40 --warmup 4 --skip-prefill --skip-cpu --skip-ref-gpu > $O/bench_shard$n.json 2> $O/bench_shard$n.err; cut -c1-160 $O/bench_shard$n.json; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 8 --warmup 3 --skip-prefill --skip-cpu --skip-ref-gpu > /dev/null 2>&1; echo "ncu launches rc=$?"
