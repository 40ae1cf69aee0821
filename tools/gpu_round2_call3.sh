#!/bin/bash
# One GPU-box session: parity tests of the new kernels, A/B timings, the bench lines (1 GPU).
O=gpurun_out/r2c3; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 --deselect tests/test_gpu_insitu.py > $O/pytest.log 2>&1; echo "pytest_rc=$?" | tee -a $O/pytest.log
tail -15 $O/pytest.log
for fr in 1 0; do
  timeout 300 python tools/microbench.py prefill --layers 4 --reps 3 --opt prefill_full_row=$fr > $O/prefill_cfg1_fullrow$fr.log 2>&1; tail -4 $O/prefill_cfg1_fullrow$fr.log
  timeout 300 python tools/microbench.py prefill --config cfg4 --layers 4 --reps 3 --batches 1 --opt prefill_full_row=$fr > $O/prefill_cfg4_fullrow$fr.log 2>&1; tail -2 $O/prefill_cfg4_fullrow$fr.log
done
timeout 900 python bench.py --steps 40 --warmup 4 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "bench rc=$?"; cut -c1-900 $O/bench_cfg1.json
timeout 600 python bench.py --steps 40 --warmup 4 --unfused-pre-attention --skip-prefill --skip-cpu --skip-ref-gpu > $O/bench_cfg1_unfused.json 2> $O/bench_cfg1_unfused.err; cut -c1-300 $O/bench_cfg1_unfused.json
timeout 600 python bench.py --steps 40 --warmup 4 --opt decode_early_kv=0 --skip-prefill --skip-cpu --skip-ref-gpu > $O/bench_cfg1_noearly.json 2> $O/bench_cfg1_noearly.err; cut -c1-300 $O/bench_cfg1_noearly.json
