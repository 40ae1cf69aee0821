#!/bin/bash
# 1 GPU: decode softmax with redux.sync.max.f32 + one 16-byte read of the four warp maxima per head.
O=gpurun_out/r2c20; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q -k "decode" > $O/pytest_decode.log 2>&1; echo "pytest decode rc=$?"; tail -2 $O/pytest_decode.log
for c in cfg2 cfg4 cfg1; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 3 --skip-prefill --skip-cpu > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"
  grep '^{' $O/bench_$c.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_launch'])
print({k: v for k, v in d['ref_gpu'].items() if k.startswith('decode') and ('us_per_layer' in k or k.endswith('_ok') or 'vs_oracle' in k)})"
done
