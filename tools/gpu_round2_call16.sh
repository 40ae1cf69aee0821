#!/bin/bash
# 1 GPU: in-kernel merge as the default -- split policy re-swept with it; cfg2 / cfg1 bench lines with the
# three GPU outputs also held against the exact oracle.
O=gpurun_out/r2c16; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q -k "decode" > $O/pytest_decode.log 2>&1; echo "pytest decode rc=$?"; tail -2 $O/pytest_decode.log
sweep() {
  timeout 400 python tools/decode_sweep.py --graph --layers 28 --hq $1 --hkv $2 --iters 100,500,800,900,1000 --targets 1,2,4 --nosplit 75,100,150 --fused 1 --out $O/decode_sweep_hq$1_hkv$2.json > $O/decode_sweep_hq$1_hkv$2.log 2>&1
  grep -v BEST $O/decode_sweep_hq$1_hkv$2.log | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('hq', $1, 'bs', r['bs'], 'tgt', r['target'], 'nosplit', r['nosplit'], 'items', r['items'], 'us', r['us'], 'frac', r['frac'])"
}
sweep 16 8
sweep 8 4
sweep 4 2
sweep 2 1
timeout 600 python bench.py --config cfg2 --steps 20 --warmup 3 --prefill-batches 3 --prefill-layers 8 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 600 python bench.py > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "cfg1 rc=$?"; tail -3 $O/bench_cfg1.err
python - <<'PY'
import json
for c in ("cfg2", "cfg1"):
    d = json.loads([l for l in open(f"gpurun_out/r2c16/bench_{c}.json") if l.startswith("{")][-1])
    print(c, {k: d.get(k) for k in ("value", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"].get("parity_max_rel_err_vs_gpu"))
    print({k: v for k, v in d["ref_gpu"].items() if "oracle" in k or "parity_b200" in k or "parity_trt" in k})
    print("prefill", d["prefill"].get("tflops"))
PY
