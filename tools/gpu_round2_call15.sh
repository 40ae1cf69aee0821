#!/bin/bash
# 1 GPU: in-kernel split-KV merge (batched loads) vs the combine launch, in captured graphs, at the
# tp1/2/4/8 shard shapes; decode parity with the merge in kernel; cfg2 bench with the per-request prefix share.
O=gpurun_out/r2c15; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q -k "decode" > $O/pytest_decode.log 2>&1; echo "pytest decode rc=$?"; tail -3 $O/pytest_decode.log
timeout 300 python -m pytest tests/test_gpu_elementwise.py -x -q > $O/pytest_elem.log 2>&1; echo "pytest elementwise rc=$?"; tail -2 $O/pytest_elem.log
sweep() {
  timeout 400 python tools/decode_sweep.py --graph --layers 28 --hq $1 --hkv $2 --iters 100,500,900,1000 --targets 2 --nosplit 75 --fused 0,1,2 --out $O/decode_sweep_hq$1_hkv$2.json > $O/decode_sweep_hq$1_hkv$2.log 2>&1
  grep -v BEST $O/decode_sweep_hq$1_hkv$2.log | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('hq', $1, 'bs', r['bs'], 'items', r['items'], 'fused', r['fused'], 'us', r['us'], 'frac', r['frac'])"
}
sweep 16 8
sweep 8 4
sweep 4 2
sweep 2 1
timeout 600 python bench.py --config cfg2 --steps 20 --warmup 3 --prefill-batches 3 --prefill-layers 8 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; grep "^{" $O/bench_cfg2.json | cut -c1-400
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2c15/bench_cfg2.json") if l.startswith("{")][-1])
print({k: d.get(k) for k in ("value", "e2e", "roofline")})
print("ref_gpu", json.dumps(d.get("ref_gpu"))[:1500])
print("prefill", json.dumps(d.get("prefill"))[:600])
PY
for sh in 0 4 8; do for fc in 2 1; do
  timeout 300 python bench.py --tp-shard $sh --steps 40 --warmup 4 --skip-cpu --skip-prefill --skip-ref-gpu --opt decode_fused_combine=$fc > $O/bench_shard${sh}_merge$fc.json 2> $O/bench_shard${sh}_merge$fc.err
  echo "shard $sh decode_fused_combine=$fc rc=$?: $(grep '^{' $O/bench_shard${sh}_merge$fc.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['e2e']['value'])")"
done; done
