#!/bin/bash
N=${1:-4}; O=gpurun_out/r2tp${N}b; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29512 bench.py --gpus $N --steps 40 --warmup 4 --skip-cpu --skip-prefill > $O/bench_b200ar.json 2> $O/bench_b200ar.err; echo "bench rc=$?"; grep "^{" $O/bench_b200ar.json | cut -c1-600; tail -3 $O/bench_b200ar.err | cut -c1-300
timeout 200 $TR --master-port 29511 tools/ar_check.py --out $O/ar_check.json > $O/ar_check.log 2>&1; echo "ar_check rc=$?"; grep "^{" $O/ar_check.log | tail -1 | cut -c1-1400
