#!/bin/bash
# N-GPU session: bench at N GPUs with the captured push all-reduce (the driver's command), plus a no-all-reduce run.
N=${1:-4}; O=gpurun_out/r2tp$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 40 --warmup 4 --skip-cpu --skip-prefill > $O/bench_b200ar.json 2> $O/bench_b200ar.err; echo "bench rc=$?"; grep "^{" $O/bench_b200ar.json | cut -c1-500; tail -3 $O/bench_b200ar.err | cut -c1-300
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 40 --warmup 4 --skip-cpu --skip-prefill --no-allreduce > $O/bench_noar.json 2> $O/bench_noar.err; echo "bench noar rc=$?"; grep "^{" $O/bench_noar.json | cut -c1-400
timeout 300 $TR --master-port 29511 tools/ar_check.py --out $O/ar_check.json > $O/ar_check.log 2>&1; echo "ar_check rc=$?"; grep "^{" $O/ar_check.log | tail -1 | cut -c1-1600
