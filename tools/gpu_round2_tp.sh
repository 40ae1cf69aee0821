#!/bin/bash
# Multi-GPU session: all-reduce checks + timings, bench at N GPUs with the captured push all-reduce and with NCCL.
N=${1:-2}; O=gpurun_out/r2tp$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/ar_check.py --out $O/ar_check.json > $O/ar_check.log 2>&1; echo "ar_check rc=$?"; grep "^{" $O/ar_check.log | cut -c1-1500; tail -5 $O/ar_check.log | cut -c1-300
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 40 --warmup 4 --skip-cpu > $O/bench_b200ar.json 2> $O/bench_b200ar.err; echo "bench rc=$?"; cut -c1-700 $O/bench_b200ar.json; tail -3 $O/bench_b200ar.err | cut -c1-300
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 40 --warmup 4 --skip-cpu --skip-prefill --allreduce nccl > $O/bench_nccl.json 2> $O/bench_nccl.err; echo "bench nccl rc=$?"; cut -c1-400 $O/bench_nccl.json
timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 40 --warmup 4 --skip-cpu --skip-prefill --no-allreduce > $O/bench_noar.json 2> $O/bench_noar.err; echo "bench noar rc=$?"; cut -c1-400 $O/bench_noar.json
