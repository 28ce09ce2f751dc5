#!/bin/bash
# usage: tools/gpu_ddpN.sh N
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $1 --steps 20 --warmup 5 > gpurun_out/bench_n$1.log 2>&1; echo "rc=$?"
grep -o '"value": [0-9.]*, "ms_per_step": [0-9.]*' gpurun_out/bench_n$1.log; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n$1.log; grep -o '"host": {[^}]*}' gpurun_out/bench_n$1.log; tail -3 gpurun_out/bench_n$1.log | grep -i error | head -3
