#!/bin/bash
mkdir -p gpurun_out
for d in noddp replicas; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --diag $d > gpurun_out/bench_n2_$d.log 2>&1; echo "$d rc=$?"
grep -o '"value": [0-9.]*, "ms_per_step": [0-9.]*' gpurun_out/bench_n2_$d.log; grep -o '"host": {[^}]*}' gpurun_out/bench_n2_$d.log
done
