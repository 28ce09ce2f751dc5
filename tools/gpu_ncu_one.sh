#!/bin/bash
# usage: tools/gpu_ncu_one.sh <kernel-regex> <out-name> [count]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$1 -c ${3:-1} -f -o gpurun_out/$2 python tools/profile_step.py > gpurun_out/ncu_$2.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; grep -o '"value": [0-9.]*, "ms_per_step": [0-9.]*' gpurun_out/bench_quick.log; grep -o '"e2e": {[^}]*}' gpurun_out/bench_quick.log
