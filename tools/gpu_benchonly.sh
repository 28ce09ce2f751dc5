#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "ms_per_step": [0-9.]*' gpurun_out/bench.log; grep -o '"e2e": {[^}]*}' gpurun_out/bench.log; grep -o '"host": {[^}]*}' gpurun_out/bench.log
