#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
grep -o '"value": [0-9.]*, "ms_per_step": [0-9.]*' gpurun_out/bench.log; grep -o '"e2e": {[^}]*}' gpurun_out/bench.log; grep -o '"roofline": {[^[]*' gpurun_out/bench.log | cut -c1-600; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/bench.log
