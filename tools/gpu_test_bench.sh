#!/bin/bash
# GPU parity suite + one bench run
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print({k:v for k,v in d['roofline'].items() if k not in ('traffic_note','note','kernel')}); print(d.get('cpu_baseline'))" || tail -5 gpurun_out/bench.log | cut -c1-600
