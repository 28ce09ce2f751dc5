#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_c.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02_pytest_gpu_c.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_bench_c.json | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks','windows_ms_per_step')}); print(d['e2e']); print({k:v for k,v in d['roofline'].items() if k not in ('traffic_note','note','kernel','top_groups')}); print(d.get('cpu_baseline')); print(d.get('stock_gpu_baseline')); print(d.get('parity')); print(d.get('host'))" || tail -5 gpurun_out/r02_bench_c.err
