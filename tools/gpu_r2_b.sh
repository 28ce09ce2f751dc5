#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/r02_pytest_gpu_b.log 2>&1; echo "pytest rc=$?"
grep "FULLSIZE\|passed\|failed\|Error\|assert " gpurun_out/r02_pytest_gpu_b.log | cut -c1-600 | head -50
timeout 600 python tools/timeline.py --out gpurun_out/r02_timeline_cfg2 > gpurun_out/r02_timeline.log 2>&1; echo "timeline rc=$?"
head -12 gpurun_out/r02_timeline_cfg2_summary.txt | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'], d['ms_per_step'])"
