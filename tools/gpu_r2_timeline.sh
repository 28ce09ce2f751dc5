#!/bin/bash
# round 2: GPU parity suite + kernel timeline of one replayed cfg2 step (tools/timeline.py)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_gpu.log | cut -c1-300
timeout 600 python tools/timeline.py --out gpurun_out/r02_timeline_cfg2 > gpurun_out/r02_timeline.log 2>&1; echo "timeline rc=$?"
tail -90 gpurun_out/r02_timeline.log | cut -c1-200
