#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/parity_stages.py > gpurun_out/r02_parity_stages.log 2>&1; echo "stages rc=$?"; tail -25 gpurun_out/r02_parity_stages.log
timeout 1500 python -m pytest tests/test_full_size_gpu.py tests/test_criterion_golden.py -m gpu -q -s > gpurun_out/r02_full_size_parity.log 2>&1; echo "pytest rc=$?"
grep "FULLSIZE\|passed\|failed\|Error\|assert" gpurun_out/r02_full_size_parity.log | cut -c1-700 | head -60
