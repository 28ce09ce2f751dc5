#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/gemm_test.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/gemm_test.log
timeout 200 python tools/perf_gemm.py > gpurun_out/perf_gemm.log 2>&1; echo "perf rc=$?"; cat gpurun_out/perf_gemm.log | tail -45
