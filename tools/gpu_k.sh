#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -2
timeout 300 python tools/perf_gemm_k.py 2>&1 | tee gpurun_out/perf_gemm_k.log | grep -E "K3072|K1024" | head -30
timeout 200 python tools/perf_gemm.py > gpurun_out/perf_gemm.log 2>&1; tail -34 gpurun_out/perf_gemm.log
