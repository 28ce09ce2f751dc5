#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "stem or pos" 2>&1 | tail -2
timeout 300 python tools/perf_gemm_small.py 2>&1 | tee gpurun_out/perf_gemm_small.log | tail -22
