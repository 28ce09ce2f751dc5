#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_seg_kernels_gpu.py tests/test_gemm_gpu.py -x -q > gpurun_out/seg_test.log 2>&1; echo "seg kernels rc=$?"; tail -12 gpurun_out/seg_test.log
timeout 900 python -m pytest tests/test_e2e_gpu.py -x -q -s -k "seg or missing" > gpurun_out/seg_e2e.log 2>&1; echo "seg e2e rc=$?"; tail -14 gpurun_out/seg_e2e.log | cut -c1-600
