#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention or colsum or layernorm" > gpurun_out/attn_test.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/attn_test.log
timeout 120 python tools/perf_attn.py 2>&1 | tail -5
