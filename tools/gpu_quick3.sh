#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_e2e_gpu.py tests/test_dropout_gpu.py tests/test_bert_kernels_gpu.py tests/test_seg_kernels_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_quick.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02_pytest_quick.log | cut -c1-300
timeout 300 python tools/perf_gemm.py 2>&1 | grep "^TN\|auto" > gpurun_out/r02_perf_gemm_tn.log; cat gpurun_out/r02_perf_gemm_tn.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches']/20)"
