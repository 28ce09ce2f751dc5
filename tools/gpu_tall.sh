#!/bin/bash
# 256-row ("tall") tiles: correctness forced wherever legal (and pairs forced), then timings tall off / auto (same box)
mkdir -p gpurun_out
RB_GEMM_TALL=1 RB_GEMM_CLUSTER=0 timeout 240 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_tall1.log 2>&1; rc=$?; echo "gemm tests (tall forced) rc=$rc"
tail -6 gpurun_out/r02_pytest_tall1.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
RB_GEMM_CLUSTER=1 RB_GEMM_TALL=0 timeout 240 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_cluster1.log 2>&1; echo "gemm tests (pairs forced) rc=$?"
tail -2 gpurun_out/r02_pytest_cluster1.log | cut -c1-300
RB_GEMM_TALL=1 timeout 600 python -m pytest tests/test_e2e_gpu.py tests/test_seg_kernels_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_tall_e2e.log 2>&1; echo "e2e tests (tall forced) rc=$?"
tail -3 gpurun_out/r02_pytest_tall_e2e.log | cut -c1-300
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_e2e_gpu.py tests/test_bert_kernels_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_auto.log 2>&1; echo "tests (auto) rc=$?"
tail -3 gpurun_out/r02_pytest_auto.log | cut -c1-300
RB_GEMM_TALL=0 RB_GEMM_CLUSTER=0 timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_t0.log 2>&1
timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_tauto.log 2>&1
RB_GEMM_TALL=1 RB_GEMM_CLUSTER=0 timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_t1.log 2>&1
paste -d"|" <(cut -c1-62 gpurun_out/r02_perf_gemm_t0.log) <(cut -c45-62 gpurun_out/r02_perf_gemm_tauto.log) <(cut -c45-62 gpurun_out/r02_perf_gemm_t1.log)
for c in 0 auto 0 auto; do
  if [ $c = auto ]; then unset RB_GEMM_TALL; unset RB_GEMM_CLUSTER; else export RB_GEMM_TALL=0; export RB_GEMM_CLUSTER=0; fi
  REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 timeout 300 python bench.py --steps 20 --warmup 5 --windows 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench tall=$c', d['value'], d['e2e']['value'], d['ms_per_step'])"
done
