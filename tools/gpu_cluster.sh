#!/bin/bash
# CTA-pair (tcgen05.mma.cta_group::2) GEMM: correctness with pairs forced wherever legal, then timings pairs on / off (same box)
mkdir -p gpurun_out
RB_GEMM_CLUSTER=1 timeout 240 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_cluster1.log 2>&1; rc=$?; echo "gemm tests (pairs forced) rc=$rc"
tail -6 gpurun_out/r02_pytest_cluster1.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_e2e_gpu.py tests/test_seg_kernels_gpu.py tests/test_bert_kernels_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_cluster.log 2>&1; echo "tests (auto) rc=$?"
tail -4 gpurun_out/r02_pytest_cluster.log | cut -c1-300
RB_GEMM_CLUSTER=0 timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_cl0.log 2>&1
timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_cl2.log 2>&1
RB_GEMM_CLUSTER=1 timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_cl1.log 2>&1
paste -d"|" <(cut -c1-62 gpurun_out/r02_perf_gemm_cl0.log) <(cut -c45-62 gpurun_out/r02_perf_gemm_cl2.log) <(cut -c45-62 gpurun_out/r02_perf_gemm_cl1.log)
for c in 0 auto 0 auto; do
  if [ $c = auto ]; then unset RB_GEMM_CLUSTER; else export RB_GEMM_CLUSTER=$c; fi
  REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 timeout 300 python bench.py --steps 20 --warmup 5 --windows 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench cluster=$c', d['value'], d['e2e']['value'], d['ms_per_step'])"
done
