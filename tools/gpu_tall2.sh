#!/bin/bash
mkdir -p gpurun_out
RB_GEMM_TALL=0 RB_GEMM_CLUSTER=0 timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_t0.log 2>&1
RB_GEMM_CLUSTER=0 timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_tauto.log 2>&1
timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_tc.log 2>&1
paste -d"|" <(cut -c1-62 gpurun_out/r02_perf_gemm_t0.log) <(cut -c45-62 gpurun_out/r02_perf_gemm_tauto.log) <(cut -c45-62 gpurun_out/r02_perf_gemm_tc.log) | grep -v "N64 \|M6720\|R320"
for c in 0 auto 0 auto; do
  if [ $c = auto ]; then unset RB_GEMM_TALL; export RB_GEMM_CLUSTER=0; else export RB_GEMM_TALL=0; export RB_GEMM_CLUSTER=0; fi
  REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 timeout 300 python bench.py --steps 20 --warmup 5 --windows 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench tall=$c', d['value'], d['e2e']['value'], d['ms_per_step'])"
done
