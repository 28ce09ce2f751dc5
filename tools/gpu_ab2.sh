#!/bin/bash
mkdir -p gpurun_out
REFTR_B200_LIB=$PWD/build/base/libreftr_b200.so timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_base.log 2>&1
RB_GEMM_TALL=0 timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_cur_t0.log 2>&1
timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_cur.log 2>&1
paste -d"|" <(cut -c1-62 gpurun_out/r02_perf_base.log) <(cut -c45-62 gpurun_out/r02_perf_cur_t0.log) <(cut -c45-62 gpurun_out/r02_perf_cur.log) | grep -v "R320"
run() {  name=$1; lib=$2; shift 2
  env REFTR_B200_LIB=$lib REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 "$@" timeout 300 python bench.py --steps 20 --warmup 5 --windows 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['e2e']['value'],1), d['windows_ms_per_step'])"
}
for rep in 1 2; do
  run base $PWD/build/base/libreftr_b200.so X=1
  run cur_tall0 $PWD/reftr_b200/libreftr_b200.so RB_GEMM_TALL=0
  run cur $PWD/reftr_b200/libreftr_b200.so X=1
done
