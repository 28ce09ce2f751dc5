#!/bin/bash
# dynamic tile scheduler: correctness, then A/B against the previous (static) kernels on the same box
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_dyn.log 2>&1; rc=$?; echo "gemm tests rc=$rc"; tail -4 gpurun_out/r02_pytest_dyn.log | cut -c1-300
if [ $rc -ne 0 ]; then exit 0; fi
RB_GEMM_CLUSTER=1 timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x 2>&1 | tail -1
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_dropout_gpu.py tests/test_seg_kernels_gpu.py tests/test_bert_kernels_gpu.py tests/test_kernels_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_dyn2.log 2>&1; echo "e2e tests rc=$?"; tail -3 gpurun_out/r02_pytest_dyn2.log | cut -c1-300
REFTR_B200_LIB=$PWD/build/base/libreftr_b200.so timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_base.log 2>&1
timeout 300 python tools/perf_gemm.py > gpurun_out/r02_perf_dyn.log 2>&1
paste -d"|" <(cut -c1-62 gpurun_out/r02_perf_base.log) <(cut -c45-62 gpurun_out/r02_perf_dyn.log) | grep -v "R320"
run() {  name=$1; lib=$2; shift 2
  env REFTR_B200_LIB=$lib REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 "$@" timeout 300 python bench.py --steps 20 --warmup 5 --windows 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['e2e']['value'],1), d['windows_ms_per_step'])"
}
for rep in 1 2; do
  run base $PWD/build/base/libreftr_b200.so X=1
  run dyn $PWD/reftr_b200/libreftr_b200.so X=1
done
