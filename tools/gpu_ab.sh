#!/bin/bash
# same-box A/B of the whole step: previous commit's kernels | current kernels without the attention occupancy change (tall off / tall auto) | current
mkdir -p gpurun_out
run() {  # name, lib, env...
  name=$1; lib=$2; shift 2
  env REFTR_B200_LIB=$lib REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 "$@" timeout 300 python bench.py --steps 20 --warmup 5 --windows 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['e2e']['value'],1), d['windows_ms_per_step'])"
}
for rep in 1 2; do
  run base $PWD/build/base/libreftr_b200.so X=1
  run cur_noattn_tall0 $PWD/build/varB/libreftr_b200.so RB_GEMM_TALL=0
  run cur_noattn_tallauto $PWD/build/varB/libreftr_b200.so X=1
  run cur $PWD/reftr_b200/libreftr_b200.so X=1
done
REFTR_B200_LIB=$PWD/reftr_b200/libreftr_b200.so timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x 2>&1 | tail -2
