#!/bin/bash
# round 2, call 1: the two experimental variants left un-run by round 1 (-DRB_EPI16, -DRB_ATTN_FAST): parity + timing, side by side
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for name in var16 varA; do
  LIBV=$PWD/build/$name/libreftr_b200.so
  REFTR_B200_LIB=$LIBV timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_e2e_gpu.py -x -q > gpurun_out/r02_pytest_$name.log 2>&1
  echo "$name tests rc=$?"; tail -3 gpurun_out/r02_pytest_$name.log
done
python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_default.log 2>&1
REFTR_B200_LIB=$PWD/build/var16/libreftr_b200.so python tools/perf_gemm.py > gpurun_out/r02_perf_gemm_var16.log 2>&1
paste -d"|" <(cut -c1-62 gpurun_out/r02_perf_gemm_default.log) <(cut -c45-62 gpurun_out/r02_perf_gemm_var16.log)
for lib in "" "$PWD/build/var16/libreftr_b200.so" "$PWD/build/varA/libreftr_b200.so"; do
  REFTR_B200_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('${lib:-default}', d['value'], d['e2e']['value'], d['ms_per_step'])"
done
