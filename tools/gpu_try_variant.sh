#!/bin/bash
# First GPU call for an experimental build variant (tools/build_variant.sh must have been run HERE first: nvcc cross-compiles, the
# built library travels with the snapshot).  Usage on the GPU box:  bash tools/gpu_try_variant.sh var16
# 1. correctness of the variant: GEMM / kernel / e2e parity tests against the variant library
# 2. per-shape GEMM timings, default vs variant, side by side
# 3. bench.py with both libraries
name=${1:-var16}
LIBV=$PWD/build/$name/libreftr_b200.so
mkdir -p gpurun_out
[ -f "$LIBV" ] || { echo "missing $LIBV (run tools/build_variant.sh $name ... before gpurun)"; exit 1; }
REFTR_B200_LIB=$LIBV timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_e2e_gpu.py tests/test_seg_kernels_gpu.py -x -q > gpurun_out/pytest_$name.log 2>&1
echo "variant tests rc=$?"; tail -3 gpurun_out/pytest_$name.log
python tools/perf_gemm.py > gpurun_out/perf_gemm_default.log 2>&1
REFTR_B200_LIB=$LIBV python tools/perf_gemm.py > gpurun_out/perf_gemm_$name.log 2>&1
paste -d"|" <(cut -c1-62 gpurun_out/perf_gemm_default.log) <(cut -c45-62 gpurun_out/perf_gemm_$name.log)
for lib in "" "$LIBV"; do
  REFTR_B200_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('${lib:-default}', d['value'], d['e2e']['value'], d['ms_per_step'])"
done
