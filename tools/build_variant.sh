#!/bin/bash
# Builds an experimental variant of the library next to the default one: tools/build_variant.sh <name> <nvcc defines...>
#   e.g. tools/build_variant.sh var16 -DRB_EPI16      (16 epilogue warps in the tcgen05 GEMM)
# and prints how to run the tests / benches against it (REFTR_B200_LIB overrides the library path, reftr_b200/_lib.py).
set -e
cd "$(dirname "$0")/../reftr_b200/csrc"
name=$1; shift
out=../../build/$name
mkdir -p $out
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
objs=""
for f in *.cu; do
  o=$out/${f%.cu}.o
  objs="$objs $o"
  nvcc $FLAGS "$@" -c "$f" -o "$o" &
done
wait
nvcc -shared -o $out/libreftr_b200.so $objs -cudart static
echo "built $(realpath $out/libreftr_b200.so)"
echo "run:  REFTR_B200_LIB=\$PWD/build/$name/libreftr_b200.so python -m pytest tests -m gpu -x -q   (and tools/perf_gemm.py, bench.py)"
