#!/bin/bash
# Builds reftr_b200/libreftr_b200.so for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libreftr_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr"
mkdir -p ../../build
objs=""
pids=""
for f in *.cu; do
  o=../../build/${f%.cu}.o
  objs="$objs $o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ host.h -nt "$o" ] || [ ../../include/reftr_b200.h -nt "$o" ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -shared -o $OUT $objs -cudart static
echo "built $(realpath $OUT)"
