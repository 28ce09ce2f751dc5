#!/bin/bash
# attention forward at 3 (default) vs 4 CTAs per SM (80 registers, 276 B spills): kernel time + whole step, same box
mkdir -p gpurun_out
for lib in reftr_b200/libreftr_b200.so build/varC/libreftr_b200.so build/varB/libreftr_b200.so; do
  REFTR_B200_LIB=$PWD/$lib timeout 200 python tools/perf_attn.py 2>&1 | tail -6 | sed "s|^|$lib: |"
done
run() {  name=$1; lib=$2; shift 2
  env REFTR_B200_LIB=$lib REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 "$@" timeout 300 python bench.py --steps 20 --warmup 5 --windows 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['e2e']['value'],1), d['windows_ms_per_step'])"
}
for rep in 1 2; do
  run cur_fwd3 $PWD/reftr_b200/libreftr_b200.so X=1
  run cur_fwd4 $PWD/build/varC/libreftr_b200.so X=1
done
