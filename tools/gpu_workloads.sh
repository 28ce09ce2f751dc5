#!/bin/bash
mkdir -p gpurun_out
for w in cfg3 cfg4 cfg5; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/bench_$w.log 2>&1; echo "$w rc=$?"
tail -3 gpurun_out/bench_$w.log | grep -o '"value": [0-9.]*, "ms_per_step": [0-9.]*' || tail -5 gpurun_out/bench_$w.log | cut -c1-400
done
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:umma_gemm_kernel -s 52 -c 14 -f -o gpurun_out/prof_final_gemm2 python tools/profile_step.py > gpurun_out/ncu_final_gemm2.log 2>&1; echo "gemm2 rc=$?"
