#!/bin/bash
mkdir -p gpurun_out
for c in 0 1; do
RB_GEMM_CLUSTER=$c timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -f -o gpurun_out/r02_gemm_shapes_cl$c python tools/prof_gemm_shapes2.py > gpurun_out/r02_ncu_gemm_cl$c.log 2>&1; echo "ncu cl=$c rc=$?"
ncu -i gpurun_out/r02_gemm_shapes_cl$c.ncu-rep --page raw --csv > gpurun_out/r02_gemm_shapes_cl$c.raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep gpurun_out/*.raw.csv
