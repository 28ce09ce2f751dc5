#!/bin/bash
# ncu --set full captures of the dominant kernels inside one eager step (cudaProfilerStart/Stop gated)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:umma_gemm_kernel -s 3 -c 6 -f -o gpurun_out/prof_gemm python tools/profile_step.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_.*_tc_kernel -c 3 -f -o gpurun_out/prof_attn python tools/profile_step.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
