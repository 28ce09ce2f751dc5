#!/bin/bash
# round evidence: tests, stock-GPU baseline, launch list (durations), DRAM bytes of every GEMM launch, full captures of the main kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl stock-gpu --steps 10 --warmup 3 > gpurun_out/bench_stock_gpu.log 2>&1; echo "stock rc=$?"; tail -1 gpurun_out/bench_stock_gpu.log | cut -c1-400
export REFTR_B200_SIDE_STREAM=0   # serialise the branches so that per-kernel numbers are not perturbed by overlap
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final.csv python tools/profile_step.py > gpurun_out/launches_final.log 2>&1; echo "launches rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k regex:umma_gemm_kernel --csv --log-file gpurun_out/gemm_dram_final.csv python tools/profile_step.py > gpurun_out/gemm_dram_final.log 2>&1; echo "dram rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_fwd_tc|attn_bwd_dq_tc|attn_bwd_dkv_tc|stem_conv" -c 4 -f -o gpurun_out/prof_final_misc python tools/profile_step.py > gpurun_out/ncu_final_misc.log 2>&1; echo "misc rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:umma_gemm_kernel -s 2 -c 12 -f -o gpurun_out/prof_final_gemm python tools/profile_step.py > gpurun_out/ncu_final_gemm.log 2>&1; echo "gemm rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*final*.csv
