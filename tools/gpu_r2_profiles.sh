#!/bin/bash
# round-2 evidence: launch list (durations), DRAM bytes of every GEMM launch, full captures of the attention / stem kernels and of a few GEMMs
mkdir -p gpurun_out
export REFTR_B200_SIDE_STREAM=0   # serialise the branches so that per-kernel numbers are not perturbed by overlap
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py > gpurun_out/r02_launches.log 2>&1; echo "launches rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches.csv 45 > gpurun_out/r02_launches_summary.txt; head -30 gpurun_out/r02_launches_summary.txt | cut -c1-150
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k regex:"umma_gemm_kernel|gemm_skinny" --csv --log-file gpurun_out/r02_gemm_dram.csv python tools/profile_step.py > gpurun_out/r02_gemm_dram.log 2>&1; echo "dram rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_fwd_tc|attn_bwd_dq_tc|attn_bwd_dkv_tc|stem_pool" -c 4 -f -o gpurun_out/r02_prof_attn python tools/profile_step.py > gpurun_out/r02_ncu_attn.log 2>&1; echo "attn rc=$?"
python tools/ncu_summary.py gpurun_out/r02_prof_attn.ncu-rep gpurun_out/r02_ncu_full_attention_stem.csv > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm_kernel -f -o gpurun_out/r02_prof_gemm python tools/prof_gemm_shapes2.py > gpurun_out/r02_ncu_gemm.log 2>&1; echo "gemm rc=$?"
python tools/ncu_summary.py gpurun_out/r02_prof_gemm.ncu-rep gpurun_out/r02_ncu_full_gemm_shapes.csv > /dev/null 2>&1
ls -la gpurun_out/r02_*.csv gpurun_out/r02_*summary*
