#!/bin/bash
mkdir -p gpurun_out
REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0 REFTR_B200_BENCH_GROUPS=gpurun_out/r02_gemm_groups_final.txt timeout 300 python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_roof.json 2> gpurun_out/r02_bench_roof.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_roof.json") if l.startswith("{")][-1])
r=d["roofline"]; print(round(d["value"],1), {k:r[k] for k in ("achieved","frac","launches","gemm_ms_per_step","avg_launch_us","gemm_gflop_per_step")}); print(r["top_groups"])
P
tail -3 gpurun_out/r02_bench_roof.err
