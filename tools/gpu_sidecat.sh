#!/bin/bash
# A/B: one side-stream chain for all off-critical-path work (round 1) vs one chain per category (transformer / BERT / conv backbone)
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
timeout 900 python -m pytest tests/test_e2e_gpu.py -x -q -k "cfg1_box or split" > gpurun_out/r02_pytest_sidecat.log 2>&1; tail -3 gpurun_out/r02_pytest_sidecat.log
for rep in 1 2; do
for v in 0 1; do
  REFTR_B200_BRANCH_PRIORITY=$v python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_sidecat$v.json 2> gpurun_out/r02_bench_sidecat$v.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_sidecat$v.json") if l.startswith("{")][-1])
print("priority=$v", round(d["value"],1), round(d["e2e"]["value"],1), d["windows_ms_per_step"])
P
done
done
