#!/bin/bash
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
timeout 600 python -m pytest tests/test_e2e_gpu.py -x -q -k "cfg1_box or multi_phrase or split" 2>&1 | tail -2
for v in 0 1 0 1; do
  REFTR_B200_FORKS=$v timeout 300 python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_prio.json 2> gpurun_out/r02_bench_prio.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_prio.json") if l.startswith("{")][-1])
print("forks $v", round(d["value"],1), round(d["e2e"]["value"],1), d["windows_ms_per_step"])
P
done
