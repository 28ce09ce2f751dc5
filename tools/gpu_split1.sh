#!/bin/bash
# one GPU: single-graph backward vs the split backward (parts handed over while later parts compute; no exchange)
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
for v in "" 1 "" 1; do
  REFTR_B200_SPLIT_BWD=$v timeout 300 python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_split1.json 2> gpurun_out/r02_bench_split1.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_split1.json") if l.startswith("{")][-1])
print("SPLIT_BWD='$v'", round(d["value"],1), round(d["e2e"]["value"],1), d["windows_ms_per_step"])
P
done
