#!/bin/bash
# A/B: SM limits of BERT's backward GEMMs (input gradients / weight gradients) next to the conv backbone's backward
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
for cfg in "0 72 36 32 0" "0 72 36 32 1" "36 72 36 32 1" "24 72 36 32 0" "0 72 36 16 0" "0 72 36 48 0" "0 72 36 32 0"; do
  set -- $cfg
  REFTR_B200_BERT_SMS_FWD=$1 REFTR_B200_BERT_SMS_BWD=$2 REFTR_B200_BERT_SMS_WGRAD=$3 REFTR_B200_SIDE_SMS_T=$4 REFTR_B200_BRANCH_PRIORITY_FWD=$5 python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_bertsm.json 2> gpurun_out/r02_bench_bertsm.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_bertsm.json") if l.startswith("{")][-1])
print("fwd/bwd/wgrad/T limits, fwd priority $cfg:", round(d["value"],1), round(d["e2e"]["value"],1), d["windows_ms_per_step"])
P
done
