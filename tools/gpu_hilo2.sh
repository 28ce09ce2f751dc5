#!/bin/bash
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
for cfg in "enc,bert 0" "enc,bert 36" "enc,bert 48" "enc,bert 72" "enc,bert 24" "enc,bert 0"; do
  set -- $cfg
  REFTR_B200_HILO="$1" REFTR_B200_BERT_SMS_FWD=$2 timeout 300 python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_hilo.json 2> gpurun_out/r02_bench_hilo.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_hilo.json") if l.startswith("{")][-1])
print("hilo '$1' bert fwd sm limit $2:", round(d["value"],1), round(d["e2e"]["value"],1), d["windows_ms_per_step"])
P
done
