#!/bin/bash
# (hi | residual) weight pairs in the forward of the encoder / BERT: box error by stage (full-size cfg2, B=16) and step time
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
for h in "" "enc" "enc,bert"; do
  echo "== REFTR_B200_HILO='$h' B=16"
  PB=16 REFTR_B200_HILO="$h" timeout 600 python tools/parity_stages.py 2>&1 | grep "memory\|boxes layer . *:\|C5"
done
for h in "" "enc,bert" "enc" ""  "enc,bert"; do
  REFTR_B200_HILO="$h" timeout 300 python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_hilo.json 2> gpurun_out/r02_bench_hilo.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_hilo.json") if l.startswith("{")][-1])
print("hilo '$h'", round(d["value"],1), round(d["e2e"]["value"],1), d["windows_ms_per_step"])
P
done
