#!/bin/bash
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
for h in "enc,bert" "enc,bert_qkv,bert_o,bert_ffn1"; do
  echo "== REFTR_B200_HILO='$h' B=16"
  PB=16 REFTR_B200_HILO="$h" timeout 600 python tools/parity_stages.py 2>&1 | grep "memory\|boxes layer . *:"
done
for h in "enc,bert" "enc,bert_qkv,bert_o,bert_ffn1" "enc,bert" "enc,bert_qkv,bert_o,bert_ffn1"; do
  REFTR_B200_HILO="$h" timeout 300 python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_hilo.json 2> gpurun_out/r02_bench_hilo.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_hilo.json") if l.startswith("{")][-1])
print("hilo '$h'", round(d["value"],1), d["windows_ms_per_step"])
P
done
