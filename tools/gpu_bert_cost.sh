#!/bin/bash
# what the BERT branch's backward costs inside the step: default vs BERT frozen (no BERT backward)
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
for v in none lang_backbone; do
  f=$v; [ $v = none ] && f=""
  REFTR_B200_BENCH_FREEZE="$f" python bench.py --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_frz_$v.json 2> gpurun_out/r02_bench_frz_$v.err
done
python - <<'P'
import json
for f in ("none","lang_backbone"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02_bench_frz_{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["launches"], d["gpu_launches"])
    except Exception as e: print(f, "failed", e)
P
