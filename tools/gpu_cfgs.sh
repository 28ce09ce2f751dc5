#!/bin/bash
# the other BASELINE configs with the final library (not the metric): cfg3 mask head bs8, cfg4 Flickr multi-phrase bs32, cfg5 ResNet-101 800x800 bs16
mkdir -p gpurun_out
export REFTR_B200_BENCH_STOCK=0 REFTR_B200_BENCH_OPTIM=0
for c in cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $c --no-cpu-baseline --windows 3 > gpurun_out/r02_bench_$c.json 2> gpurun_out/r02_bench_$c.err
  python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_$c.json") if l.startswith("{")][-1])
    print("$c", round(d["value"],1), round(d["e2e"]["value"],1), d["windows_ms_per_step"], d["config"]["global_batch"])
except Exception as e:
    print("$c failed", e); print(open("gpurun_out/r02_bench_$c.err").read()[-600:])
P
done
