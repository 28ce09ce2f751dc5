#!/bin/bash
# end-of-round check on one GPU: the whole GPU test suite, smoke(), the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final2.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_final2.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_final2.json") if l.startswith("{")][-1])
print(d["value"], d["e2e"]["value"], d["windows_ms_per_step"], "roofline.frac", round(d["roofline"]["frac"],4), "stale", d["encoder_mha"]["stale"],
      "x stock", round(d["stock_gpu_baseline"]["ratio_ours_over_stock"],2), "parity max", d["parity"]["max"], "launches", d["gpu_launches"])
P
