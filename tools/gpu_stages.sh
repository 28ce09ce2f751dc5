#!/bin/bash
mkdir -p gpurun_out
for f in 1 0; do
  echo "== REFTR_B200_STEM_FUSED=$f B=16"
  PB=16 REFTR_B200_STEM_FUSED=$f timeout 600 python tools/parity_stages.py 2>&1 | tail -21
done
