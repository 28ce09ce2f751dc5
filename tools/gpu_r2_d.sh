#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r2_profiles.sh
unset REFTR_B200_SIDE_STREAM
for sc in 1024 16384 65536; do
  REFTR_B200_GRAD_SCALE=$sc REFTR_B200_DYNAMIC_SCALE=0 timeout 600 python -m pytest tests/test_full_size_gpu.py -m gpu -q -s -k "cfg2 or cfg5" 2>&1 | grep "gradient rel-L2\|passed\|failed" | sed "s/^/scale $sc: /" | cut -c1-330
done
