#!/bin/bash
# ncu launch list (durations) of one eager train-mode step, branches serialised
mkdir -p gpurun_out
export REFTR_B200_SIDE_STREAM=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python tools/profile_step.py > gpurun_out/launches_train.log 2>&1; echo "launches rc=$?"
python tools/summarize_launches.py gpurun_out/launches_train.csv 45
