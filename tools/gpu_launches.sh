#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/launches.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/launches.log; wc -l gpurun_out/launches.csv
