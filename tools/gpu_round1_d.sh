#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/d_launches_step.csv python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/d_ncu.log 2>&1
tail -2 gpurun_out/d_ncu.log; wc -l gpurun_out/d_launches_step.csv
