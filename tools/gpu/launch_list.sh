#!/bin/bash
# One ResNet-50 ERK-80 b256 train step under ncu: per-launch time + DRAM bytes (-> profiles/*_step_launches_*.md).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/step_for_ncu.py --steps 1 --warmup 2 \
  > gpurun_out/step.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
