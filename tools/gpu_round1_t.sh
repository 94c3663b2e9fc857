#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/rate_probe tools/sm100_rate_probe.cu && timeout 200 /tmp/rate_probe > gpurun_out/t_rate_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/t_rate_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/t_launches.csv python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/t_step.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/t_launches.csv
