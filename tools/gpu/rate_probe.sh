#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/rate_probe tools/sm100_rate_probe.cu && timeout 200 /tmp/rate_probe > gpurun_out/rate_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/rate_probe.log
