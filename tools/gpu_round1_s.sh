#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/rate_probe tools/sm100_rate_probe.cu && timeout 120 /tmp/rate_probe > gpurun_out/s_rate_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/s_rate_probe.log
for k in k_halo3x3_kmajor k_halo3x3_wgrad; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -f -o gpurun_out/s_$k python tools/bench_conv_layer.py --shapes r50s1 --iters 1 > gpurun_out/s_ncu_$k.log 2>&1
  ncu -i gpurun_out/s_$k.ncu-rep --page raw --csv > gpurun_out/s_$k.raw.csv 2>/dev/null
  ls -la gpurun_out/s_$k.ncu-rep
done
timeout 400 ncu --set full --clock-control none -k regex:k_igemm_kmajor2 -s 2 -c 2 -f -o gpurun_out/s_kmajor2 python tools/bench_conv_layer.py --shapes r50_1x1 --iters 1 > gpurun_out/s_ncu_kmajor2.log 2>&1
ncu -i gpurun_out/s_kmajor2.ncu-rep --page raw --csv > gpurun_out/s_kmajor2.raw.csv 2>/dev/null
ls -la gpurun_out/
