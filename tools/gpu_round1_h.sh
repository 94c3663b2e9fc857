#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_igemm -c 170 \
  -o gpurun_out/h_igemm_full python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/h_ncu_full.log 2>&1
tail -2 gpurun_out/h_ncu_full.log; ls -la gpurun_out/h_igemm_full.ncu-rep
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:k_bn_ -c 120 \
  -o gpurun_out/h_bn_full python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/h_ncu_bn.log 2>&1
ls -la gpurun_out/h_bn_full.ncu-rep
timeout 600 python bench.py > gpurun_out/h_bench_default.json 2> gpurun_out/h_bench_default.err
echo "bench exit $?"; tail -3 gpurun_out/h_bench_default.err; cat gpurun_out/h_bench_default.json
