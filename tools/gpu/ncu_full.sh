#!/bin/bash
# ncu --set full of the top kernels of the C2 step (one GPU; ~40 replays per captured launch)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_igemm_kmajor2 -s 20 -c 3 \
  -o gpurun_out/r2_ncu_kmajor2 -f python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/r2_ncu_kmajor2.log 2>&1; echo "ncu kmajor2 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_bn_colsum -s 60 -c 2 \
  -o gpurun_out/r2_ncu_bn -f python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/r2_ncu_bn.log 2>&1; echo "ncu bn exit $?"
ls -la gpurun_out/*.ncu-rep
