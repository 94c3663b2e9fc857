#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/q_layer.jsonl
for cfg in "2,2" "1,3" "1,4" "1,2"; do
  RIGL_HALO_CFG=$cfg timeout 200 python tools/bench_conv_layer.py --shapes r50s1 --tag "halo_cfg=$cfg" >> gpurun_out/q_layer.jsonl 2> gpurun_out/q_err_$cfg.log
done
RIGL_HALO3X3=0 timeout 200 python tools/bench_conv_layer.py --shapes r50s1 --tag "generic" >> gpurun_out/q_layer.jsonl 2>> gpurun_out/q_err.log
timeout 300 python tools/bench_conv_layer.py --shapes r50_3x3 --tag "default" >> gpurun_out/q_layer.jsonl 2>> gpurun_out/q_err.log
timeout 300 python tools/bench_conv_layer.py --shapes r50_1x1 --tag "default" >> gpurun_out/q_layer.jsonl 2>> gpurun_out/q_err.log
cat gpurun_out/q_layer.jsonl
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "halo" 2>&1 | tail -3
