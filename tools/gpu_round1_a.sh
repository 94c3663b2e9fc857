#!/bin/bash
# First GPU call: mask path parity + SIMT conv parity + mask-update microbench + ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=40 > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
tail -40 gpurun_out/a_pytest.log
timeout 300 python tools/bench_mask_update.py > gpurun_out/a_bench_mask.json 2> gpurun_out/a_bench_mask.err
timeout 300 python tools/bench_mask_update.py --noise >> gpurun_out/a_bench_mask.json 2>> gpurun_out/a_bench_mask.err
cat gpurun_out/a_bench_mask.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv \
  --log-file gpurun_out/a_launches_mask.csv python tools/bench_mask_update.py --iters 2 --warmup 1 > gpurun_out/a_ncu.log 2>&1
tail -3 gpurun_out/a_ncu.log
