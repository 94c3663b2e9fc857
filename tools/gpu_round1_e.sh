#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=15 -k "bn_gpu or train_step or conv_gpu" > gpurun_out/e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -25 gpurun_out/e_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/e_layers.json > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo "bench exit $?"; tail -5 gpurun_out/e_bench.err; cat gpurun_out/e_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/e_launches_step.csv python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/e_ncu.log 2>&1
tail -2 gpurun_out/e_ncu.log
