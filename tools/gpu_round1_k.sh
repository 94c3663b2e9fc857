#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=10 > gpurun_out/k_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/k_pytest.log; tail -6 gpurun_out/k_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/k_layers.json > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
echo "bench exit $?"; grep -i "error\|Traceback" -A3 gpurun_out/k_bench.err | tail -8; cat gpurun_out/k_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/k_launches_step.csv python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/k_ncu.log 2>&1
