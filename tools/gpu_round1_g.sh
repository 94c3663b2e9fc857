#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_tc.py --big > gpurun_out/g_debug_tc.log 2>&1; grep -E "BAD|EXC|DEBUG_TC" gpurun_out/g_debug_tc.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=15 > gpurun_out/g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/g_pytest.log
tail -12 gpurun_out/g_pytest.log
timeout 300 python tools/bench_mask_update.py > gpurun_out/g_bench_mask.json 2> gpurun_out/g_bench_mask.err; cat gpurun_out/g_bench_mask.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/g_layers.json > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench exit $?"; tail -5 gpurun_out/g_bench.err; cat gpurun_out/g_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/g_launches_step.csv python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/g_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv \
  --log-file gpurun_out/g_launches_mask.csv python tools/bench_mask_update.py --iters 2 --warmup 1 > gpurun_out/g_ncu2.log 2>&1
