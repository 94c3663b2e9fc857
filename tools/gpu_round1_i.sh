#!/bin/bash
# Correctness of the new kernels (multicast clusters, small-Cin stem path) + bench + SMALL ncu full captures
mkdir -p gpurun_out
timeout 300 python tools/debug_tc.py --big > gpurun_out/i_debug_tc.log 2>&1; grep -E "BAD|EXC|DEBUG_TC|fatal" gpurun_out/i_debug_tc.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=15 > gpurun_out/i_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/i_pytest.log
tail -8 gpurun_out/i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/i_layers.json > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench exit $?"; tail -5 gpurun_out/i_bench.err; cat gpurun_out/i_bench.json
RIGL_CLUSTER_MC=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/i_bench_nomc.json 2> gpurun_out/i_bench_nomc.err
echo "bench (no multicast) exit $?"; cut -c1-330 gpurun_out/i_bench_nomc.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/i_launches_step.csv python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/i_ncu.log 2>&1
