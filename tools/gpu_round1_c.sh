#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/c_smoke.log 2>&1; tail -3 gpurun_out/c_smoke.log
timeout 600 python tools/debug_tc.py --big > gpurun_out/c_debug_tc.log 2>&1; grep -E "BAD|EXC|DEBUG_TC|\(3, 64|224" gpurun_out/c_debug_tc.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=15 > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -15 gpurun_out/c_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --cpu-batch 8 --layer-report gpurun_out/c_layers.json > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench exit $?"; tail -5 gpurun_out/c_bench.err; cat gpurun_out/c_bench.json
