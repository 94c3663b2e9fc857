#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=15 > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -25 gpurun_out/c_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/c_smoke.log 2>&1; tail -3 gpurun_out/c_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --cpu-batch 8 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
echo "bench exit $?"; tail -5 gpurun_out/c_bench.err; cat gpurun_out/c_bench.json
