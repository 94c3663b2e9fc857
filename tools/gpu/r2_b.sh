#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_whole_step_parity_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2b_whole.log 2>&1; echo "whole exit $?"; tail -25 gpurun_out/r2b_whole.log
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_whole_step_parity_gpu.py --durations=15 ) > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/r2b_layers.json > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench exit $?"; cat gpurun_out/r2b_bench.json | head -c 3000
