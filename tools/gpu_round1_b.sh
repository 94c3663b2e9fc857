#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/debug_tc.py --big > gpurun_out/b_debug_tc.log 2>&1
echo "debug_tc exit $?" >> gpurun_out/b_debug_tc.log
cat gpurun_out/b_debug_tc.log | tail -45
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x -k "not mask_update and not masks_gpu" > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
tail -30 gpurun_out/b_pytest.log
