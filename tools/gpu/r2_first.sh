#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/sw32 tools/umma_sw32_probe.cu -lcuda > gpurun_out/r2a_probe.log 2>&1 && timeout 60 /tmp/sw32 >> gpurun_out/r2a_probe.log 2>&1; echo "probe exit $?"; tail -20 gpurun_out/r2a_probe.log
RIGL_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_mask_update_gpu.py tests/test_optimizers_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2a_pytest_exp1.log 2>&1; echo "exit $?"; tail -30 gpurun_out/r2a_pytest_exp1.log
RIGL_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k s2d > gpurun_out/r2a_pytest_s2d.log 2>&1; echo "exit $?"; tail -40 gpurun_out/r2a_pytest_s2d.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
