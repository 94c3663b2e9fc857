#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_whole_step_parity_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4
bash tools/gpu/r2_g.sh
