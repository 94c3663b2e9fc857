#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_shift_probe tools/umma_shift_probe.cu && timeout 60 /tmp/umma_shift_probe > gpurun_out/o_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/o_probe.log
timeout 300 python -m pytest tests/test_bn_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/o_pytest_bn.log 2>&1; tail -3 gpurun_out/o_pytest_bn.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/o_layers.json > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/o_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'], d['clocks'])"
