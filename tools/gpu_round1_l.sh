#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bn_gpu.py tests/test_train_step_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/l_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/l_pytest.log; tail -3 gpurun_out/l_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/l_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'])"
