#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench$i.json 2> gpurun_out/r2o_bench$i.err; python -c "
import json; d=json.load(open('gpurun_out/r2o_bench$i.json')); print('run$i', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['step_ms']['max'])" || tail -5 gpurun_out/r2o_bench$i.err
done
timeout 200 python -m pytest tests/test_mask_update_gpu.py -m gpu -q -p no:cacheprovider -k "three_optimizer" 2>&1 | tail -2
