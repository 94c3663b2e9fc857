#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "graph" > gpurun_out/ov_pytest.log 2>&1; tail -8 gpurun_out/ov_pytest.log
for cfg in "1 1" "1 0" "0 1"; do
  set -- $cfg
  RIGL_WGRAD_OVERLAP=$1 RIGL_BN_FUSED=$2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ov_bench_$1_$2.json 2> gpurun_out/ov_bench_$1_$2.err
  echo "overlap=$1 bnfused=$2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/ov_bench_$1_$2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'])"
done
