#!/bin/bash
mkdir -p gpurun_out
# 1) correctness of the forked-gradient BN path + whole test-suite subset
timeout 600 python -m pytest tests/test_bn_gpu.py tests/test_train_step_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/v_pytest.log 2>&1; tail -5 gpurun_out/v_pytest.log
# 2) bench with / without the fork
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/v_layers.json > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/v_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'])"
# 3) where do the pair kernel's warps wait?  (source-level stall samples)
for pair in 1 0; do
  RIGL_CTA_PAIR=$pair timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_igemm_kmajor -s 7 -c 1 -f -o gpurun_out/v_c3_pair$pair python tools/bench_conv_layer.py --shapes r50_33c3 --iters 1 > gpurun_out/v_ncu_pair$pair.log 2>&1
  ncu -i gpurun_out/v_c3_pair$pair.ncu-rep --page source --csv > gpurun_out/v_c3_pair$pair.source.csv 2>/dev/null
  ncu -i gpurun_out/v_c3_pair$pair.ncu-rep --page raw --csv > gpurun_out/v_c3_pair$pair.raw.csv 2>/dev/null
  ls -la gpurun_out/v_c3_pair$pair.*
done
