#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_tc.py --big > gpurun_out/p_debug_tc.log 2>&1; echo "debug exit $?"; grep -E "BAD|EXC|DEBUG_TC|fatal" gpurun_out/p_debug_tc.log | head -30
grep -E "^\((2, 14, 14|2, 28, 28, 32|2, 13|5, 6|4, 56, 56, 64, 64|32, 56, 56, 64, 64)" gpurun_out/p_debug_tc.log
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "halo" > gpurun_out/p_pytest_halo.log 2>&1; tail -15 gpurun_out/p_pytest_halo.log
if grep -q "DEBUG_TC bad=0" gpurun_out/p_debug_tc.log; then
for halo in 1 0; do
  RIGL_HALO3X3=$halo timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/p_layers_halo$halo.json > gpurun_out/p_bench_halo$halo.json 2> gpurun_out/p_bench_halo$halo.err
  echo "halo=$halo bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/p_bench_halo$halo.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'])"
done
fi
