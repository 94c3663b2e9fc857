#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_tc.py --big > gpurun_out/n_debug_tc.log 2>&1; echo "debug exit $?"; grep -E "BAD|EXC|DEBUG_TC|fatal" gpurun_out/n_debug_tc.log | head -20
for pair in 1 0; do
  RIGL_CTA_PAIR=$pair timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/n_layers_pair$pair.json > gpurun_out/n_bench_pair$pair.json 2> gpurun_out/n_bench_pair$pair.err
  echo "pair=$pair bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/n_bench_pair$pair.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'])"
done
