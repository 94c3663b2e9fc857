#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_tc.py --big > gpurun_out/m_debug_tc.log 2>&1; echo "debug exit $?"; grep -E "BAD|EXC|DEBUG_TC|fatal" gpurun_out/m_debug_tc.log | head -40
if grep -q "DEBUG_TC bad=0" gpurun_out/m_debug_tc.log; then
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/m_layers.json > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
  echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/m_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'])"
  timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=10 > gpurun_out/m_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/m_pytest.log; tail -4 gpurun_out/m_pytest.log
else
  echo "pair kernel failed; rerun with RIGL_CTA_PAIR=0 for a sanity check"
  RIGL_CTA_PAIR=0 timeout 300 python tools/debug_tc.py > gpurun_out/m_debug_tc_nopair.log 2>&1; grep -E "DEBUG_TC" gpurun_out/m_debug_tc_nopair.log
fi
