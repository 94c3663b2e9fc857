#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_tc.py --big > gpurun_out/conv_debug_tc.log 2>&1; echo "debug exit $?"; grep -E "BAD|EXC|DEBUG_TC|fatal" gpurun_out/conv_debug_tc.log | head -30
if grep -q "DEBUG_TC bad=0" gpurun_out/conv_debug_tc.log; then
: > gpurun_out/conv_layer.jsonl
timeout 300 python tools/bench_conv_layer.py --shapes r50_3x3 --tag "prod" >> gpurun_out/conv_layer.jsonl 2>> gpurun_out/conv_err.log
timeout 300 python tools/bench_conv_layer.py --shapes r50_1x1 --tag "prod" >> gpurun_out/conv_layer.jsonl 2>> gpurun_out/conv_err.log
RIGL_CTA_PAIR=0 timeout 300 python tools/bench_conv_layer.py --shapes r50_3x3 --tag "prod_nopair" >> gpurun_out/conv_layer.jsonl 2>> gpurun_out/conv_err.log
RIGL_CTA_PAIR=0 timeout 300 python tools/bench_conv_layer.py --shapes r50_1x1 --tag "prod_nopair" >> gpurun_out/conv_layer.jsonl 2>> gpurun_out/conv_err.log
python - <<'PY'
import json
for l in open('gpurun_out/conv_layer.jsonl'):
    d=json.loads(l); print(d['tag'], d['shape'], d['op'], d['us'], d['tflops_dense'], d['gbps_algorithmic'])
PY
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/conv_layers.json > gpurun_out/conv_bench.json 2> gpurun_out/conv_bench.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/conv_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'])"
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=10 > gpurun_out/conv_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/conv_pytest.log; tail -4 gpurun_out/conv_pytest.log
fi
