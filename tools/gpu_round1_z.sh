#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_tc.py --big > gpurun_out/z_debug_tc.log 2>&1; echo "debug exit $?"; grep -E "BAD|EXC|DEBUG_TC|fatal" gpurun_out/z_debug_tc.log | head -30
if grep -q "DEBUG_TC bad=0" gpurun_out/z_debug_tc.log; then
: > gpurun_out/z_layer.jsonl
for lb in 1 0; do
RIGL_PAIR_LOCALBAR=$lb timeout 300 python tools/bench_conv_layer.py --shapes r50_3x3 --tag "localbar=$lb" >> gpurun_out/z_layer.jsonl 2>> gpurun_out/z_err.log
RIGL_PAIR_LOCALBAR=$lb timeout 300 python tools/bench_conv_layer.py --shapes r50_1x1 --tag "localbar=$lb" >> gpurun_out/z_layer.jsonl 2>> gpurun_out/z_err.log
done
python - <<'PY'
import json
d={}
for l in open('gpurun_out/z_layer.jsonl'):
    r=json.loads(l); d.setdefault((tuple(r['shape']),r['op']),{})[r['tag']]=r['us']
for k,v in d.items():
    if k[1]!='wgrad': print(k, v)
PY
for lb in 1 0; do
RIGL_PAIR_LOCALBAR=$lb timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/z_bench$lb.json 2> gpurun_out/z_bench$lb.err
echo "localbar=$lb bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/z_bench$lb.json')); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_step_by_kind'])"
done
fi
