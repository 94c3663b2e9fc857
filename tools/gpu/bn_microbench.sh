#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bn_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/bn_pytest_bn.log 2>&1; tail -4 gpurun_out/bn_pytest_bn.log
: > gpurun_out/bn_bn.jsonl
timeout 300 python tools/bench_bn_layer.py --tag fused >> gpurun_out/bn_bn.jsonl 2> gpurun_out/bn_err1.log
RIGL_BN_FUSED=0 timeout 300 python tools/bench_bn_layer.py --tag unfused >> gpurun_out/bn_bn.jsonl 2> gpurun_out/bn_err0.log
python - <<'PY'
import json
rows={}
for l in open('gpurun_out/bn_bn.jsonl'):
    d=json.loads(l); rows.setdefault((d['rows'],d['c'],d['op']),{})[d['tag']]=d
for k,v in rows.items():
    f=v.get('fused',{}); u=v.get('unfused',{})
    print(k, 'MB', f.get('tensor_mb'), 'fused us', f.get('us'), 'GB/s', f.get('gbps'), '| unfused us', u.get('us'), 'GB/s', u.get('gbps'))
PY
tail -3 gpurun_out/bn_err1.log
for fz in 1 0; do
RIGL_BN_FUSED=$fz timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/bn_layers$fz.json > gpurun_out/bn_bench$fz.json 2> gpurun_out/bn_bench$fz.err
echo "fused=$fz bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bn_bench$fz.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'])"
done
