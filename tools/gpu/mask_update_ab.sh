#!/bin/bash
# the batched mask update: bit-exactness tests on every kernel variant, then the micro-benchmark of variants 2 and 3
# (CUDA events) and the per-kernel launch list of variant 3
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_mask_update_gpu.py tests/test_optimizers_gpu.py tests/test_bn_gpu.py -q -m gpu -p no:cacheprovider -x --tb=short ) > gpurun_out/mu_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/mu_pytest.log
for v in 2 3; do
  timeout 300 python tools/bench_mask_update.py --variant $v > gpurun_out/mu_bench_v$v.json 2> gpurun_out/mu_bench_v$v.err; cat gpurun_out/mu_bench_v$v.json
  timeout 300 python tools/bench_mask_update.py --variant $v --inkernel-noise > gpurun_out/mu_bench_v${v}_noise.json 2>> gpurun_out/mu_bench_v$v.err; cat gpurun_out/mu_bench_v${v}_noise.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mu_launches_v3.csv python tools/bench_mask_update.py --variant 3 --iters 2 --warmup 1 > /dev/null 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/mu_launches_v3.csv')))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
ki, vi = rows[h].index('Kernel Name'), rows[h].index('Metric Value')
out = [(r[ki][:44], r[vi]) for r in rows[h + 2:] if len(r) > vi and 'k_pack' not in r[ki] and 'rigl::' in r[ki]]
for k, v in out[-8:]: print(k, v)
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/mu_bench_c2.json 2> gpurun_out/mu_bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/mu_bench_c2.json')); print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['mask_update_ms'], d['roofline']['ms_per_step_by_kind'])" || tail -5 gpurun_out/mu_bench_c2.err
