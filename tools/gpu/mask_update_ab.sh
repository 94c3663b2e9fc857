#!/bin/bash
# the batched mask update: bit-exactness tests (default scan-block size and a large one), the micro-benchmark over
# scan-block sizes (CUDA events) and the per-kernel launch list
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_mask_update_gpu.py tests/test_optimizers_gpu.py tests/test_whole_step_parity_gpu.py -q -m gpu -p no:cacheprovider -x --tb=short ) > gpurun_out/mu_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/mu_pytest.log
( time RIGL_MASK_CHUNK=32768 timeout 900 python -m pytest tests/test_mask_update_gpu.py -q -m gpu -p no:cacheprovider -x --tb=short ) > gpurun_out/mu_pytest_32k.log 2>&1; echo "pytest(32768) exit $?"; tail -4 gpurun_out/mu_pytest_32k.log
for ch in 32768 16384 8192 4096; do
  RIGL_MASK_CHUNK=$ch timeout 300 python tools/bench_mask_update.py > gpurun_out/mu_bench_$ch.json 2> gpurun_out/mu_bench_$ch.err
  RIGL_MASK_CHUNK=$ch timeout 300 python tools/bench_mask_update.py --inkernel-noise > gpurun_out/mu_bench_${ch}_noise.json 2>> gpurun_out/mu_bench_$ch.err
  python -c "
import json; d=json.load(open('gpurun_out/mu_bench_$ch.json')); e=json.load(open('gpurun_out/mu_bench_${ch}_noise.json')); print('scan block $ch', round(d['ms_median'],4), round(d['ms_min'],4), 'noise', round(e['ms_median'],4))" || tail -3 gpurun_out/mu_bench_$ch.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mu_launches.csv python tools/bench_mask_update.py --iters 2 --warmup 1 --inkernel-noise > /dev/null 2>&1; echo "ncu exit $?"
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/mu_launches.csv')))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
ki, vi = rows[h].index('Kernel Name'), rows[h].index('Metric Value')
out = [(r[ki][:44], r[vi]) for r in rows[h + 2:] if len(r) > vi and 'k_pack' not in r[ki] and 'rigl::' in r[ki]]
for k, v in out[-7:]: print(k, v)
PY
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/mu_bench_c2.json 2> gpurun_out/mu_bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/mu_bench_c2.json')); print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['mask_update_ms'], d['roofline']['ms_per_step_by_kind'])" || tail -5 gpurun_out/mu_bench_c2.err
