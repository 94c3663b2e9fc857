#!/bin/bash
# the batched mask update: bit-exactness tests on every kernel variant (default and a small scan-block size), then
# the micro-benchmark over variant x scan-block size (CUDA events) and per-kernel launch lists
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_mask_update_gpu.py tests/test_optimizers_gpu.py -q -m gpu -p no:cacheprovider -x --tb=short ) > gpurun_out/mu_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/mu_pytest.log
( time RIGL_MASK_CHUNK=8192 timeout 900 python -m pytest tests/test_mask_update_gpu.py -q -m gpu -p no:cacheprovider -x --tb=short ) > gpurun_out/mu_pytest_8k.log 2>&1; echo "pytest(8192) exit $?"; tail -4 gpurun_out/mu_pytest_8k.log
for ch in 32768 16384 8192 4096; do for v in 2 3; do
  RIGL_MASK_CHUNK=$ch timeout 300 python tools/bench_mask_update.py --variant $v > gpurun_out/mu_bench_v${v}_$ch.json 2> gpurun_out/mu_bench_v${v}_$ch.err
  python -c "
import json; d=json.load(open('gpurun_out/mu_bench_v${v}_$ch.json')); print('variant $v chunk $ch', round(d['ms_median'],4), round(d['ms_min'],4))" || tail -3 gpurun_out/mu_bench_v${v}_$ch.err
done; done
RIGL_MASK_CHUNK=8192 timeout 300 python tools/bench_mask_update.py --variant 2 --inkernel-noise | python -c "import json,sys; d=json.load(sys.stdin); print('v2 8192 noise', d['ms_median'])"
for v in 2 3; do
RIGL_MASK_CHUNK=8192 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mu_launches_v${v}_8k.csv python tools/bench_mask_update.py --variant $v --iters 2 --warmup 1 > /dev/null 2>&1; echo "ncu exit $?"
python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/mu_launches_v${v}_8k.csv')))
h = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
ki, vi = rows[h].index('Kernel Name'), rows[h].index('Metric Value')
out = [(r[ki][:44], r[vi]) for r in rows[h + 2:] if len(r) > vi and 'k_pack' not in r[ki] and 'rigl::' in r[ki]]
for k, v in out[-7:]: print(k, v)
PY
done
