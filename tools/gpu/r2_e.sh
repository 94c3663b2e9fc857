#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_whole_step_parity_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2e_whole.log 2>&1; echo "whole exit $?"; tail -12 gpurun_out/r2e_whole.log
python - <<'PY'
import json
for m in ('resnet50','wrn22_2','mobilenet_v1'):
  try:
    d=json.load(open('gpurun_out/whole_step_parity_%s.json'%m))
  except Exception as e:
    print(m, e); continue
  v=list(d['rel_l2'].values())
  import statistics
  print(m, 'loss', d['loss_cuda'], d['loss_oracle'], 'rel first %.4f median %.4f max %.4f last %.4f'%(v[0], statistics.median(v), max(v), v[-1]))
PY
timeout 200 python tools/debug_whole_step.py resnet50 > gpurun_out/r2e_dbg_r50.log 2>&1; tail -8 gpurun_out/r2e_dbg_r50.log
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect tests/test_whole_step_parity_gpu.py ) > gpurun_out/r2e_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2e_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/r2e_layers.json > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['mask_update_ms'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'], d['roofline']['frac_step'])"
for c in c3 c4 c5; do
timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2e_bench_$c.json 2> gpurun_out/r2e_bench_$c.err; echo "bench $c exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_$c.json')); print(d['metric'], d['value'], d['ms_per_step'], d['e2e']['value'], d['mask_update_ms'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac_step'])" || tail -5 gpurun_out/r2e_bench_$c.err
done
