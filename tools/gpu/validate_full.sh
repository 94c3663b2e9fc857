#!/bin/bash
# final single-GPU validation: what the driver runs (GPU tests, smoke, bench both arms) + the depthwise A/B
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/full_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/full_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/full_bench_default.json 2> gpurun_out/full_bench_default.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/full_bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['frac_step'], d['cpu_baseline']['value'], d['cpu_baseline']['spread'], d['clocks'], d['config']['step_ms'], d['mask_update_ms'])"
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/full_bench_reference.json 2> gpurun_out/full_bench_reference.err; echo "ref exit $?"; head -c 600 gpurun_out/full_bench_reference.json; echo
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/full_bench_s20.json 2> gpurun_out/full_bench_s20.err; python -c "
import json; d=json.load(open('gpurun_out/full_bench_s20.json')); print('s20', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['step_ms'], d['config']['mask_update_steps'])"
for v in 1 0; do
RIGL_NATIVE_DEPTHWISE=$v timeout 600 python bench.py --config c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/full_bench_c4_dw$v.json 2> gpurun_out/full_bench_c4_dw$v.err; python -c "
import json; d=json.load(open('gpurun_out/full_bench_c4_dw$v.json')); print('c4 native=$v', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/full_bench_c4_dw$v.err
done
RIGL_NATIVE_DEPTHWISE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/full_launches_c4.csv python tools/step_for_ncu.py --config c4 --steps 1 --warmup 2 > /dev/null 2>&1; grep -c depthwise gpurun_out/full_launches_c4.csv
timeout 300 python bench.py --config c5 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/full_bench_c5.json 2> gpurun_out/full_bench_c5.err; python -c "
import json; d=json.load(open('gpurun_out/full_bench_c5.json')); print('c5', d['value'], d['ms_per_step'], d['e2e']['value'])"
