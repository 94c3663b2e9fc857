#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/r2j_pytest.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r2j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/r2j_layers.json > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['mask_update_ms'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'], d['roofline']['frac_step'])"
RIGL_BN_RELU_BITS=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_nobits.json 2> gpurun_out/r2j_bench_nobits.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_nobits.json')); print('nobits', d['value'], d['ms_per_step'], d['roofline']['ms_per_step_by_kind'])"
RIGL_WGRAD_FIXUP=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2j_bench_nofix.json 2> gpurun_out/r2j_bench_nofix.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_nofix.json')); print('nofixup', d['value'], d['ms_per_step'], d['roofline']['ms_per_step_by_kind'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2j_launches_c4.csv python tools/step_for_ncu.py --config c4 --steps 1 --warmup 2 > gpurun_out/r2j_step_c4.log 2>&1; echo "ncu c4 exit $?"
