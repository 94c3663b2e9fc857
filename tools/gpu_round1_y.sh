#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=10 > gpurun_out/y_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/y_pytest.log; tail -4 gpurun_out/y_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/y_smoke.log 2>&1; tail -2 gpurun_out/y_smoke.log
timeout 900 python bench.py > gpurun_out/y_bench_default.json 2> gpurun_out/y_bench_default.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/y_bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac'], d['cpu_baseline'], d['clocks'], d['mask_update_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/y_launches.csv python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/y_step.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/y_launches.csv
