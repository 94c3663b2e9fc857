#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=10 > gpurun_out/validate_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/validate_pytest.log; tail -6 gpurun_out/validate_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/validate_smoke.log 2>&1; tail -1 gpurun_out/validate_smoke.log
timeout 900 python bench.py > gpurun_out/validate_bench_default.json 2> gpurun_out/validate_bench_default.err
echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/validate_bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'])"
