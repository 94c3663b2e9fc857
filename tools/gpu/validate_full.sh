#!/bin/bash
# single-GPU validation of the tree: what the driver runs (GPU tests, smoke, bench both arms), the other BASELINE
# configs, and the ncu launch list of the C2 step (-> profiles/)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/full_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/full_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/full_bench_default.json 2> gpurun_out/full_bench_default.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/full_bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['frac_step'], d['cpu_baseline']['value'], d['cpu_baseline']['spread'], d['clocks'], d['config']['step_ms'], d['mask_update_ms'])"
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/full_bench_reference.json 2> gpurun_out/full_bench_reference.err; echo "ref exit $?"; head -c 600 gpurun_out/full_bench_reference.json; echo
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/full_bench_s20.json 2> gpurun_out/full_bench_s20.err; python -c "
import json; d=json.load(open('gpurun_out/full_bench_s20.json')); print('s20', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['step_ms'], d['config']['mask_update_steps'], d['roofline']['ms_per_step_by_kind'])"
for c in c3 c4 c5; do
timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/full_bench_$c.json 2> gpurun_out/full_bench_$c.err; python -c "
import json; d=json.load(open('gpurun_out/full_bench_$c.json')); print('$c', d['value'], d['ms_per_step'], d['e2e']['value'], d['mask_update_ms'])" || tail -5 gpurun_out/full_bench_$c.err
done
bash tools/gpu/launch_list.sh
