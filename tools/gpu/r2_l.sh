#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_whole_step_parity_gpu.py tests/test_train_step_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "depthwise or mobilenet" 2>&1 | tail -15
for v in 1 0; do
RIGL_NATIVE_DEPTHWISE=$v timeout 600 python bench.py --config c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_c4_dw$v.json 2> gpurun_out/r2l_bench_c4_dw$v.err; echo "bench c4 native=$v exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_c4_dw$v.json')); print(d['metric'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_step_by_kind'], d['roofline']['frac_step'])" || tail -5 gpurun_out/r2l_bench_c4_dw$v.err
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2l_launches_c4.csv python tools/step_for_ncu.py --config c4 --steps 1 --warmup 2 > gpurun_out/r2l_step_c4.log 2>&1; echo "ncu c4 exit $?"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_c2.json 2> gpurun_out/r2l_bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/r2l_bench_c2.json')); print('c2', d['value'], d['ms_per_step'], d['e2e']['value'], d['mask_update_ms'], d['roofline']['ms_per_step_by_kind'])"
