#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/debug_tc.py --big > gpurun_out/i_debug_tc.log 2>&1; grep -E "BAD|EXC|DEBUG_TC|fatal" gpurun_out/i_debug_tc.log
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --maxfail=15 > gpurun_out/i_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/i_pytest.log
tail -4 gpurun_out/i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layer-report gpurun_out/i_layers.json > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench exit $?"; tail -5 gpurun_out/i_bench.err; cat gpurun_out/i_bench.json
RIGL_CLUSTER_MC=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/i_bench_mc.json 2> gpurun_out/i_bench_mc.err
echo "bench (multicast) exit $?"; cut -c1-330 gpurun_out/i_bench_mc.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/i_launches_step.csv python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/i_ncu.log 2>&1
# small ncu --set full captures: 10 launches of each kernel family at batch 64, converted to CSV on the box
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:k_igemm_kmajor -s 20 -c 10 \
  -o gpurun_out/i_kmajor_full python tools/step_for_ncu.py --steps 1 --warmup 2 --batch 128 > gpurun_out/i_ncu_f1.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:k_igemm_wgrad -s 20 -c 10 \
  -o gpurun_out/i_wgrad_full python tools/step_for_ncu.py --steps 1 --warmup 2 --batch 128 > gpurun_out/i_ncu_f2.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:k_bn_ -s 40 -c 12 \
  -o gpurun_out/i_bn_full python tools/step_for_ncu.py --steps 1 --warmup 2 --batch 128 > gpurun_out/i_ncu_f3.log 2>&1
for f in i_kmajor_full i_wgrad_full i_bn_full; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.csv 2>/dev/null
  ls -la gpurun_out/$f.ncu-rep gpurun_out/$f.csv
  sz=$(stat -c %s gpurun_out/$f.ncu-rep); if [ "$sz" -gt 12000000 ]; then rm gpurun_out/$f.ncu-rep; fi
done
du -sh gpurun_out
