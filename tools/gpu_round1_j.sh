#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_step_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/j_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/j_pytest.log; tail -6 gpurun_out/j_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/j_bench_graph.json 2> gpurun_out/j_bench_graph.err
echo "bench(graph) exit $?"; grep -i "warn\|error" gpurun_out/j_bench_graph.err | tail -5; cut -c1-900 gpurun_out/j_bench_graph.json
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/j_bench_eager.json 2> gpurun_out/j_bench_eager.err
echo "bench(eager) exit $?"; cut -c1-330 gpurun_out/j_bench_eager.json
# DRAM traffic of every conv-family / BN launch in one step (2 metrics => cheap)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
  -k regex:"k_igemm|k_im2col|k_splitk|k_bn_|k_pack|k_maxpool" --csv --log-file gpurun_out/j_traffic_step.csv python tools/step_for_ncu.py --steps 1 --warmup 2 > gpurun_out/j_ncu.log 2>&1
tail -2 gpurun_out/j_ncu.log; wc -l gpurun_out/j_traffic_step.csv
