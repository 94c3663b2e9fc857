#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/f_bench_n2.json 2> gpurun_out/f_bench_n2.err
echo "bench n2 exit $?"; grep -v "^  " gpurun_out/f_bench_n2.err | tail -6; cat gpurun_out/f_bench_n2.json
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 --cpu-batch 4 > gpurun_out/f_bench_ref_n2.json 2> gpurun_out/f_bench_ref_n2.err
echo "ref exit $?"; cat gpurun_out/f_bench_ref_n2.json | cut -c1-500
