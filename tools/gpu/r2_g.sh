#!/bin/bash
# 2-GPU data-parallel check: overlapped bucketed all-reduce (default) vs one blocking all-reduce, N=1 for reference
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_$tag.json 2> gpurun_out/r2g_$tag.err; echo "$tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2g_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['e2e']['value'], d['config'])" || tail -15 gpurun_out/r2g_$tag.err; }
run overlap RIGL_DP_OVERLAP=1
run blocking RIGL_DP_OVERLAP=0
run overlap_nowgradfork RIGL_WGRAD_OVERLAP=0
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2g_n1.json 2> gpurun_out/r2g_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_n1.json')); print('n1', d['value'], d['ms_per_step'])"
