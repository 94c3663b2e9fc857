#!/bin/bash
# NG=2|4|8: the data-parallel bench with the overlapped (RIGL_DP_OVERLAP=1) and the one-all-reduce exchange, back to back
# usage: tools/gpu/retry_n.sh <N> <log> <timeout> "NG=<N> bash tools/gpu/multi_gpu_ab.sh"
mkdir -p gpurun_out
N=${NG:-2}
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/dp_ab_${tag}_n$N.json 2> gpurun_out/dp_ab_${tag}_n$N.err; echo "$tag exit $?"; python -c "
import json; d=json.load(open('gpurun_out/dp_ab_${tag}_n$N.json')); print('$tag', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['masks_identical_across_replicas'], d['config']['cuda_graph'])" || tail -15 gpurun_out/dp_ab_${tag}_n$N.err; }

s=$(date +%s); run overlap RIGL_DP_OVERLAP=1; echo "wall $(( $(date +%s) - s )) s"
s=$(date +%s); run blocking RIGL_DP_OVERLAP=0; echo "wall $(( $(date +%s) - s )) s"
