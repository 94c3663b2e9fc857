#!/bin/bash
# NG=<n> CFG=<c2|c3|c4|c5>: one data-parallel bench line of a BASELINE config on n GPUs (default gradient exchange)
# usage: tools/gpu/retry_n.sh <n> <log> <timeout> "NG=<n> CFG=c3 bash tools/gpu/multi_gpu_config.sh"
mkdir -p gpurun_out
N=${NG:-8}; C=${CFG:-c3}; S=${STEPS:-20}
out=gpurun_out/dp_${C}_n$N
s=$(date +%s)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $C --steps $S --warmup 3 --no-cpu-baseline > $out.json 2> $out.err; echo "exit $? wall $(( $(date +%s) - s )) s"
python -c "
import json; d=json.load(open('$out.json')); print('$C', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('masks_identical_across_replicas'), d['config'].get('mask_update_steps'), d['config'].get('global_batch'))" || tail -15 $out.err
