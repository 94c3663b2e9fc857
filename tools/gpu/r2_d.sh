#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/debug_whole_step.py mobilenet_v1 > gpurun_out/r2d_dbg_mbv1.log 2>&1; tail -32 gpurun_out/r2d_dbg_mbv1.log
timeout 200 python tools/debug_whole_step.py resnet50 > gpurun_out/r2d_dbg_r50.log 2>&1; tail -60 gpurun_out/r2d_dbg_r50.log
RIGL_FUSE_BN_STATS=0 timeout 200 python tools/debug_whole_step.py resnet50 > gpurun_out/r2d_dbg_r50_nostats.log 2>&1; tail -12 gpurun_out/r2d_dbg_r50_nostats.log
timeout 200 python tools/debug_whole_step.py resnet50 32 128 > gpurun_out/r2d_dbg_r50_big.log 2>&1; tail -12 gpurun_out/r2d_dbg_r50_big.log
for d in 0 1 2 3; do
  STATS_DBG=$d timeout 200 python tools/bench_conv_layer.py --shapes stats --iters 10 --tag dbg$d 2>/dev/null | grep -E '"fprop' | python -c "
import sys, json
for l in sys.stdin:
  d = json.loads(l); print(d['tag'], d['shape'], d['op'], d['us'])"
done
