#!/bin/bash
mkdir -p gpurun_out
for v in "RIGL_FUSE_BN_STATS=0" "RIGL_PACK_AHEAD=0" "RIGL_STEM_S2D=0" "RIGL_FUSED_SGD=0"; do
  echo "=== $v"
  env $v timeout 300 python -m pytest tests/test_whole_step_parity_gpu.py -m gpu -q --tb=line -p no:cacheprovider 2>&1 | grep -E "rel L2|passed|failed|Error" | head -8
  cp gpurun_out/whole_step_parity_resnet50.json gpurun_out/r2c_whole_r50_${v%%=*}.json 2>/dev/null
done
timeout 600 python -m pytest tests/test_bn_gpu.py tests/test_conv_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "epilogue or batched_pack" 2>&1 | tail -15
bash tools/gpu/r2_d.sh
