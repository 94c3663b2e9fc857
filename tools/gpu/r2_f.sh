#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_whole_step_parity_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2f_whole.log 2>&1; echo "whole exit $?"; tail -25 gpurun_out/r2f_whole.log
python - <<'PY'
import json, statistics
for m in ('resnet50','wrn22_2','mobilenet_v1'):
  try:
    d=json.load(open('gpurun_out/whole_step_parity_%s.json'%m))
  except Exception as e:
    print(m, e); continue
  v=list(d['rel_l2'].values())
  print(m, 'loss', d['loss_cuda'], d['loss_oracle'], 'free-running rel first %.4f median %.4f max %.4f last %.4f'%(v[0], statistics.median(v), max(v), v[-1]))
  t=list(d['teacher_forced'].values())
  print('   teacher-forced max: fprop %.2e dgrad %.2e dense wgrad %.2e'%(max(x[0] for x in t), max(x[1] for x in t), max(x[2] for x in t)))
PY
# ncu launch list of one C2 step (shares + DRAM bytes)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/r2f_launches.csv python tools/step_for_ncu.py --steps 1 --warmup 2 \
  > gpurun_out/r2f_step.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r2f_launches.csv
timeout 300 python tools/bench_mask_update.py > gpurun_out/r2f_mask.json 2> gpurun_out/r2f_mask.err; tail -3 gpurun_out/r2f_mask.json
timeout 300 python tools/bench_mask_update.py --inkernel-noise >> gpurun_out/r2f_mask.json 2>> gpurun_out/r2f_mask.err; tail -1 gpurun_out/r2f_mask.json
timeout 300 python tools/bench_mask_update.py --noise >> gpurun_out/r2f_mask.json 2>> gpurun_out/r2f_mask.err; tail -1 gpurun_out/r2f_mask.json
