#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_whole_step_parity_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2h_whole.log 2>&1; echo "whole exit $?"; grep -E "^E  |passed|failed" gpurun_out/r2h_whole.log | head -20
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
timeout 300 python -m pytest tests/test_mask_update_gpu.py tests/test_optimizers_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5
for a in "" "--inkernel-noise"; do timeout 300 python tools/bench_mask_update.py $a | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['inkernel_noise'], d['ms_median'])"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_mask_launches.csv python tools/bench_mask_update.py --iters 2 --warmup 1 --inkernel-noise > /dev/null 2>&1; python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r2h_mask_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
agg=collections.Counter(); cnt=collections.Counter()
for r in rows[hdr+1:]:
  if len(r)<10: continue
  name=r[4].split('(')[0][:40]; 
  try: v=float(r[-1].replace(',',''))
  except: continue
  agg[name]+=v; cnt[name]+=1
for k,v in agg.most_common(12): print('%-42s %3d launches %9.1f us total  %.1f us each'%(k,cnt[k],v/1e3 if v>1e4 else v, (v/cnt[k])/1e3 if v>1e4 else v/cnt[k]))
PY
