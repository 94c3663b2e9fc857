#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import torch, time
x = torch.empty(77072384 // 2, dtype=torch.bfloat16).pin_memory()
torch.cuda.synchronize()
for _ in range(2):
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record(); 
  for _ in range(10): y = x.to('cuda', non_blocking=True)
  e.record(); torch.cuda.synchronize()
  print('H2D pinned GB/s', 10 * 77072384 / (s.elapsed_time(e) * 1e-3) / 1e9)
PY
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_new$i.json 2> gpurun_out/r2n_bench_new$i.err; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_new$i.json')); print('new', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
git_old=tools/gpu/bench_prev.py
timeout 600 python $git_old --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_old.json 2> gpurun_out/r2n_bench_old.err; python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_old.json')); print('old', d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -5 gpurun_out/r2n_bench_old.err
bash tools/gpu/ncu_full.sh
