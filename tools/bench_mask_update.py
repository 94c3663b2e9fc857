"""Microbenchmark of one batched RigL mask update on the ResNet-50 layer set
(54 masked tensors, 25.5 M weights, ERK sparsities from the golden fixture).

  python tools/bench_mask_update.py [--sparsity 0.8] [--iters 20] [--noise]
Timing: CUDA events on the launching stream, L2 flushed between iterations."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigl_b200 import _cabi  # noqa: E402
from rigl_b200.masks import MaskUpdateEngine, MaskVariable  # noqa: E402


def build_layers(tag, noise, dev='cuda:0', seed=0):
  with open(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'sparse_utils_golden.json')) as f:
    case = [c for c in json.load(f)['cases'] if c['tag'] == tag][0]
  g = torch.Generator(device=dev).manual_seed(seed)
  specs = []
  for name, shape in case['layers']:
    n = int(np.prod(shape))
    s = float.fromhex(case['sparsities_hex'][name + '/mask:0'])
    mv = MaskVariable(name, shape, dev)
    keep = torch.rand(n, device=dev, generator=g) >= s
    mv.assign(keep.float().view(shape))
    spec = dict(mask=mv, weights=torch.randn(n, device=dev, generator=g) * 0.05,
                score_grow=torch.randn(n, device=dev, generator=g) * 1e-3,
                slots=[torch.randn(n, device=dev, generator=g)])
    if noise:
      spec['noise'] = torch.randn(n, device=dev, generator=g) * 1e-5
    specs.append(spec)
  return specs


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--tag', default='r50_erk80')
  ap.add_argument('--iters', type=int, default=20)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--noise', action='store_true', help='drop-score noise read from a tensor')
  ap.add_argument('--inkernel-noise', action='store_true', help='drop-score noise drawn inside the kernels')
  args = ap.parse_args()
  specs = build_layers(args.tag, args.noise)
  total_n = sum(s['mask'].size for s in specs)
  eng = MaskUpdateEngine()
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda:0')
  times = []
  for it in range(args.warmup + args.iters):
    flush.zero_()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    if args.inkernel_noise:
      eng.run(specs, np.float32(0.3), noise_std=1e-5, noise_seed=(77 << 32) | it)
    else:
      eng.run(specs, np.float32(0.3))
    stop.record()
    torch.cuda.synchronize()
    if it >= args.warmup:
      times.append(start.elapsed_time(stop))
  stats = eng.stats()
  ms = float(np.median(times))
  alg_bytes = (8.25 + (4 if args.noise else 0)) * total_n
  print(json.dumps({'bench': 'mask_update', 'tag': args.tag, 'layers': len(specs), 'weights': total_n,
                    'scan_block': int(os.environ.get('RIGL_MASK_CHUNK', 8192)), 'noise': args.noise, 'inkernel_noise': args.inkernel_noise, 'ms_median': ms, 'ms_min': float(min(times)),
                    'algorithmic_GBps': alg_bytes / ms / 1e6, 'workspace_MB': eng.workspace_bytes / 2 ** 20,
                    'max_drop_candidates': max(s[3] for s in stats),
                    'max_grow_candidates': max(s[4] for s in stats),
                    'launches_total': _cabi.launch_count()}))


if __name__ == '__main__':
  main()
