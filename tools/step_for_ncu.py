"""Runs a few train steps of a BASELINE config eagerly (for ncu launch lists / captures).
  python tools/step_for_ncu.py [--config c2|c3|c4|c5] [--steps 2] [--warmup 2] [--batch N]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rigl_b200 import workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='c2')
ap.add_argument('--steps', type=int, default=2)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--batch', type=int, default=0)
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
batch = args.batch or cfg['batch']
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = bench.build_model(cfg, dev)
h = workloads.TrainHarness(model, lr=0.1, weight_decay=5e-4 if cfg['model'] == 'wrn22_2' else 1e-4,
                           label_smoothing=0.0 if cfg['model'] == 'wrn22_2' else 0.1)
x = torch.randn(batch, 3, cfg['image'], cfg['image'], device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
y = torch.randint(0, cfg['classes'], (batch,), device=dev)
for _ in range(args.warmup):
  h.step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
  h.step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
