"""Runs a few ResNet-50 train steps (for ncu launch lists / captures).
  python tools/step_for_ncu.py [--steps 2] [--warmup 2] [--batch 256]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rigl_b200 import workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=2)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--batch', type=int, default=256)
args = ap.parse_args()
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = workloads.ResNet50(device=dev)
workloads.init_masks(model, 'erdos_renyi_kernel', 0.8, seed=0)
h = workloads.TrainHarness(model, lr=0.1)
x = torch.randn(args.batch, 3, 224, 224, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
y = torch.randint(0, 1000, (args.batch,), device=dev)
for _ in range(args.warmup):
  h.step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
  h.step(x, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
