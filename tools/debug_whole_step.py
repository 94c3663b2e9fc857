"""Layer-by-layer forward comparison of a CUDA model with the CPU oracle net (development tool): relative L2
error of every batch-norm output in execution order -> the first layer that deviates.

  python tools/debug_whole_step.py [resnet50|mobilenet_v1|wrn22_2] [batch] [image]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from oracle import cpu_train_step as cpu  # noqa: E402
from rigl_b200 import workloads  # noqa: E402
from rigl_b200.norm import FusedBatchNormReLU  # noqa: E402
import test_whole_step_parity_gpu as T  # noqa: E402

DEV = 'cuda:0'


def main():
  which = sys.argv[1] if len(sys.argv) > 1 else 'resnet50'
  batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
  image = int(sys.argv[3]) if len(sys.argv) > 3 else 64
  torch.manual_seed(0)
  if which == 'resnet50':
    net, model, classes = cpu.CpuResNet50(0.8, 11, bf16_weights=True), workloads.ResNet50(device=DEV), 1000
  elif which == 'mobilenet_v1':
    net, model, classes = cpu.CpuMobileNetV1(0.9, 13, bf16_weights=True), workloads.MobileNetV1(device=DEV), 1000
  else:
    net, model, classes = cpu.CpuWideResNet(22, 2, 0.95, 12, bf16_weights=True), \
        workloads.WideResNet(22, 2, droprate=0.0, device=DEV), 10
  net.bn_init = T._bn_init(11)
  net.bf16_act = os.environ.get('ORACLE_FP32') != '1'       # default: the bf16-storage oracle
  net.trace = []
  images = torch.randn(batch, 3, image, image).to(torch.bfloat16)
  labels = torch.randint(0, classes, (batch,))
  net.forward_backward(images.float(), labels)
  T._load(model, net)
  with torch.no_grad():
    if which == 'wrn22_2':
      model.conv_1.weight.copy_(net.p['conv_1'].detach().permute(3, 2, 0, 1).to(DEV))
    if which == 'mobilenet_v1':
      model.initial_conv.weight.copy_(net.p['initial_conv'].detach().permute(3, 2, 0, 1).to(DEV))
      for i, blk in enumerate(model.blocks):
        blk.depthwise.weight.copy_(net.p['depthwise_%d' % i].detach().to(DEV))
  got = []
  bns = [m for m in model.modules() if isinstance(m, FusedBatchNormReLU)]
  for m in bns:
    m.register_forward_hook(lambda mod, inp, out: got.append(((out[0] if isinstance(out, tuple) else out).detach().float().cpu(),
                                                              inp[0].detach().float().cpu())))
  model.train()
  xd = images.to(DEV).contiguous(memory_format=torch.channels_last)
  model(xd)
  torch.cuda.synchronize()
  print('%-4s %-44s %-22s %10s' % ('#', 'oracle key', 'shape', 'rel L2 of the BN output'))
  for i, ((key, want), (g, g_in)) in enumerate(zip(net.trace, got)):
    rel = float((g - want).norm() / (want.norm() + 1e-30))
    print('%-4d %-44s %-22s %10.4f   in: mean %.4f std %.4f' % (i, key, tuple(want.shape), rel, float(g_in.mean()),
                                                               float(g_in.std())))


if __name__ == '__main__':
  main()
