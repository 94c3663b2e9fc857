"""CPU restatement of the reference's sparse ResNet-50 train step -- TEST / BASELINE
INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/rigl_oracle.py header).

What the reference executes per step on its CPU path (TF1 graph):
  masked_weights = mask * weights (dense fp32 tensor, every step)
  fwd: conv(x, masked_weights) for 53 convs + BN/ReLU/pool + masked FC
  bwd: TWO gradient sets -- wrt `weights` (masked) and wrt `masked_weights` (DENSE,
       sparse_optimizers_base.py:478-485; TF1 cond semantics compute it every step)
  then either the Nesterov-momentum step or, on update iterations, the drop/grow
  update (two full sorts per layer, base.py:276-343).
TensorFlow is not installable here, so this is a torch-CPU port of that math
(fp32, dense-executed like TF's Eigen kernels), used as `cpu_baseline` (kind
"port") and by `bench.py --impl reference`.  Parity: unpinned against TF itself.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import rigl_oracle as orc


class CpuResNet50(object):

  def __init__(self, sparsity=0.8, seed=0, num_classes=1000):
    rng = np.random.RandomState(seed)
    self.layers = orc.resnet50_masked_layers()
    masks = [orc.FakeMask(n + '/mask:0', sh) for n, sh, _, _ in self.layers]
    sp = orc.get_sparsities(masks, 'erdos_renyi_kernel', sparsity, {})
    self.w, self.m, self.mom = {}, {}, {}
    for n, sh, _, _ in self.layers:
      fan_in = int(np.prod(sh[:-1]))
      self.w[n] = torch.from_numpy((rng.standard_normal(sh) * np.sqrt(2.0 / fan_in)).astype(np.float32))
      self.m[n] = torch.from_numpy(orc.get_mask_random_numpy(sh, sp[n + '/mask:0'], rng).astype(np.float32))
      self.mom[n] = torch.zeros(sh)
    self.bn = {}
    self.fc_bias = torch.zeros(num_classes)

  def _bn(self, x, key, relu=True):
    c = x.shape[1]
    if key not in self.bn:
      self.bn[key] = (torch.ones(c, requires_grad=True), torch.zeros(c, requires_grad=True))
    g, b = self.bn[key]
    x = F.batch_norm(x, None, None, g, b, training=True, momentum=0.1, eps=1e-5)
    return F.relu(x) if relu else x

  def _conv(self, x, name, masked, stride):
    w = masked[name]                       # HWIO
    k = w.shape[0]
    return F.conv2d(x, w.permute(3, 2, 0, 1), stride=stride, padding=(k - 1) // 2)

  def forward_backward(self, images, labels):
    """Returns (loss, dense grads dict).  images [N,3,H,W] fp32."""
    masked = {}
    for n, _, _, _ in self.layers:
      masked[n] = (self.m[n] * self.w[n]).requires_grad_(True)      # materialised every step
    p = 'resnet_model/'
    x = self._bn(self._conv(images, p + 'initial_conv', masked, 2), 'bn0')
    x = F.max_pool2d(x, 3, 2, 1)
    for g, (n_blocks, stride) in enumerate(((3, 1), (4, 2), (6, 2), (3, 2)), 1):
      for b in range(n_blocks):
        sfx = ('block_group_projection_block_group%d' % g) if b == 0 else ('block_group%d_%d_1' % (g, b))
        s = stride if b == 0 else 1
        sc = x
        if b == 0:
          sc = self._bn(self._conv(x, p + 'bottleneck_projection_' + sfx, masked, s), sfx + 'p', relu=False)
        y = self._bn(self._conv(x, p + 'bottleneck_1_' + sfx, masked, 1), sfx + '1')
        y = self._bn(self._conv(y, p + 'bottleneck_2_' + sfx, masked, s), sfx + '2')
        y = self._bn(self._conv(y, p + 'bottleneck_3_' + sfx, masked, 1), sfx + '3', relu=False)
        x = F.relu(y + sc)
    x = x.mean(dim=(2, 3))
    logits = x @ masked[p + 'final_dense'] + self.fc_bias
    loss = F.cross_entropy(logits, labels, label_smoothing=0.1)
    loss.backward()
    dense = {n: masked[n].grad for n, _, _, _ in self.layers}
    return float(loss.detach()), dense

  def optimizer_step(self, dense, lr=0.1, momentum=0.9, wd=1e-4):
    for n, _, _, _ in self.layers:
      g = self.m[n] * dense[n] + wd * self.w[n]              # dL/dweights + l2 on raw weights
      self.mom[n].mul_(momentum).add_(g)
      self.w[n].sub_(lr * (g + momentum * self.mom[n]))

  def mask_update(self, dense, drop_fraction=0.3):
    for n, _, _, _ in self.layers:
      r = orc.rigl_mask_update(self.m[n].numpy(), self.w[n].numpy(), dense[n].numpy(), drop_fraction,
                               slots=[self.mom[n].numpy()])
      self.m[n] = torch.from_numpy(r['mask'])
      self.w[n] = torch.from_numpy(r['weights'])
      self.mom[n] = torch.from_numpy(r['slots'][0])


def time_train_steps(batch, steps, warmup=1, image_hw=224, sparsity=0.8, seed=0):
  """Seconds per train step (median) of the CPU port at `batch` images."""
  torch.manual_seed(seed)
  net = CpuResNet50(sparsity=sparsity, seed=seed)
  images = torch.randn(batch, 3, image_hw, image_hw)
  labels = torch.randint(0, 1000, (batch,))
  times = []
  dense = None
  for i in range(warmup + steps):
    t0 = time.perf_counter()
    _, dense = net.forward_backward(images, labels)
    net.optimizer_step(dense)
    dt = time.perf_counter() - t0
    if i >= warmup:
      times.append(dt)
  return float(np.median(times)), net, dense


def time_mask_update(net, dense, repeats=1):
  times = []
  for _ in range(repeats):
    t0 = time.perf_counter()
    net.mask_update(dense, 0.3)
    times.append(time.perf_counter() - t0)
  return float(np.median(times))
