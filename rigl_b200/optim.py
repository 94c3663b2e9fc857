"""The wrapped ("inner") optimizer of the reference's training scripts on one fused launch.

`FusedMomentumSGD` is tf.train.MomentumOptimizer(learning_rate, momentum, use_nesterov) + the l2 term on the raw
weights, as imagenet_train_eval.py:355-365 / cifar resnet_train_eval.py build it under the sparse wrapper, for
every parameter of a model in ONE kernel (csrc/sgd.cu).  For masked layers the gradient it consumes is
`mask * dense_grad` formed while loading the dense gradient (sparse_optimizers_base.py:478-485), so the masked
gradient tensor is never materialised.  The learning rate lives in device memory: `set_lr` (or assigning
param_groups[...]['lr'] between steps) takes effect in CUDA-graph replays without re-capture.

It is a `torch.optim.Optimizer`: the sparse wrappers only call `.step()`, `.zero_grad()`, `.state` (slot name
'momentum_buffer', as torch.optim.SGD) and `.param_groups`.
"""
import ctypes as C

import torch

from . import _cabi


class FusedMomentumSGD(torch.optim.Optimizer):

  def __init__(self, params, lr=0.1, momentum=0.9, nesterov=True, weight_decay=0.0):
    if momentum < 0 or lr < 0 or weight_decay < 0:
      raise ValueError('lr, momentum and weight_decay must be non-negative')
    super(FusedMomentumSGD, self).__init__(params, dict(lr=lr, momentum=momentum, nesterov=nesterov,
                                                        weight_decay=weight_decay))
    if len(self.param_groups) != 1:
      raise ValueError('FusedMomentumSGD supports a single parameter group')
    self._masked = {}            # id(weight) -> masked layer (dense gradient + bitmap replace weight.grad)
    self._grad_scale = 1.0
    self._plan, self._key = C.c_void_p(None), None
    self._lr_dev, self._lr_uploaded = None, None

  def __del__(self):
    try:
      self._destroy()
    except Exception:
      pass

  def _destroy(self):
    if self._plan and self._plan.value:
      _cabi.lib().rigl_sgd_plan_destroy(self._plan)
      self._plan = C.c_void_p(None)

  # ---- masked layers: consume mask * dense_grad (* grad_scale) instead of weight.grad
  def attach_masked_layers(self, layers, grad_scale=1.0, other_grad_scale=1.0):
    """grad_scale multiplies the masked layers' dense gradients, other_grad_scale every other gradient
    (1 / replicas when the buffers hold cross-replica SUMS)."""
    self._masked = {id(l.weight): l for l in layers}
    self._grad_scale = float(grad_scale)
    self._other_scale = float(other_grad_scale)
    self._key = None
    return self

  # ---- learning rate in device memory
  def set_lr(self, lr):
    """Sets the learning rate (eagerly: call it OUTSIDE graph capture / between replays)."""
    lr = float(lr)
    self.param_groups[0]['lr'] = lr
    self._upload_lr()

  def _upload_lr(self):
    g = self.param_groups[0]
    dev = g['params'][0].device
    if self._lr_dev is None or self._lr_dev.device != dev:
      self._lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
      self._lr_uploaded = None
    lr = float(g['lr'])
    if self._lr_uploaded != lr:
      self._lr_dev.fill_(lr)
      self._lr_uploaded = lr

  @torch.no_grad()
  def prepare(self):
    """Creates the momentum slots, the device learning rate and the launch plan for the CURRENT gradient
    buffers.  Allocates, so it cannot run under stream capture: call it once before capturing `step()`
    (TrainHarness.enable_cuda_graph does)."""
    g = self.param_groups[0]
    self._upload_lr()
    ents = []
    for p in g['params']:
      layer = self._masked.get(id(p))
      if layer is not None:
        grad, bits, scale = layer.masked_weights.dense_grad, layer.mask.bits, self._grad_scale
      else:
        if p.grad is None:
          continue
        grad, bits, scale = p.grad, None, getattr(self, '_other_scale', 1.0)
      if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or grad.dtype != torch.float32 \
          or not grad.is_contiguous():
        raise ValueError('FusedMomentumSGD needs contiguous float32 CUDA parameters and gradients')
      st = self.state[p]
      if 'momentum_buffer' not in st:
        st['momentum_buffer'] = torch.zeros_like(p, memory_format=torch.preserve_format)
      ents.append((p.data_ptr(), st['momentum_buffer'].data_ptr(), grad.data_ptr(),
                   0 if bits is None else bits.data_ptr(), p.numel(), float(g['weight_decay']), float(scale)))
    key = tuple(ents)
    if key != self._key:
      self._destroy()
      if ents:
        descs = (_cabi.SgdDesc * len(ents))()
        for d, (pp, mp, gp, bp, n, wd, sc) in zip(descs, ents):
          d.param, d.momentum, d.grad, d.mask_bits, d.n, d.weight_decay, d.grad_scale = pp, mp, gp, bp or None, n, wd, sc
        plan = C.c_void_p(None)
        _cabi.check(_cabi.lib().rigl_sgd_plan_create(descs, len(ents), C.byref(plan)), 'rigl_sgd_plan_create')
        self._plan = plan
      self._key = key

  def _current_key_matches(self):
    """Cheap check under capture: the gradient buffers are the ones the plan was built for."""
    if self._key is None:
      return False
    i = 0
    for p in self.param_groups[0]['params']:
      layer = self._masked.get(id(p))
      grad = layer.masked_weights.dense_grad if layer is not None else p.grad
      if grad is None:
        continue
      if i >= len(self._key) or self._key[i][0] != p.data_ptr() or self._key[i][2] != grad.data_ptr():
        return False
      i += 1
    return i == len(self._key)

  @torch.no_grad()
  def step(self, closure=None):
    loss = None
    if closure is not None:
      with torch.enable_grad():
        loss = closure()
    g = self.param_groups[0]
    if torch.cuda.is_current_stream_capturing():
      if not self._current_key_matches():
        raise RuntimeError('FusedMomentumSGD.step() under stream capture needs prepare() with the same gradient '
                           'buffers first')
    else:
      self.prepare()
    if self._plan and self._plan.value:
      _cabi.check(_cabi.lib().rigl_sgd_plan_run(self._plan, self._lr_dev.data_ptr(), float(g['momentum']),
                                                int(bool(g['nesterov'])), _cabi.stream_ptr()), 'rigl_sgd_plan_run')
    return loss


def imagenet_lr_schedule(current_epoch, base_learning_rate=0.1, train_batch_size=4096, architecture='resnet',
                         training_steps_multiplier=1.0):
  """lr_schedule of imagenet_train_eval.py:317-330 (step schedule, no SGDR) with set_lr_schedule :280-299:
  (multiplier, start epoch) pairs, linear ramp from 0 to the first multiplier over the first start epoch."""
  if architecture in ('mobilenet_v1', 'mobilenet_v2'):
    sched = [(1.0, 8), (0.1, 40), (0.01, 75), (0.001, 95), (.0003, 120)]
  elif architecture == 'resnet' or architecture.startswith('vgg'):
    sched = [(1.0, 0), (0.1, 30), (0.01, 70), (0.001, 90), (.0001, 120)]
  else:
    raise ValueError('Unknown architecture ' + architecture)
  if training_steps_multiplier != 1.0:
    sched = [(x, y * training_steps_multiplier) for x, y in sched]
  scaled_lr = base_learning_rate * (train_batch_size / 256.0)
  # the ramp term is only selected while current_epoch < first start epoch (tf.where), so a zero-length ramp
  # (ResNet) is never read
  rate = scaled_lr * sched[0][0] * current_epoch / sched[0][1] if sched[0][1] > 0 else scaled_lr * sched[0][0]
  for mult, start_epoch in sched:
    if not current_epoch < start_epoch:
      rate = scaled_lr * mult
  return rate


def make_imagenet_lr_fn(base_learning_rate=0.1, train_batch_size=4096, num_train_images=1281167,
                        architecture='resnet', training_steps_multiplier=1.0):
  """global_step -> learning rate, as train_function computes it (imagenet_train_eval.py:350-354)."""
  steps_per_epoch = num_train_images / float(train_batch_size)

  def lr_fn(global_step):
    return imagenet_lr_schedule(float(int(global_step)) / steps_per_epoch, base_learning_rate, train_batch_size,
                                architecture, training_steps_multiplier)
  return lr_fn
