"""Fused batch-norm + ReLU (+ residual add) on NHWC bf16 activations.

Host mirror of `batch_norm_relu(inputs, is_training, relu, init_zero)`
(rigl/imagenet_resnet/resnet_model.py:41-80; BATCH_NORM_DECAY 0.9, EPSILON 1e-5) and of the
`relu(inputs + shortcut)` tail of the bottleneck block (:501), backed by csrc/bn.cu.
Not a masked op in the reference -- it is the HBM-bound glue between the masked convs
(SURVEY 8f row 1), so it gets streaming kernels instead of tensor cores.
"""
import torch
from torch import nn

from . import _cabi
from .layers import _timed, _workspace


def _p(t):
  return None if t is None else t.data_ptr()


import os as _os
RELU_BITMASK = _os.environ.get('RIGL_BN_RELU_BITS', '1') != '0'    # 0: the residual backward re-reads the block output


class _BNFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, y, gamma, beta, residual, mod, partial, fork=False):
    n, c, h, w = y.shape
    rows = n * h * w
    dev = y.device
    out = torch.empty_like(y, memory_format=torch.channels_last)
    save = torch.empty((4, c), dtype=torch.float32, device=dev)      # mean, rstd, scale, shift
    ws = _workspace(dev, _cabi.lib().rigl_bn_workspace_bytes(rows, c) + 8 * c + 256)
    # residual form: the backward needs only the SIGN of the block output -> one bit per element, written by the
    # apply pass (1/16 of re-reading the bf16 tensor in the backward reduce pass)
    bits = torch.empty(rows * c // 8, dtype=torch.uint8, device=dev) if (residual is not None and RELU_BITMASK) else None

    def run():
      if partial is not None:       # statistics already reduced by the producing conv's epilogue
        _cabi.check(_cabi.lib().rigl_bn_forward_train_partials(
            y.data_ptr(), _p(residual), gamma.data_ptr(), beta.data_ptr(), partial[0].data_ptr(), partial[1],
            rows, c, mod.eps, mod.momentum, int(mod.relu), mod.running_mean.data_ptr(),
            mod.running_var.data_ptr(), save[0].data_ptr(), save[1].data_ptr(), save[2].data_ptr(),
            save[3].data_ptr(), out.data_ptr(), _p(bits), _cabi.stream_ptr()), 'rigl_bn_forward_train_partials')
        return
      _cabi.check(_cabi.lib().rigl_bn_forward_train(
          y.data_ptr(), _p(residual), gamma.data_ptr(), beta.data_ptr(), rows, c, mod.eps, mod.momentum,
          int(mod.relu), mod.running_mean.data_ptr(), mod.running_var.data_ptr(), save[0].data_ptr(),
          save[1].data_ptr(), save[2].data_ptr(), save[3].data_ptr(), out.data_ptr(), ws.data_ptr(),
          ws.numel(), _p(bits), _cabi.stream_ptr()), 'rigl_bn_forward_train')
    _timed('bn_fwd', mod, run)
    ctx.mod, ctx.has_res, ctx.fork = mod, residual is not None, bool(fork)
    ctx.save_for_backward(y, (bits if bits is not None else out) if residual is not None else None, save)
    ctx.has_bits = bits is not None
    if fork:
      # Two handles on the same activation for its two consumers (next block's first conv and its
      # shortcut): backward then receives their gradients SEPARATELY and the sum is folded into
      # the column-sum pass (rigl_bn_backward2) instead of autograd's elementwise add.
      ctx.set_materialize_grads(False)
      return out, out.detach()
    return out

  @staticmethod
  def backward(ctx, da, da_b=None):
    y, act, save = ctx.saved_tensors
    mod = ctx.mod
    n, c, h, w = y.shape
    rows = n * h * w

    def as_grad(t):
      t = t.contiguous(memory_format=torch.channels_last)
      return t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16)
    if ctx.fork:
      if da is None and da_b is None:
        return None, None, None, None, None, None, None
      if da is None:
        da, da_b = da_b, None
    da = as_grad(da)
    if da_b is not None:
      da_b = as_grad(da_b)
      if not ctx.has_res:           # only the residual form sums in-kernel
        da, da_b = da + da_b, None
    dy = torch.empty_like(y, memory_format=torch.channels_last)
    dres = torch.empty_like(y, memory_format=torch.channels_last) if ctx.has_res else None
    dgb = torch.empty((2, c), dtype=torch.float32, device=y.device)
    ws = _workspace(y.device, _cabi.lib().rigl_bn_workspace_bytes(rows, c) + 8 * c + 256)

    def run():
      _cabi.check(_cabi.lib().rigl_bn_backward2(
          da.data_ptr(), _p(da_b), y.data_ptr(), None if ctx.has_bits else _p(act), save[0].data_ptr(),
          save[1].data_ptr(), save[2].data_ptr(), save[3].data_ptr(), rows, c, int(mod.relu), dy.data_ptr(), _p(dres),
          dgb[0].data_ptr(), dgb[1].data_ptr(), ws.data_ptr(), ws.numel(), _p(act) if ctx.has_bits else None,
          _cabi.stream_ptr()), 'rigl_bn_backward2')
    _timed('bn_bwd', mod, run)
    return dy, dgb[0], dgb[1], dres, None, None, None


class FusedBatchNormReLU(nn.Module):
  """y -> [relu](BN(y) [+ residual]); training mode uses batch statistics."""

  def __init__(self, channels, relu=True, init_zero=False, eps=1e-5, decay=0.9, device='cuda', name=None):
    super(FusedBatchNormReLU, self).__init__()
    if channels % 8:
      raise ValueError('FusedBatchNormReLU needs channels % 8 == 0')
    self.channels, self.relu, self.eps, self.momentum = channels, bool(relu), float(eps), 1.0 - float(decay)
    self.scope = name or 'batch_normalization'
    self.weight = nn.Parameter(torch.zeros(channels, device=device) if init_zero
                               else torch.ones(channels, device=device))
    self.bias = nn.Parameter(torch.zeros(channels, device=device))
    self.register_buffer('running_mean', torch.zeros(channels, device=device))
    self.register_buffer('running_var', torch.ones(channels, device=device))

  def forward(self, y, residual=None, producer=None, fork=False):
    """`producer`: the SparseConv2d whose output `y` is; if its epilogue emitted the batch statistics
    of exactly this tensor (layer.bn_partial), the stats pass is skipped.
    `fork`: return the activation TWICE (same storage) for its two consumers; their gradients are
    then summed inside the backward kernel instead of by a separate elementwise add."""
    if y.dim() != 4 or y.shape[1] != self.channels:
      raise ValueError('expected [N,%d,H,W]' % self.channels)
    partial = None
    if producer is not None and getattr(producer, 'bn_partial', None) is not None:
      part, nrows, ptr = producer.bn_partial
      if ptr == y.data_ptr() and y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last):
        partial = (part, nrows)
      producer.bn_partial = None
    y = y.contiguous(memory_format=torch.channels_last)
    if y.dtype != torch.bfloat16:
      y = y.to(torch.bfloat16)
    if residual is not None:
      residual = residual.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    if self.training:
      return _BNFn.apply(y, self.weight, self.bias, residual, self, partial, fork)
    scale = self.weight.detach() * torch.rsqrt(self.running_var + self.eps)
    shift = self.bias.detach() - self.running_mean * scale
    out = torch.empty_like(y, memory_format=torch.channels_last)
    n, c, h, w = y.shape
    _cabi.check(_cabi.lib().rigl_bn_apply(y.data_ptr(), _p(residual), scale.data_ptr(), shift.data_ptr(),
                                          n * h * w, c, int(self.relu), out.data_ptr(), _cabi.stream_ptr()),
                'rigl_bn_apply')
    return (out, out) if fork else out


class _MaxPoolFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, ksize, stride):
    n, c, h, w = x.shape
    oh, ow = (h + stride - 1) // stride, (w + stride - 1) // stride
    y = torch.empty((n, c, oh, ow), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    arg = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device)
    _cabi.check(_cabi.lib().rigl_maxpool_same_forward(x.data_ptr(), n, h, w, c, ksize, stride, y.data_ptr(),
                                                      arg.data_ptr(), _cabi.stream_ptr()),
                'rigl_maxpool_same_forward')
    ctx.save_for_backward(arg)
    ctx.geom = (n, h, w, c, ksize, stride)
    return y

  @staticmethod
  def backward(ctx, dy):
    arg, = ctx.saved_tensors
    n, h, w, c, ksize, stride = ctx.geom
    dy = dy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dx = torch.empty((n, c, h, w), dtype=torch.bfloat16, device=dy.device, memory_format=torch.channels_last)
    _cabi.check(_cabi.lib().rigl_maxpool_same_backward(dy.data_ptr(), arg.data_ptr(), n, h, w, c, ksize, stride,
                                                       dx.data_ptr(), _cabi.stream_ptr()),
                'rigl_maxpool_same_backward')
    return dx, None, None


def max_pool_same(x, ksize=3, stride=2):
  """tf.layers.max_pooling2d(pool_size, strides, padding='SAME') on NHWC bf16
  (resnet_model.py:636-642); x is [N,C,H,W] channels_last."""
  x = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
  return _MaxPoolFn.apply(x, int(ksize), int(stride))
