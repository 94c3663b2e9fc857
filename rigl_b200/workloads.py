"""Workload definitions for the hot path: the masked-layer graphs of the
reference's models, built on rigl_b200.layers, plus the train-step harness.

Only what BASELINE.json's configs need -- the shapes, variable names and wiring
that the masked conv/linear kernels and the RigL update run on:
  ResNet50   rigl/imagenet_resnet/resnet_model.py:396-731 (v1.5: stride on the 3x3;
             BN after every conv, zero-init gamma on the last BN of a block)
  MnistFC    rigl/mnist/mnist_train_eval.py:112-160 (784-300-100-10, all masked)
BN+ReLU(+residual) run on the fused streaming kernels of csrc/bn.cu;
pooling / loss are stock PyTorch kernels over channels_last bf16 tensors.
"""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from . import pruning
from . import sparse_utils
from .layers import SparseConv2d, SparseLinear, variance_scaling_
from .norm import FusedBatchNormReLU, max_pool_same
from .sparse_optimizers import SparseRigLOptimizer
from .sparse_optimizers_base import GlobalStep

BATCH_NORM_DECAY = 0.9
BATCH_NORM_EPSILON = 1e-5


class DenseConv2d(nn.Conv2d):
  """An un-masked conv of the reference models (WRN `conv_1`, MobileNet `initial_conv` / depthwise): float32
  master weights like every TF variable, bf16 compute (stock cuDNN kernels -- not a masked op)."""

  def forward(self, x):
    return F.conv2d(x, self.weight.to(torch.bfloat16), None, self.stride, self.padding, self.dilation, self.groups)


def _BNReLU(channels, relu=True, init_zero=False, device='cuda'):
  """batch_norm_relu (resnet_model.py:41-80) on the fused streaming kernels (csrc/bn.cu)."""
  return FusedBatchNormReLU(channels, relu=relu, init_zero=init_zero, eps=BATCH_NORM_EPSILON,
                            decay=BATCH_NORM_DECAY, device=device)


# Block outputs are handed to their two consumers as two handles so that the gradient sum happens
# inside the BN backward kernel (rigl_bn_backward2) rather than in a separate elementwise add.
FORK_BLOCK_OUTPUTS = True


class _Bottleneck(nn.Module):

  def __init__(self, cin, filters, strides, use_projection, name, device, registry):
    super(_Bottleneck, self).__init__()
    mk = lambda ci, co, k, s, n: SparseConv2d(ci, co, k, strides=s, padding='FIXED', name='resnet_model/' + n,
                                              device=device, registry=registry)
    self.proj = None
    if use_projection:
      self.proj = mk(cin, 4 * filters, 1, strides, 'bottleneck_projection_%s' % name)
      self.proj_bn = _BNReLU(4 * filters, relu=False, device=device)
    self.conv1 = mk(cin, filters, 1, 1, 'bottleneck_1_%s' % name)
    self.bn1 = _BNReLU(filters, device=device)
    self.conv2 = mk(filters, filters, 3, strides, 'bottleneck_2_%s' % name)
    self.bn2 = _BNReLU(filters, device=device)
    self.conv3 = mk(filters, 4 * filters, 1, 1, 'bottleneck_3_%s' % name)
    for conv in (self.proj, self.conv1, self.conv2, self.conv3):
      if conv is not None:
        conv.collect_bn_stats = True
    # last BN of the block: zero-init gamma; the residual add + final ReLU are fused into it
    self.bn3 = _BNReLU(4 * filters, relu=True, init_zero=True, device=device)

  def forward(self, x, x_skip=None, fork=False):
    """`x` feeds conv1, `x_skip` (the same activation, second handle) the shortcut branch; with
    `fork` the block output comes back as two handles as well (see FusedBatchNormReLU.forward)."""
    # every conv feeds a BN: its epilogue emits that BN's batch statistics (producer=...)
    if x_skip is None:
      x_skip = x
    shortcut = x_skip if self.proj is None else self.proj_bn(self.proj(x_skip), producer=self.proj)
    y = self.bn1(self.conv1(x), producer=self.conv1)
    y = self.bn2(self.conv2(y), producer=self.conv2)
    return self.bn3(self.conv3(y), residual=shortcut, producer=self.conv3, fork=fork)   # relu(BN(conv3) + shortcut)


class ResNet50(nn.Module):
  """ResNet-50 with every conv and the classifier masked (54 masked tensors,
  names = the reference variable scopes)."""

  def __init__(self, num_classes=1000, device='cuda', registry=None):
    super(ResNet50, self).__init__()
    self.registry = registry if registry is not None else pruning.MaskedLayerRegistry()
    reg = self.registry
    self.initial_conv = SparseConv2d(3, 64, 7, strides=2, padding='FIXED', name='resnet_model/initial_conv',
                                     device=device, registry=reg)
    self.initial_bn = _BNReLU(64, device=device)
    self.initial_conv.collect_bn_stats = True
    blocks = []
    cin = 64
    for g, (filters, n_blocks, stride) in enumerate(((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)), 1):
      blocks.append(_Bottleneck(cin, filters, stride, True, 'block_group_projection_block_group%d' % g,
                                device, reg))
      cin = 4 * filters
      for b in range(1, n_blocks):
        blocks.append(_Bottleneck(cin, filters, 1, False, 'block_group%d_%d_1' % (g, b), device, reg))
    self.blocks = nn.ModuleList(blocks)
    self.final_dense = SparseLinear(
        2048, num_classes, name='resnet_model/final_dense', device=device, registry=reg,
        out_dtype=torch.float32,
        kernel_initializer=lambda w: w.normal_(0., .01))      # resnet_model.py:713

  def forward(self, x):
    x = self.initial_bn(self.initial_conv(x), producer=self.initial_conv)
    x = max_pool_same(x, 3, 2)                      # 'SAME' 3x3/2 pool, resnet_model.py:636-642
    x_skip, last = x, len(self.blocks) - 1
    for i, blk in enumerate(self.blocks):
      if i < last and FORK_BLOCK_OUTPUTS:
        x, x_skip = blk(x, x_skip, fork=True)
      else:
        x = blk(x, x_skip)
        x_skip = x
    x = x.mean(dim=(2, 3))
    return self.final_dense(x)


class WideResNet(nn.Module):
  """Pre-activation WideResNet-(6n+4)-k, cifar_resnet/resnet_model.py:70-235 (BASELINE C5:
  depth 22, width 2).  `conv_1` (3x3x3x16) is a plain dense conv unless prune_first_layer
  (resnet_train_eval.py:96); residual 3x3 convs use TF 'SAME', the 1x1 skip convs 'VALID'
  with the block stride; dropout 0.3 between the two convs of a block."""

  def __init__(self, depth=22, width=2, num_classes=10, droprate=0.3, device='cuda', registry=None):
    super(WideResNet, self).__init__()
    if (depth - 4) % 6 != 0:
      raise ValueError('Depth of ResNet specified not sufficient.')
    self.registry = registry if registry is not None else pruning.MaskedLayerRegistry()
    reg, n_blocks = self.registry, (depth - 4) // 6
    self.conv_1 = DenseConv2d(3, 16, 3, padding=1, bias=False, device=device)
    self.droprate = droprate
    blocks, cin = [], 16
    for name, size, subsample in (('conv_2', 16 * width, False), ('conv_3', 32 * width, True),
                                  ('conv_4', 64 * width, True)):
      for n in range(n_blocks):
        stride = 2 if (subsample and n == 0) else 1
        blk = nn.Module()
        blk.bn_a = _BNReLU(cin, device=device)
        blk.skip = None
        if cin != size:
          blk.skip = SparseConv2d(cin, size, 1, strides=stride, padding='VALID',
                                  name='resnet_model/skip_%s' % name, device=device, registry=reg)
        blk.conv_a = SparseConv2d(cin, size, 3, strides=stride, padding='SAME',
                                  name='resnet_model/%s_%d_1' % (name, n), device=device, registry=reg)
        blk.bn_b = _BNReLU(size, device=device)
        blk.conv_b = SparseConv2d(size, size, 3, strides=1, padding='SAME',
                                  name='resnet_model/%s_%d_2' % (name, n), device=device, registry=reg)
        blocks.append(blk)
        cin = size
    self.blocks = nn.ModuleList(blocks)
    self.final_bn = _BNReLU(cin, device=device)
    self.logits = SparseLinear(cin, num_classes, name='resnet_model/logits', device=device, registry=reg,
                               out_dtype=torch.float32)

  def forward(self, x):
    net = self.conv_1(x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
    for blk in self.blocks:
      skip = net
      net = blk.bn_a(net)
      if blk.skip is not None:
        skip = blk.skip(net)
      net = blk.conv_a(net)
      net = F.dropout(blk.bn_b(net), self.droprate, self.training)
      net = blk.conv_b(net) + skip
    net = self.final_bn(net)
    return self.logits(net.mean(dim=(2, 3)))


class MobileNetV1(nn.Module):
  """MobileNet-v1 as the reference sparsifies it (mobilenetv1_model.py:156-342, BASELINE C4):
  only the 13 pointwise 1x1 convs and `final_dense` are masked; `initial_conv` and the
  depthwise 3x3 convs are dense (stock grouped convs)."""

  CFG = ((64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1), (512, 1), (512, 1),
         (512, 1), (1024, 2), (1024, 1))

  def __init__(self, num_classes=1000, device='cuda', registry=None):
    super(MobileNetV1, self).__init__()
    self.registry = registry if registry is not None else pruning.MaskedLayerRegistry()
    reg = self.registry
    self.initial_conv = DenseConv2d(3, 32, 3, stride=2, padding=1, bias=False, device=device)
    self.initial_bn = _BNReLU(32, device=device)
    blocks, cin = [], 32
    for i, (filters, stride) in enumerate(self.CFG):
      blk = nn.Module()
      blk.depthwise = DepthwiseConv2d(cin, stride=stride, device=device)
      blk.bn_dw = _BNReLU(cin, device=device)
      blk.pointwise = SparseConv2d(cin, filters, 1, strides=1, padding='FIXED',
                                   name='resnet_model/contraction_1x1_%d' % i, device=device, registry=reg)
      blk.bn_pw = _BNReLU(filters, device=device)
      blocks.append(blk)
      cin = filters
    self.blocks = nn.ModuleList(blocks)
    self.final_dense = SparseLinear(cin, num_classes, name='resnet_model/final_dense', device=device,
                                    registry=reg, out_dtype=torch.float32)

  def forward(self, x):
    x = self.initial_bn(self.initial_conv(x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)))
    for blk in self.blocks:
      x = blk.bn_dw(blk.depthwise(x))
      x = blk.bn_pw(blk.pointwise(x))
    return self.final_dense(x.mean(dim=(2, 3)))


class MnistFC(nn.Module):
  """mnist_network_fc with model_pruning=True: three masked dense layers + ReLU."""

  def __init__(self, hidden=(300, 100), device='cuda', registry=None):
    super(MnistFC, self).__init__()
    self.registry = registry if registry is not None else pruning.MaskedLayerRegistry()
    dims = (784,) + tuple(hidden) + (10,)
    self.layers = nn.ModuleList([
        SparseLinear(dims[i], dims[i + 1], name='layer%d' % (i + 1), device=device,
                     registry=self.registry,
                     out_dtype=torch.float32 if i == len(dims) - 2 else torch.bfloat16)
        for i in range(len(dims) - 1)])

  def forward(self, x):
    for i, l in enumerate(self.layers):
      x = l(x)
      if i + 1 < len(self.layers):
        x = F.relu(x)
    return x


def init_masks(model, method, sparsity, custom_sparsity_map=None, seed=0, erk_power_scale=1.0):
  """Runs the reference's mask-init path (get_mask_init_fn) on a model's masks."""
  np.random.seed(seed)
  fn = sparse_utils.get_mask_init_fn(model.registry.get_masks(), method, sparsity,
                                     custom_sparsity_map or {}, erk_power_scale=erk_power_scale)
  return fn()


class TrainHarness(object):
  """One sparse training step, wired like imagenet_train_eval.py:355-430:
  Nesterov momentum 0.9, weight decay on the raw weights, label smoothing 0.1,
  SparseRigLOptimizer(drop 0.3 cosine, every 100 steps), optional data parallelism."""

  def __init__(self, model, lr=0.1, momentum=0.9, weight_decay=1e-4, label_smoothing=0.1,
               drop_fraction=0.3, drop_fraction_anneal='cosine', begin_step=0, end_step=25000,
               frequency=100, data_parallel=None, optimizer_cls=SparseRigLOptimizer, lr_schedule=None,
               fused_optimizer=None):
    """lr_schedule: optional callable(global_step) -> learning rate, evaluated on the host before every step
    (optim.make_imagenet_lr_fn is the reference's, imagenet_train_eval.py:317-354); `lr` is then only the
    initial value.  fused_optimizer: the inner optimizer is optim.FusedMomentumSGD (one launch, mask * dense_grad
    fused in, device-resident learning rate); default on for CUDA models (RIGL_FUSED_SGD=0 -> torch.optim.SGD)."""
    import os
    self.model = model
    self.label_smoothing = label_smoothing
    self.lr_schedule = lr_schedule
    on_cuda = next(model.parameters()).is_cuda
    if fused_optimizer is None:
      fused_optimizer = on_cuda and os.environ.get('RIGL_FUSED_SGD', '1') != '0'
    self.fused = bool(fused_optimizer)
    if self.fused:
      from .optim import FusedMomentumSGD
      self.inner = FusedMomentumSGD(model.parameters(), lr=lr, momentum=momentum, nesterov=True,
                                    weight_decay=weight_decay)
    else:
      self.inner = torch.optim.SGD(model.parameters(), lr=lr, momentum=momentum, nesterov=True,
                                   weight_decay=weight_decay, foreach=True)
    self.opt = optimizer_cls(self.inner, begin_step, end_step, frequency, drop_fraction=drop_fraction,
                             drop_fraction_anneal=drop_fraction_anneal,
                             use_tpu=data_parallel is not None).bind(model.registry)
    self.global_step = GlobalStep(0)
    self.dp = data_parallel
    # Gradient exchange under data parallelism.  Default: ONE all-reduce of the flat buffer between the two graph
    # replays (backward; optimizer).  RIGL_DP_OVERLAP=1: bucketed all-reduces launched from inside backward on a
    # communication stream and captured with the step.  Measured on 8 x B200 (profiles/r02_bench_n8*.json): 22.12 vs
    # 22.21 ms per step (N = 1: 21.6) -- the 102 MB exchange takes ~0.4 ms over NVSwitch, and hiding it costs as
    # much in SM contention as it saves, so the simpler form is the default.
    self._dp_overlap = os.environ.get('RIGL_DP_OVERLAP', '0') == '1'
    self._pack_ahead = on_cuda and os.environ.get('RIGL_PACK_AHEAD', '1') != '0'
    if self.dp is not None:
      self.dp.attach(model)
    if self.fused:
      # masked layers: the optimizer reads dense_grad (+ bitmap) directly; under data parallelism the buffers
      # hold the SUM over replicas, the weight update takes the mean (CrossShardOptimizer)
      inv = 1.0 / self.dp.world if self.dp is not None else 1.0
      self.inner.attach_masked_layers(model.registry.layers(), grad_scale=inv, other_grad_scale=inv)
      if self.dp is not None:
        self.dp.masked_grads_in_optimizer = True
        self.dp.other_scale_in_optimizer = True

  def _apply_lr_schedule(self):
    if self.lr_schedule is None:
      return
    lr = float(self.lr_schedule(self.global_step.value))
    if self.fused:
      self.inner.set_lr(lr)
    else:
      if getattr(self, 'graphed', False) and lr != self.inner.param_groups[0]['lr']:
        raise RuntimeError('torch.optim.SGD bakes the learning rate into a captured graph: use the fused optimizer '
                           'with an lr schedule')
      for g in self.inner.param_groups:
        g['lr'] = lr

  # ---- CUDA-graph mode: the forward+backward and the inner optimizer step are captured once and
  # replayed (inter-kernel launch gaps and all host work disappear); the data-parallel
  # all-reduce, the schedule logic and the (rare) mask update stay eager between the replays.
  def enable_cuda_graph(self, images, labels, warmup=3, overlap_wgrad=None):
    """Captures the step for fixed input shapes.  Returns False (and stays eager) if capture fails.
    overlap_wgrad: run the dense wgrad kernels on a forked stream inside the graph (layers.WGRAD_SIDE_STREAM);
    default from RIGL_WGRAD_OVERLAP (on unless '0')."""
    import os
    if overlap_wgrad is None:
      # default on (RIGL_WGRAD_OVERLAP=0 keeps the serial backward); under data parallelism the bucketed
      # all-reduces run behind the forked wgrad kernels on a third stream
      overlap_wgrad = os.environ.get('RIGL_WGRAD_OVERLAP', '1') != '0' 
    self._overlap = bool(overlap_wgrad)
    self._sx, self._sy = images.clone(), labels.clone()
    ok = self._capture(warmup)
    if not ok and self.dp is not None and self._dp_overlap:
      # the collectives could not be captured on this stack: capture the step without them and all-reduce
      # (blocking, one call) between the two replays instead of falling back to the eager step
      self._dp_overlap = False
      ok = self._capture(warmup)
    return ok

  def release_cuda_graph(self):
    """Drops the captured graphs (back to the eager step).  Needed before torch.distributed is shut down: NCCL does
    not finalise a communicator while graphs that captured its collectives exist."""
    self.graphed = False
    for name in ('_g_fb', '_g_opt', '_sloss'):
      if hasattr(self, name):
        setattr(self, name, None)

  def _capture(self, warmup):
    try:
      side = torch.cuda.Stream()
      side.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(side):
        for _ in range(warmup):
          self._forward_backward(self._sx, self._sy, set_to_none=False)
      torch.cuda.current_stream().wait_stream(side)
      torch.cuda.synchronize()
      from . import _cabi
      self._g_fb = torch.cuda.CUDAGraph()
      before = _cabi.launch_count()
      # (thread-local capture mode: the NCCL watchdog thread keeps issuing CUDA calls while we capture)
      with torch.cuda.graph(self._g_fb, capture_error_mode='thread_local'):
        self._sloss = self._forward_backward(self._sx, self._sy, set_to_none=False)
      self.graph_kernel_launches = _cabi.launch_count() - before     # rigl kernels inside one replay
      self.replayed_kernel_launches = 0
      self._g_opt = torch.cuda.CUDAGraph()
      if self.fused:
        self.inner.prepare()                   # slots / device lr / launch plan: allocated OUTSIDE the capture
      with torch.cuda.graph(self._g_opt, pool=self._g_fb.pool(), capture_error_mode='thread_local'):
        self.inner.step()
      self.graphed = True
    except Exception as e:      # stay on the eager path, but say why
      import warnings
      warnings.warn('CUDA-graph capture failed, running eagerly: %r' % (e,))
      torch.cuda.synchronize()
      self.graphed = False
    return self.graphed

  def _forward_backward(self, images, labels, set_to_none):
    from . import layers
    for mw in self.model.registry.get_masked_weights():
      mw.fresh = False
    self.inner.zero_grad(set_to_none=set_to_none)
    if self._pack_ahead:
      # ONE launch packs the operands (mask * W -> bf16, both layouts) of every layer
      layers.pack_all(self.model.registry.layers())
    logits = self.model(images)
    loss = F.cross_entropy(logits.float(), labels, label_smoothing=self.label_smoothing)
    layers.WGRAD_SIDE_STREAM = bool(getattr(self, '_overlap', False))
    layers.MASKED_GRAD_IN_OPTIMIZER = self.fused      # mask * dense_grad is formed inside the optimizer kernel
    if self.dp is not None and self._dp_overlap:
      self.dp.begin_backward()                        # bucketed all-reduces launched from inside backward
      layers.DP_HOOK = self.dp
    try:
      loss.backward()
    finally:
      layers.WGRAD_SIDE_STREAM = False
      layers.MASKED_GRAD_IN_OPTIMIZER = False
      layers.DP_HOOK = None
      if self.dp is not None and self._dp_overlap:
        # (only a side stream that was forked in THIS backward may be waited on: under capture a wait on a
        #  stream outside the capture is an error)
        side = list(layers._SIDE.values()) if getattr(self, '_overlap', False) else []
        self.dp.finish(self.model, producer_streams=side)       # head bucket + join of the communication stream
      layers.join_side_streams()                # (no-op when nothing was forked)
    return loss

  def _graphed_step(self, images, labels):
    self._apply_lr_schedule()
    self._sx.copy_(images, non_blocking=True)
    self._sy.copy_(labels, non_blocking=True)
    self._g_fb.replay()
    self.replayed_kernel_launches += self.graph_kernel_launches
    if self.dp is not None and not self._dp_overlap:
      self.dp.reduce_gradients(self.model)
    self.opt.collect_masked_grads()
    gs = self.global_step
    self.opt._global_step = gs
    # same decision as SparseRigLOptimizerBase.apply_gradients, with the inner step replayed
    def inner_step():
      self._g_opt.replay()
      gs.increment()
    self.opt.cond_mask_update_op(gs, inner_step)
    return self._sloss

  def step(self, images, labels):
    """images: bf16 [N,3,H,W] channels_last; labels: int64 [N].  Returns the loss tensor."""
    if getattr(self, 'graphed', False):
      return self._graphed_step(images, labels)
    self._apply_lr_schedule()
    # without DP the grads are re-created by autograd (no zero-fill, no accumulate pass);
    # with DP they are views of the flat all-reduce buffer and must persist
    loss = self._forward_backward(images, labels, set_to_none=self.dp is None)
    if self.dp is not None and not self._dp_overlap:
      self.dp.reduce_gradients(self.model)
    self.opt.collect_masked_grads()
    self.opt.apply_gradients(None, global_step=self.global_step)
    return loss


class _DepthwiseFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x, weight, stride):
    from . import _cabi
    n, c, h, w = x.shape
    oh, ow = (h - 1) // stride + 1, (w - 1) // stride + 1
    y = torch.empty((n, c, oh, ow), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    _cabi.check(_cabi.lib().rigl_depthwise3x3_fprop(x.data_ptr(), weight.data_ptr(), n, h, w, c, stride, y.data_ptr(),
                                                    _cabi.stream_ptr()), 'rigl_depthwise3x3_fprop')
    ctx.save_for_backward(x, weight)
    ctx.stride = stride
    return y

  @staticmethod
  def backward(ctx, dy):
    from . import _cabi
    from .layers import _workspace
    x, weight = ctx.saved_tensors
    n, c, h, w = x.shape
    dy = dy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dx = dw = None
    if ctx.needs_input_grad[0]:
      dx = torch.empty_like(x, memory_format=torch.channels_last)
      _cabi.check(_cabi.lib().rigl_depthwise3x3_dgrad(dy.data_ptr(), weight.data_ptr(), n, h, w, c, ctx.stride,
                                                      dx.data_ptr(), _cabi.stream_ptr()), 'rigl_depthwise3x3_dgrad')
    if ctx.needs_input_grad[1]:
      dw = torch.empty_like(weight)
      ws = _workspace(x.device, _cabi.lib().rigl_depthwise3x3_workspace_bytes(n, h, w, c, ctx.stride))
      _cabi.check(_cabi.lib().rigl_depthwise3x3_wgrad(x.data_ptr(), dy.data_ptr(), n, h, w, c, ctx.stride, dw.data_ptr(),
                                                      0.0, ws.data_ptr(), ws.numel(), _cabi.stream_ptr()),
                  'rigl_depthwise3x3_wgrad')
    return dx, dw, None


class DepthwiseConv2d(nn.Module):
  """depthwise_conv2d_fixed_padding(kernel_size=3) of the reference's MobileNet-v1 (mobilenetv1_model.py:120-153):
  dense (un-masked), fp32 master weights in the torch depthwise layout [C,1,3,3], bf16 compute on the streaming
  kernels of csrc/depthwise.cu when RIGL_NATIVE_DEPTHWISE=1; default: the stock cuDNN grouped conv, which was
  the faster of the two on B200 when measured."""

  def __init__(self, channels, stride=1, device='cuda'):
    super(DepthwiseConv2d, self).__init__()
    import math
    import os
    self.channels, self.stride = int(channels), int(stride)
    self.weight = nn.Parameter(torch.empty(channels, 1, 3, 3, device=device))
    nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
    self.native = os.environ.get('RIGL_NATIVE_DEPTHWISE', '0') == '1'

  def forward(self, x):
    if self.native and x.is_cuda and self.channels % 8 == 0:
      x = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
      return _DepthwiseFn.apply(x, self.weight, self.stride)
    return F.conv2d(x, self.weight.to(torch.bfloat16), None, self.stride, 1, 1, self.channels)
