"""ResNet-50 sparse train step on the CUDA hot path (small batch / image so it runs in seconds)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from rigl_b200 import workloads

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _reference_forward(model, x):
  """Same network with stock torch convs on mask*W (bf16, channels_last)."""
  def conv(layer, t):
    w = (layer.weight.detach() * layer.mask.to_dense()).to(torch.bfloat16).permute(3, 2, 0, 1).contiguous()
    return F.conv2d(t, w, stride=layer.stride, padding=layer.pad)
  t = model.initial_bn(conv(model.initial_conv, x))
  t = F.max_pool2d(F.pad(t, (0, 1, 0, 1), value=float('-inf')), 3, 2, 0)     # TF 'SAME': pad at the end only
  for blk in model.blocks:
    sc = t if blk.proj is None else blk.proj_bn(conv(blk.proj, t))
    y = blk.bn1(conv(blk.conv1, t))
    y = blk.bn2(conv(blk.conv2, y))
    t = blk.bn3(conv(blk.conv3, y), residual=sc)
  t = t.mean(dim=(2, 3))
  fc = model.final_dense
  return t.float() @ (fc.weight.detach() * fc.mask.to_dense()).to(torch.bfloat16).float() + fc.bias.detach()


def test_resnet50_layer_table_and_erk_counts(golden):
  torch.manual_seed(0)
  model = workloads.ResNet50(device=DEV)
  case = [c for c in golden['cases'] if c['tag'] == 'r50_erk80'][0]
  names = [m.name for m in model.registry.get_masks()]
  assert names == [n + '/mask:0' for n, _ in case['layers']]
  assert [list(m.shape) for m in model.registry.get_masks()] == [sh for _, sh in case['layers']]
  sp = workloads.init_masks(model, 'erdos_renyi_kernel', 0.8, seed=0)
  for m in model.registry.get_masks():
    assert float(sp[m.name]).hex() == case['sparsities_hex'][m.name]
    assert m.count_ones() == case['nnz'][m.name]


def test_forward_matches_stock_torch_convs():
  torch.manual_seed(1)
  model = workloads.ResNet50(device=DEV)
  workloads.init_masks(model, 'erdos_renyi_kernel', 0.8, seed=1)
  for blk in model.blocks:                      # make the residual branch non-trivial
    torch.nn.init.ones_(blk.bn3.weight)
  model.eval()                                  # BN in inference mode: deterministic comparison
  x = torch.randn(4, 3, 64, 64, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
  with torch.no_grad():
    got = model(x)
    want = _reference_forward(model, x)
  scale = float(want.abs().max())
  assert torch.isfinite(got).all()
  assert float((got - want).abs().max()) <= 5e-2 * scale      # 50+ bf16 roundings deep


def test_train_steps_update_semantics_and_conservation():
  torch.manual_seed(2)
  model = workloads.ResNet50(device=DEV)
  workloads.init_masks(model, 'erdos_renyi_kernel', 0.8, seed=2)
  ones_before = [m.count_ones() for m in model.registry.get_masks()]
  h = workloads.TrainHarness(model, lr=0.05, frequency=3, end_step=100, fused_optimizer=False)   # (weight.grad is read below)
  x = torch.randn(8, 3, 64, 64, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
  y = torch.randint(0, 1000, (8,), device=DEV)
  incs, losses = [], []
  for _ in range(8):
    before = h.global_step.value
    losses.append(float(h.step(x, y)))
    incs.append(h.global_step.value - before)
  assert incs == [0, 1, 1, 1, 0, 1, 1, 1]       # gs 0 and gs 3 are update iterations
  assert all(np.isfinite(losses))
  assert [m.count_ones() for m in model.registry.get_masks()] == ones_before
  # dense gradient is dense; masked gradient is zero off-mask
  l = model.blocks[3].conv2
  assert float((l.masked_weights.dense_grad != 0).float().mean()) > 0.9
  off = l.mask.to_dense() == 0
  assert float(l.weight.grad[off].abs().max()) == 0.0
  assert losses[-1] < losses[0] * 1.5


def test_mnist_fc_trains():
  torch.manual_seed(3)
  model = workloads.MnistFC(device=DEV)
  workloads.init_masks(model, 'random', 0.9, {'layer2': 0.81, 'layer3': 0.0}, seed=3)
  assert [m.count_ones() for m in model.registry.get_masks()] == [23520, 5700, 1000]
  h = workloads.TrainHarness(model, lr=0.2, weight_decay=0.0, label_smoothing=0.0, frequency=100, end_step=50000)
  x = torch.randn(100, 784, device=DEV)
  y = (x[:, :10].argmax(1)).long()
  first = float(h.step(x, y))
  for _ in range(60):
    last = float(h.step(x, y))
  assert last < 0.7 * first
  assert [m.count_ones() for m in model.registry.get_masks()] == [23520, 5700, 1000]


def test_wrn22_2_layer_table_and_step(golden):
  """BASELINE C5 workload: names/shapes/ERK-0.95 counts equal the reference-generated fixture."""
  torch.manual_seed(4)
  model = workloads.WideResNet(depth=22, width=2, device=DEV)
  case = [c for c in golden['cases'] if c['tag'] == 'wrn22_2_erk95'][0]
  got = sorted((m.name, list(m.shape)) for m in model.registry.get_masks())
  want = sorted((n + '/mask:0', sh) for n, sh in case['layers'])
  assert got == want
  sp = workloads.init_masks(model, 'erdos_renyi_kernel', 0.95, seed=4)
  for m in model.registry.get_masks():
    assert float(sp[m.name]).hex() == case['sparsities_hex'][m.name] and m.count_ones() == case['nnz'][m.name]
  h = workloads.TrainHarness(model, lr=0.1, weight_decay=5e-4, label_smoothing=0.0, frequency=100, end_step=75000)
  x = torch.randn(32, 3, 32, 32, device=DEV)
  y = torch.randint(0, 10, (32,), device=DEV)
  losses = [float(h.step(x, y).detach()) for _ in range(6)]
  assert all(np.isfinite(losses)) and h.global_step.value == 5        # step 0 was the mask update
  assert [m.count_ones() for m in model.registry.get_masks()] == \
      [case['nnz'][m.name] for m in model.registry.get_masks()]


def test_mobilenet_v1_masked_pointwise_step(golden):
  """BASELINE C4 workload: 13 pointwise convs + classifier masked at 0.9 (overall ~0.89)."""
  torch.manual_seed(5)
  model = workloads.MobileNetV1(device=DEV)
  case = [c for c in golden['cases'] if c['tag'] == 'mobilenetv1_uniform90'][0]
  assert [(m.name, list(m.shape)) for m in model.registry.get_masks()] == \
      [(n + '/mask:0', sh) for n, sh in case['layers']]
  workloads.init_masks(model, 'random', 0.9, seed=5)
  assert [m.count_ones() for m in model.registry.get_masks()] == [case['nnz'][m.name] for m in model.registry.get_masks()]
  h = workloads.TrainHarness(model, lr=0.05, frequency=2, end_step=100)
  x = torch.randn(8, 3, 64, 64, device=DEV)
  y = torch.randint(0, 1000, (8,), device=DEV)
  incs = []
  for _ in range(4):
    before = h.global_step.value
    assert np.isfinite(float(h.step(x, y).detach()))
    incs.append(h.global_step.value - before)
  assert incs == [0, 1, 1, 0]


def test_cuda_graph_mode_matches_eager():
  """Graph-replayed steps produce the same masks / step counter behaviour and finite, decreasing loss."""
  def run(graph):
    torch.manual_seed(7)
    model = workloads.MnistFC(device=DEV)
    workloads.init_masks(model, 'random', 0.9, {'layer2': 0.81, 'layer3': 0.0}, seed=7)
    h = workloads.TrainHarness(model, lr=0.1, weight_decay=0.0, label_smoothing=0.0, frequency=4, end_step=1000)
    x = torch.randn(100, 784, device=DEV)
    y = (x[:, :10].argmax(1)).long()
    h.step(x, y)
    h.step(x, y)
    if graph:
      assert h.enable_cuda_graph(x, y)
    losses = [float(h.step(x, y).detach()) for _ in range(12)]
    return losses, h.global_step.value, [m.numpy().copy() for m in model.registry.get_masks()]
  le, ge, me = run(False)
  lg, gg, mg = run(True)
  assert ge == gg
  assert all(np.isfinite(lg)) and lg[-1] < lg[0]
  # the three graph warm-up passes do not move weights (no optimizer step), so trajectories agree closely
  assert np.allclose(le, lg, rtol=2e-2, atol=2e-3)
  for a, b in zip(me, mg):
    assert a.sum() == b.sum()


@pytest.mark.parametrize('fused', [False, True])
def test_cuda_graph_wgrad_side_stream_matches_serial_backward(fused):
  """In graph mode the dense wgrad kernels run on a forked stream (layers.WGRAD_SIDE_STREAM) next to the
  BN-backward chain.  Every kernel is deterministic, so replaying the captured forward+backward must
  reproduce the serial eager backward BIT FOR BIT (a missing dependency or a recycled buffer would not)."""
  torch.manual_seed(3)
  model = workloads.ResNet50(num_classes=10, device=DEV)
  workloads.init_masks(model, 'erdos_renyi_kernel', 0.8, seed=3)
  h = workloads.TrainHarness(model, lr=0.1, fused_optimizer=fused)
  x = torch.randn(8, 3, 64, 64, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
  y = torch.randint(0, 10, (8,), device=DEV)
  h._forward_backward(x, y, set_to_none=False)            # serial reference (no fork: _overlap is unset)
  torch.cuda.synchronize()
  layers_ = model.registry.layers()
  ref_dense = [l.masked_weights.dense_grad.clone() for l in layers_]
  # (with the fused inner optimizer mask * dense_grad is formed inside the step kernel: no weight.grad)
  ref_grad = [None if fused else l.weight.grad.clone() for l in layers_]
  assert h.enable_cuda_graph(x, y, overlap_wgrad=True) and h._overlap
  for _ in range(3):
    h._g_fb.replay()
    torch.cuda.synchronize()
    for l, d, g in zip(layers_, ref_dense, ref_grad):
      assert torch.equal(l.masked_weights.dense_grad, d), l.scope
      assert (l.weight.grad is None) if fused else torch.equal(l.weight.grad, g), l.scope


def test_fused_momentum_sgd_matches_torch_sgd_on_the_same_gradients():
  """optim.FusedMomentumSGD (one launch, mask * dense_grad formed in-kernel, device-resident lr) against
  torch.optim.SGD(nesterov, weight_decay) fed the materialised masked gradients: same trajectories up to fp32
  fma contraction, masked-out weights only decay, momentum slots agree, an lr change takes effect."""
  from rigl_b200 import pruning
  from rigl_b200.layers import SparseLinear
  from rigl_b200.optim import FusedMomentumSGD
  pruning.reset_default_registry()
  torch.manual_seed(9)
  la = SparseLinear(130, 77, name='a', device=DEV)          # 10010 weights: not a multiple of 4 -> scalar tail
  lb = SparseLinear(130, 77, name='b', device=DEV)
  rng = np.random.RandomState(9)
  m = (rng.rand(130, 77) > 0.7).astype(np.float32)
  la.mask.assign(m); lb.mask.assign(m)
  with torch.no_grad():
    lb.weight.copy_(la.weight); lb.bias.copy_(la.bias)
  fused = FusedMomentumSGD(la.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-2)
  fused.attach_masked_layers([la], grad_scale=0.5)
  ref = torch.optim.SGD(lb.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-2)
  for step in range(5):
    dense = torch.randn(130 * 77, device=DEV)
    gb = torch.randn(77, device=DEV)
    la.masked_weights.dense_grad.copy_(dense)
    la.bias.grad = gb.clone()
    lb.weight.grad = (dense.view(130, 77) * torch.from_numpy(m).to(DEV) * 0.5)
    lb.bias.grad = gb.clone()
    if step == 3:
      fused.set_lr(0.01)
      ref.param_groups[0]['lr'] = 0.01
    fused.step(); ref.step()
    for pa, pb in ((la.weight, lb.weight), (la.bias, lb.bias)):
      assert torch.allclose(pa, pb, rtol=1e-5, atol=1e-6), step
      assert torch.allclose(fused.state[pa]['momentum_buffer'], ref.state[pb]['momentum_buffer'], rtol=1e-5, atol=1e-6)
  assert la.weight.grad is None
