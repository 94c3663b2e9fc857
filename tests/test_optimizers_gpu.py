"""Port of the reference's optimizer tests (rigl/sparse_optimizers_test.py) to the
PyTorch/CUDA backend: same fixtures (one masked fully-connected layer, constant
inputs, analytically known gradients), same assertions, plus oracle parity of a
full RigL update through the public optimizer API."""
import itertools

import numpy as np
import pytest
import torch

from oracle import rigl_oracle as orc
from rigl_b200 import pruning, sparse_optimizers
from rigl_b200.layers import SparseLinear
from rigl_b200.sparse_optimizers_base import GlobalStep

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _setup_set(n_inp, n_out, drop_frac, start_iter=1, end_iter=4, freq_iter=2, cls=None):
  pruning.reset_default_registry()
  torch.manual_seed(0)
  np.random.seed(0)
  layer = SparseLinear(n_inp, n_out, name='fully_connected', device=DEV, out_dtype=torch.float32)
  optim = torch.optim.SGD(layer.parameters(), lr=0.1)
  cls = cls or sparse_optimizers.SparseSETOptimizer
  sparse_optim = cls(optim, start_iter, end_iter, freq_iter, drop_fraction=drop_frac)
  layer.mask.assign(np.random.choice([0, 1], size=(n_inp, n_out), p=[1. / 2, 1. / 2]))
  gs = GlobalStep(0)

  def train_op():
    x = torch.rand(1, n_inp, device=DEV)
    loss = layer(x).mean()
    sparse_optim.minimize(loss, gs)

  return train_op, layer, gs, sparse_optim


@pytest.mark.parametrize('n_inp,n_out,drop_frac', [(15, 25, 0.5), (15, 25, 0.2), (3, 5, 0.2)])
def test_set_mask_non_update_iterations(n_inp, n_out, drop_frac):
  train_op, layer, _, _ = _setup_set(n_inp, n_out, drop_frac)
  for i in range(1, 6):
    before = layer.mask.numpy()
    train_op()
    if i not in (1, 3):
      assert np.array_equal(before, layer.mask.numpy())


@pytest.mark.parametrize('n_inp,n_out,drop_frac', [(15, 25, 0.5), (15, 25, 0.7), (30, 10, 0.9)])
def test_set_update_iterations(n_inp, n_out, drop_frac):
  train_op, layer, _, _ = _setup_set(n_inp, n_out, drop_frac)
  for i in range(1, 5):
    before = layer.mask.numpy()
    train_op()
    after = layer.mask.numpy()
    if i in (1, 3):
      assert before.sum() == after.sum()
      assert not np.array_equal(before, after)


@pytest.mark.parametrize('start_iter,end_iter,freq_iter', [(3, 7, 2), (1, 5, 3), (0, 4, 1)])
def test_set_no_drop(start_iter, end_iter, freq_iter):
  train_op, layer, _, _ = _setup_set(3, 5, 0, start_iter, end_iter, freq_iter)
  for _ in range(end_iter + 2):
    before = layer.mask.numpy()
    train_op()
    assert np.array_equal(before, layer.mask.numpy())


def test_set_new_connection_zero_init():
  train_op, layer, _, _ = _setup_set(3, 5, 0.5, start_iter=0, end_iter=4, freq_iter=1)
  for _ in range(5):
    before = layer.mask.numpy()
    train_op()
    after = layer.mask.numpy()
    w = layer.weight.detach().cpu().numpy()
    assert np.all(w[np.logical_and(before == 0, after == 1)] == 0)


@pytest.mark.parametrize('shape,init_type', list(itertools.product(
    ((3, 7, 2), (5, 3), (1,)), ('zeros', 'random_normal', 'random_uniform'))))
def test_shape_and_dtype_of_grow_tensor(shape, init_type):
  pruning.reset_default_registry()
  p = torch.nn.Parameter(torch.rand(shape, device=DEV))
  so = sparse_optimizers.SparseSETOptimizer(torch.optim.SGD([p], lr=0.1), 0, 0, 1, use_stateless=False)
  for dtype in (torch.float32, torch.float64):
    w = torch.rand(shape, device=DEV, dtype=dtype) * 5
    g = so.get_grow_tensor(w, init_type)
    assert g.shape == w.shape and g.dtype == w.dtype


@pytest.mark.parametrize('method', ['ones', 'zero', None, 0])
def test_value_error_of_grow_tensor(method):
  p = torch.nn.Parameter(torch.rand(3, 4, device=DEV))
  so = sparse_optimizers.SparseSETOptimizer(torch.optim.SGD([p], lr=0.1), 0, 0, 1, use_stateless=False)
  with pytest.raises(ValueError):
    so.get_grow_tensor(p, method)


@pytest.mark.parametrize('n_inp,n_out,drop_frac', [(15, 25, 0.5), (15, 25, 0.7), (30, 10, 0.9)])
def test_static_mask_never_changes(n_inp, n_out, drop_frac):
  train_op, layer, _, _ = _setup_set(n_inp, n_out, drop_frac, cls=sparse_optimizers.SparseStaticOptimizer)
  first = layer.mask.numpy()
  for _ in range(5):
    train_op()
    assert np.array_equal(first, layer.mask.numpy())


def _setup_rigl(n_inp, n_out, drop_frac, start_iter=1, end_iter=4, freq_iter=2):
  pruning.reset_default_registry()
  layer = SparseLinear(n_inp, n_out, name='fully_connected', device=DEV, out_dtype=torch.float32)
  optim = torch.optim.SGD(layer.parameters(), lr=1e-3)
  gs = GlobalStep(0)
  so = sparse_optimizers.SparseRigLOptimizer(optim, start_iter, end_iter, freq_iter, drop_fraction=drop_frac)

  def train_op():
    x = torch.ones(1, n_inp, device=DEV)
    y = layer(x)
    scale = (torch.arange(y.numel(), device=DEV, dtype=y.dtype).reshape(y.shape) * float(gs.value))
    loss = (y * scale).sum()
    so.minimize(loss, gs)
    return scale

  return train_op, layer, gs, so


@pytest.mark.parametrize('n_inp,n_out', [(3, 4), (5, 2), (2, 5)])
def test_rigl_masked_gradient_calculation(n_inp, n_out):
  # sparse_optimizers_test.py:330-347: the dense gradient equals the broadcast scale vector.
  train_op, layer, gs, so = _setup_rigl(n_inp, n_out, 0., start_iter=0, end_iter=3, freq_iter=1)
  for _ in range(6):
    scale = train_op()
    dense = so._weight2masked_grads[layer.weight.name].view(n_inp, n_out)
    assert torch.equal(dense, scale.expand(n_inp, n_out).float())


@pytest.mark.parametrize('sched,expect', [
    ((3, 7, 2), [1, 1, 1, 0, 1, 1, 0, 1, 1, 0, 1]),
    ((1, 5, 3), [1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 1]),
    ((0, 4, 1), [0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1])])
def test_rigl_apply_gradients(sched, expect):
  train_op, layer, gs, _ = _setup_rigl(3, 5, .5, *sched)
  for one_if_incremented in expect:
    before = gs.value
    train_op()
    assert gs.value - before == one_if_incremented


def test_rigl_update_through_public_api_matches_oracle():
  """One real RigL update (momentum slots, injected noise stream) vs the oracle."""
  pruning.reset_default_registry()
  torch.manual_seed(1)
  rng = np.random.RandomState(1)
  layers = [SparseLinear(784, 300, name='layer1', device=DEV), SparseLinear(300, 100, name='layer2', device=DEV),
            SparseLinear(100, 10, name='layer3', device=DEV, out_dtype=torch.float32)]
  for l, s in zip(layers, (0.9, 0.81, 0.0)):
    l.mask.assign(orc.get_mask_random_numpy(tuple(l.weight.shape), s, rng))
  params = [p for l in layers for p in l.parameters()]
  optim = torch.optim.SGD(params, lr=0.2, momentum=0.9, nesterov=True)
  so = sparse_optimizers.SparseRigLOptimizer(optim, 0, 50000, 100, drop_fraction=0.3,
                                             drop_fraction_anneal='cosine', initial_acc_scale=0.25)
  gs = GlobalStep(0)
  x = torch.randn(100, 784, device=DEV)
  target = torch.randint(0, 10, (100,), device=DEV)

  def loss_fn():
    h = torch.relu(layers[0](x))
    h = torch.relu(layers[1](h))
    return torch.nn.functional.cross_entropy(layers[2](h).float(), target)

  # step 0 is an update iteration (last_update = -freq): no optimizer step, gs frozen
  so.minimize(loss_fn(), gs)
  assert gs.value == 0 and so.last_update_was_mask_update
  # a normal step creates the momentum slots
  so.minimize(loss_fn(), gs)
  assert gs.value == 1
  for _ in range(99):
    so.minimize(loss_fn(), gs)
  assert gs.value == 100
  # snapshot, then the update at gs=100
  snap = []
  loss = loss_fn()
  grads_and_vars = so.compute_gradients(loss)
  for l in layers:
    snap.append(dict(mask=l.mask.numpy(), w=l.weight.detach().cpu().numpy().copy(),
                     g=l.masked_weights.dense_grad.cpu().numpy().reshape(l.weight.shape).copy(),
                     mom=optim.state[l.weight]['momentum_buffer'].cpu().numpy().copy()))
  so.apply_gradients(grads_and_vars, gs)
  assert gs.value == 100 and so.last_update_was_mask_update
  frac = orc.get_drop_fraction('cosine', 0.3, 100, 0, 50000, True)
  assert np.float32(so.drop_fraction) == frac
  for l, s in zip(layers, snap):
    noise = so.last_update_noise(l.weight).cpu().numpy().reshape(s['w'].shape)       # drawn in-kernel, re-materialised
    want = orc.rigl_mask_update(s['mask'], s['w'], s['g'], frac, noise=noise, initial_acc_scale=0.25,
                                slots=[s['mom']])
    assert np.array_equal(l.mask.numpy(), want['mask'])
    assert l.weight.detach().cpu().numpy().tobytes() == want['weights'].tobytes()
    assert optim.state[l.weight]['momentum_buffer'].cpu().numpy().tobytes() == want['slots'][0].tobytes()
    assert l.mask.count_ones() == int(s['mask'].sum())


# ---- SparseMomentumOptimizer (sparse_optimizers.py:126-214).  The host logic is also covered on the CPU
# (tests/test_oracle_other_optimizers.py); these two run the real layers + select kernels.
@pytest.mark.parametrize('n_inp,n_out,momentum', [(3, 4, 0.5), (5, 2, 0.), (2, 5, 1.)])
def test_momentum_update(n_inp, n_out, momentum):
  """sparse_optimizers_test.py:276-295 (testMomentumUpdate)."""
  pruning.reset_default_registry()
  layer = SparseLinear(n_inp, n_out, name='fully_connected', device=DEV, out_dtype=torch.float32)
  optim = torch.optim.SGD(layer.parameters(), lr=0.1)
  gs = GlobalStep(0)
  so = sparse_optimizers.SparseMomentumOptimizer(optim, 1, 4, 2, drop_fraction=0.5, momentum=momentum)
  current = np.zeros((n_inp, n_out))
  for _ in range(6):
    x = torch.ones(1, n_inp, device=DEV)
    y = layer(x)
    loss = (y * torch.arange(y.numel(), device=DEV, dtype=y.dtype).reshape(y.shape)).sum()
    so.minimize(loss, gs)
    current = current * momentum + (1 - momentum) * np.arange(n_out)
    got = so.ema_average(layer.weight).view(n_inp, n_out).cpu().numpy()
    assert np.array_equal(got, current.astype(np.float32))


def test_momentum_mask_update_matches_oracle():
  pruning.reset_default_registry()
  torch.manual_seed(2)
  rng = np.random.RandomState(2)
  layer = SparseLinear(64, 48, name='layer1', device=DEV, out_dtype=torch.float32)
  layer.mask.assign(orc.get_mask_random_numpy((64, 48), 0.8, rng))
  optim = torch.optim.SGD(layer.parameters(), lr=0.05)
  so = sparse_optimizers.SparseMomentumOptimizer(optim, 0, 1000, 5, drop_fraction=0.3, momentum=0.9)
  gs = GlobalStep(1)
  x = torch.randn(32, 64, device=DEV)
  t = torch.randn(32, 48, device=DEV)
  loss_fn = lambda: ((layer(x) - t) ** 2).mean()
  for _ in range(4):                                    # steps 1..4: plain steps, EMA accumulates
    so.minimize(loss_fn(), gs)
  assert gs.value == 5
  gv = so.compute_gradients(loss_fn())
  mask0, w0 = layer.mask.numpy(), layer.weight.detach().cpu().numpy().copy()
  ema_before = so.ema_average(layer.weight).cpu().numpy().copy()
  g = layer.masked_weights.dense_grad.cpu().numpy().copy()
  so.apply_gradients(gv, gs)                            # SET-style: step, then the update on step 6? no: freq 5
  # (SET semantics, base.apply_gradients: the optimizer step runs first, then the mask update if due)
  ema_after = orc.momentum_ema_update(ema_before, g, 0.9)
  assert np.array_equal(so.ema_average(layer.weight).cpu().numpy(), ema_after)
  assert layer.mask.count_ones() == int(mask0.sum())


@pytest.mark.parametrize('clear', ['inner_zero_grad', 'module_zero_grad', 'set_to_none', 'never'])
def test_dense_grads_do_not_accumulate_across_steps_whichever_zero_grad(clear):
  """torch-style loop `backward(); opt.step()` where the caller clears gradients on the INNER optimizer /
  the module (or not at all): the dense gradient RigL ranks must be this step's gradient, like the
  reference's per-step compute_gradients (base.py:478-485), not a running sum."""
  pruning.reset_default_registry()
  layer = SparseLinear(6, 5, name='fully_connected', device=DEV, out_dtype=torch.float32)
  inner = torch.optim.SGD(layer.parameters(), lr=0.0)
  so = sparse_optimizers.SparseRigLOptimizer(inner, 100, 200, 50, drop_fraction=0.3)      # no update in range
  gs = GlobalStep(0)
  x = torch.ones(1, 6, device=DEV)
  coeff = torch.arange(5, device=DEV, dtype=torch.float32)
  for step in range(4):
    if clear == 'inner_zero_grad':
      inner.zero_grad()
    elif clear == 'module_zero_grad':
      layer.zero_grad()
    elif clear == 'set_to_none':
      inner.zero_grad(set_to_none=True)
    (layer(x) * coeff).sum().backward()
    g = layer.masked_weights.dense_grad.view(6, 5).cpu().numpy()
    assert np.array_equal(g, np.tile(np.arange(5, dtype=np.float32), (6, 1))), (clear, step)
    so.step(gs)
  assert gs.value == 4


def test_get_update_op_uses_scores_verbatim():
  """Public `_get_update_op` with NEGATIVE grow scores (the rigl_tf2 updaters pass -|g|): ranked as given."""
  pruning.reset_default_registry()
  rng = np.random.RandomState(3)
  layer = SparseLinear(40, 30, name='fully_connected', device=DEV, out_dtype=torch.float32)
  m = orc.get_mask_random_numpy((40, 30), 0.6, rng).astype(np.float32)
  layer.mask.assign(m)
  w = layer.weight.detach().cpu().numpy().copy()
  so = sparse_optimizers.SparseSETOptimizer(torch.optim.SGD(layer.parameters(), lr=0.1), 0, 10, 1, drop_fraction=0.5)
  so.drop_fraction = np.float32(0.5)
  sd = (np.abs(w) * m).astype(np.float32)
  sg = (-rng.rand(40, 30)).astype(np.float32)
  so._get_update_op(torch.from_numpy(sd).to(DEV), torch.from_numpy(sg).to(DEV), layer.mask, layer.weight)
  want = orc.get_update_op(sd, sg, m, w, np.float32(0.5))
  assert np.array_equal(layer.mask.numpy(), want['mask'])
  assert layer.weight.detach().cpu().numpy().tobytes() == want['weights'].tobytes()


# ---- SparseSnipOptimizer / SparseDNWOptimizer on the batched select kernels, against the masks produced by
# EXECUTING the reference's apply_gradients (tests/golden/snip_dnw_golden.json, tools/make_golden_snip_dnw.py).
import json as _json
import os as _os

with open(_os.path.join(_os.path.dirname(__file__), 'golden', 'snip_dnw_golden.json')) as _f:
  _SNIP_DNW = _json.load(_f)


def _gdec(e):
  return np.frombuffer(bytes.fromhex(e['hex']), dtype=np.dtype(e['dtype'])).reshape(e['shape']).copy()


@pytest.mark.parametrize('case', _SNIP_DNW['cases'], ids=[c['tag'] for c in _SNIP_DNW['cases']])
def test_snip_dnw_product_path_on_reference_executed_golden(case):
  from rigl_b200.layers import SparseConv2d
  pruning.reset_default_registry()
  layers_ = []
  for i, sh in enumerate(case['shapes'], 1):
    if len(sh) == 2:
      l = SparseLinear(sh[0], sh[1], use_bias=False, name='layer%d' % i, device=DEV)
    else:
      l = SparseConv2d(sh[2], sh[3], sh[0], name='layer%d' % i, device=DEV)
    with torch.no_grad():
      l.weight.copy_(torch.from_numpy(_gdec(case['weights'][i - 1])).to(DEV))
    layers_.append(l)
  params = [l.weight for l in layers_]
  inner = torch.optim.SGD(params, lr=0.0)                      # the golden generator's inner optimizer is a no-op
  gs = GlobalStep(0 if case['kind'] == 'snip' else 3)
  if case['kind'] == 'snip':
    so = sparse_optimizers.SparseSnipOptimizer(inner, case['sparsity'], case['method'], custom_sparsity_map=case['custom'])
    gv = [(torch.from_numpy(_gdec(g)).to(DEV), p) for g, p in zip(case['grads'], params)]
    assert so.apply_gradients(gv, global_step=gs) is True and so.is_snipped and gs.value == 0
    masks = [l.mask.numpy().copy() for l in layers_]
    assert so.apply_gradients(gv, global_step=gs) is False and gs.value == 1      # already snipped: plain step
    assert all(np.array_equal(a, l.mask.numpy()) for a, l in zip(masks, layers_))
  else:
    so = sparse_optimizers.SparseDNWOptimizer(inner, case['sparsity'], case['method'], custom_sparsity_map=case['custom'])
    gv = [(torch.zeros_like(p), p) for p in params]
    so.apply_gradients(gv, global_step=gs)
    assert gs.value == 4
    masks = [l.mask.numpy() for l in layers_]
  for i, (got, l) in enumerate(zip(masks, layers_)):
    want = _gdec(case['masks'][i])
    assert np.array_equal(got, want), (case['tag'], i)
    assert l.weight.detach().cpu().numpy().tobytes() == _gdec(case['weights'][i]).tobytes()    # weights untouched


def test_dnw_trains_with_dense_gradient_and_tracks_topk():
  """DNW end to end on a real layer: the weight update uses the DENSE gradient (masked-out weights move too) and
  after every step the mask equals the oracle's top-|w| mask of the UPDATED weights."""
  pruning.reset_default_registry()
  torch.manual_seed(5)
  layer = SparseLinear(48, 40, use_bias=False, name='layer1', device=DEV, out_dtype=torch.float32)
  layer.mask.assign(orc.get_mask_random_numpy((48, 40), 0.75, np.random.RandomState(5)))
  inner = torch.optim.SGD(layer.parameters(), lr=0.1)
  so = sparse_optimizers.SparseDNWOptimizer(inner, 0.75, 'random')
  gs = GlobalStep(0)
  x, t = torch.randn(16, 48, device=DEV), torch.randn(16, 40, device=DEV)
  for _ in range(3):
    m0, w0 = layer.mask.numpy().copy(), layer.weight.detach().cpu().numpy().copy()
    so.minimize(((layer(x) - t) ** 2).mean(), gs)
    w1 = layer.weight.detach().cpu().numpy()
    assert np.abs((w1 - w0)[m0 == 0]).max() > 0                  # masked-out weights received the dense gradient
    assert np.array_equal(layer.mask.numpy(), orc.dnw_mask(w1, 0.75))
  assert gs.value == 3


def test_tf2_style_schedule_drives_the_cuda_mask_update():
  """update_schedules.CosineUpdateSchedule (rigl_tf2/mask_updaters.py:251-344) over MaskUpdaterAdapter: at every
  update step the masks / weights equal the oracle's restatement of the TF2 generic_mask_update (:99-154) with the
  schedule's float32 drop fraction; `prune` drops without growing."""
  from rigl_b200 import update_schedules as us
  pruning.reset_default_registry()
  rng = np.random.RandomState(8)
  torch.manual_seed(8)
  layer = SparseLinear(60, 50, use_bias=False, name='layer1', device=DEV, out_dtype=torch.float32)
  layer.mask.assign(orc.get_mask_random_numpy((60, 50), 0.8, rng))
  inner = torch.optim.SGD(layer.parameters(), lr=0.05, momentum=0.9)
  so = sparse_optimizers.SparseRigLOptimizer(inner, 10 ** 9, 10 ** 9 + 1, 1, drop_fraction=0.3)   # own schedule never fires
  sched = us.CosineUpdateSchedule(us.MaskUpdaterAdapter(so), 0.3, update_freq=4, last_update_step=40)
  gs = GlobalStep(0)
  x, t = torch.randn(32, 60, device=DEV), torch.randn(32, 50, device=DEV)
  n_updates = 0
  for step in range(1, 13):
    gv = so.compute_gradients(((layer(x) - t) ** 2).mean())
    so.apply_gradients(gv, gs)
    if sched.is_update_iter(step):
      m0, w0 = layer.mask.numpy().copy(), layer.weight.detach().cpu().numpy().copy()
      g = layer.masked_weights.dense_grad.view(60, 50).cpu().numpy().copy()
      mom0 = inner.state[layer.weight]['momentum_buffer'].cpu().numpy().copy()
      sched.update(step)
      frac = sched.get_drop_fraction(step)
      assert sched.last_drop_fraction == frac and frac.dtype == np.float32
      want_mask, want_w = orc.tf2_generic_mask_update(m0, w0, np.abs(m0 * w0), np.abs(g), frac)
      assert np.array_equal(layer.mask.numpy(), want_mask)
      assert layer.weight.detach().cpu().numpy().tobytes() == want_w.tobytes()
      new = (want_mask == 1) & (m0 == 0)
      mom1 = inner.state[layer.weight]['momentum_buffer'].cpu().numpy()
      assert (mom1[new] == 0).all() and np.array_equal(mom1[~new], mom0[~new])       # reset_momentum (:156-162)
      n_updates += 1
  assert n_updates == 3 and gs.value == 12
  # prune: score_grow = None -> the layer keeps its top n_ones - int(n_ones * f) by |mask * w|, nothing grows
  m0, w0 = layer.mask.numpy().copy(), layer.weight.detach().cpu().numpy().copy()
  sched.prune(0.25)
  n_ones = int(m0.sum())
  n_keep = n_ones - int(np.float32(n_ones) * np.float32(0.25))
  order = np.argsort(-np.abs(m0 * w0).ravel(), kind='stable')
  want = np.zeros(m0.size, np.float32)
  want[order[:n_keep]] = 1
  assert np.array_equal(layer.mask.numpy().ravel(), want)
  assert layer.weight.detach().cpu().numpy().tobytes() == w0.tobytes()
