"""Oracle restatements of SparseMomentum / SparseSnip / SparseDNW (rigl/sparse_optimizers.py:126-480)
against the expectations of the reference's own tests (rigl/sparse_optimizers_test.py:247-590),
re-expressed on plain arrays (the reference builds a 1-layer masked_fully_connected graph whose
gradients are known in closed form)."""
import numpy as np
import pytest

from oracle import rigl_oracle as orc


@pytest.mark.parametrize('n_inp,n_out,momentum', [(3, 4, 0.5), (5, 2, 0.), (2, 5, 1.)])
def test_momentum_ema_trajectory(n_inp, n_out, momentum):
  """sparse_optimizers_test.py:276-295 (testMomentumUpdate): x = ones, y_j scaled by j, so the dense
  gradient of every weight in column j is j."""
  g = np.broadcast_to(np.arange(n_out, dtype=np.float32), (n_inp, n_out))
  ema = np.zeros((n_inp, n_out), np.float32)
  want = np.zeros((n_inp, n_out))
  for _ in range(6):
    assert np.array_equal(ema, want.astype(np.float32))
    ema = orc.momentum_ema_update(ema, g, momentum)
    want = want * momentum + (1 - momentum) * np.arange(n_out)
    assert np.array_equal(ema, want.astype(np.float32))


def test_momentum_update_grows_where_the_ema_is_largest():
  rng = np.random.RandomState(0)
  mask = orc.get_mask_random_numpy((6, 7), 0.5, rng).astype(np.float32)
  w = rng.standard_normal((6, 7)).astype(np.float32)
  ema = rng.standard_normal((6, 7)).astype(np.float32)
  r = orc.momentum_mask_update(mask, w, ema, 0.5)
  assert r['mask'].sum() == mask.sum()                                   # sparsity preserved
  grown = (r['mask2'] == 1)
  cand = (r['mask1'] == 0)                                               # everything not kept competes
  assert np.abs(ema)[grown].min() >= np.abs(ema)[cand & ~grown].max()
  assert (r['weights'][r['new_connections']] == 0).all()                 # zeros grow-init
  # identical to the RigL update when the EMA equals the instantaneous dense gradient
  rr = orc.rigl_mask_update(mask, w, ema, 0.5)
  assert np.array_equal(r['mask'], rr['mask'])


@pytest.mark.parametrize('n_inp,n_out,sparsity', [(3, 4, 0.5), (5, 3, 0.8), (8, 5, 0.8)])
def test_snip_sparsity_and_scores(n_inp, n_out, sparsity):
  """testSnipSparsity + testGradientUsed (sparse_optimizers_test.py:407-438)."""
  rng = np.random.RandomState(n_inp * 10 + n_out)
  inp = np.arange(1, n_inp + 1)
  scale = rng.uniform(size=(n_out,)) - 0.5
  grads = np.outer(inp, scale).astype(np.float32)                        # closed-form dL/dW of the test graph
  w = rng.standard_normal((n_inp, n_out)).astype(np.float32)
  m = orc.snip_mask(grads, w, sparsity)
  assert m.size - m.sum() == orc.get_n_zeros(m.size, sparsity)
  scores = np.abs(grads * w)
  assert scores[m == 0].max() <= scores[m == 1].min()


def test_snip_ties_keep_the_lower_flat_index_and_control_flow():
  g = np.ones((2, 4), np.float32)
  w = np.array([[1, 2, 2, 1], [2, 1, 2, 2]], np.float32)                 # five scores of 2, three of 1
  m = orc.snip_mask(g, w, 0.5)                                           # keep 4 of 8
  assert np.array_equal(m.ravel(), [0, 1, 1, 0, 1, 0, 1, 0])             # the LAST tied 2 (flat 7) is dropped
  sim = orc.SnipSim()
  assert sim.is_snip_iter(0)
  sim.is_snipped = True
  assert not sim.is_snip_iter(0) and not sim.is_snip_iter(3)            # sparse_optimizers_test.py:451-468


@pytest.mark.parametrize('n_inp,n_out,sparsity', [(3, 4, 0.5), (5, 3, 0.8), (8, 5, 0.8)])
def test_dnw_keeps_the_largest_magnitudes(n_inp, n_out, sparsity):
  """testDNWSparsity + testWeightsUsed (sparse_optimizers_test.py:515-546)."""
  rng = np.random.RandomState(n_inp + 100 * n_out)
  w = rng.standard_normal((n_inp, n_out)).astype(np.float32)
  m = orc.dnw_mask(w, sparsity)
  assert m.size - m.sum() == orc.get_n_zeros(m.size, sparsity)
  assert np.abs(w)[m == 0].max() <= np.abs(w)[m == 1].min()
  # the mask follows the weights: after a step that shrinks a kept weight to zero it is dropped
  w2 = w.copy()
  w2.ravel()[np.argmax(np.abs(w))] = 0.
  m2 = orc.dnw_mask(w2, sparsity)
  assert m2.ravel()[np.argmax(np.abs(w))] == 0 and m2.sum() == m.sum()


@pytest.mark.parametrize('n_inp,n_out,momentum', [(3, 4, 0.5), (5, 2, 0.), (2, 5, 1.), (4, 4, 0.9)])
def test_momentum_optimizer_host_logic_matches_oracle(n_inp, n_out, momentum):
  """The product class' EMA bookkeeping (rigl_b200.sparse_optimizers.SparseMomentumOptimizer) on CPU
  tensors: same trajectory as the oracle / the reference's testMomentumUpdate; the grow score handed
  to the select kernels is that EMA."""
  import torch
  from rigl_b200.sparse_optimizers import SparseMomentumOptimizer

  class W(object):
    name = 'layer/weights:0'
  w = torch.nn.Parameter(torch.zeros(n_inp, n_out))
  opt = SparseMomentumOptimizer(torch.optim.SGD([w], lr=0.1), 1, 4, 2, drop_fraction=0.5, momentum=momentum)
  g = torch.arange(n_out, dtype=torch.float32).repeat(n_inp, 1).contiguous().view(-1)
  opt.get_weights = lambda: [W]
  opt.get_masked_weights = lambda: [__import__('types').SimpleNamespace(dense_grad=g, fresh=True)]
  opt.set_masked_grads([g], [W])
  ema = np.zeros(n_inp * n_out, np.float32)
  for _ in range(6):
    opt._before_apply_gradients(None)
    ema = orc.momentum_ema_update(ema, g.numpy(), momentum)
    assert np.array_equal(opt.ema_average(W).numpy(), ema)
    assert opt._score_grow_for(None, W) is opt.ema_average(W)
