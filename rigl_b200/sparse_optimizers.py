"""Concrete sparse optimizers bound to the PyTorch masked-layer registry.

Mirror of rigl/sparse_optimizers.py:46-123 (PruningGetterTf1Mixin,
SparseSETOptimizer, SparseRigLOptimizer, SparseStaticOptimizer) and of the
momentum / SNIP / DNW optimizers (:126-480), which reuse the same batched
select kernels with different scores.
"""
from . import pruning
from . import sparse_optimizers_base as sparse_opt_base


class PruningGetterTorchMixin(object):
  """Variable retrieval from a MaskedLayerRegistry (default: the global one).

  `bind(model_or_registry)` scopes the optimizer to one model's layers."""

  _registry = None

  def bind(self, model_or_registry):
    if isinstance(model_or_registry, pruning.MaskedLayerRegistry):
      self._registry = model_or_registry
    else:
      self._registry = pruning.MaskedLayerRegistry.from_module(model_or_registry)
    return self

  def _reg(self):
    return self._registry if self._registry is not None else pruning.default_registry()

  def get_weights(self):
    return self._reg().get_weights()

  def get_masks(self):
    return self._reg().get_masks()

  def get_masked_weights(self):
    return self._reg().get_masked_weights()


class SparseSETOptimizer(PruningGetterTorchMixin, sparse_opt_base.SparseSETOptimizerBase):
  pass


class SparseRigLOptimizer(PruningGetterTorchMixin, sparse_opt_base.SparseRigLOptimizerBase):
  pass


class SparseStaticOptimizer(SparseSETOptimizer):
  """Keeps the connectivity fixed but re-initialises the weakest connections:
  grow score = the current mask, reinit_when_same=True (reference :109-123)."""

  def __init__(self, optimizer, begin_step, end_step, frequency, drop_fraction=0.1,
               drop_fraction_anneal='constant', use_locking=False, grow_init='zeros',
               name='SparseStaticOptimizer', stateless_seed_offset=0):
    super(SparseStaticOptimizer, self).__init__(
        optimizer, begin_step, end_step, frequency, drop_fraction=drop_fraction,
        drop_fraction_anneal=drop_fraction_anneal, grow_init=grow_init,
        use_locking=use_locking, name=name, stateless_seed_offset=stateless_seed_offset)

  def _score_grow_for(self, mask, weights):
    return mask.to_dense().view(-1)

  def _layer_spec(self, mask, weights, noise_std, score_drop=None, score_grow=None,
                  reinit_when_same=True, noise=None, signed_grow=False):
    return super(SparseStaticOptimizer, self)._layer_spec(
        mask, weights, noise_std, score_drop=score_drop, score_grow=score_grow,
        reinit_when_same=True, noise=noise, signed_grow=signed_grow)


class SparseMomentumOptimizer(SparseSETOptimizer):
  """Grows where the exponential moving average of the DENSE gradient is largest; no
  redistribution of sparsity (reference sparse_optimizers.py:126-214, after Dettmers & Zettlemoyer).

  Same drop rule and same batched select kernels as SET / RigL; only the grow score differs.  The
  EMA follows tf.train.ExponentialMovingAverage on a Tensor: zero-initialised, updated BEFORE every
  weight update by shadow -= (shadow - grad) * (1 - momentum)."""

  def __init__(self, optimizer, begin_step, end_step, frequency, drop_fraction=0.1,
               drop_fraction_anneal='constant', use_locking=False, grow_init='zeros', momentum=0.9,
               use_tpu=False, name='SparseMomentumOptimizer', stateless_seed_offset=0):
    super(SparseMomentumOptimizer, self).__init__(
        optimizer, begin_step, end_step, frequency, drop_fraction=drop_fraction,
        drop_fraction_anneal=drop_fraction_anneal, grow_init=grow_init, use_locking=use_locking,
        name='SparseMomentumOptimizer', stateless_seed_offset=stateless_seed_offset)
    self._momentum = float(momentum)
    self._use_tpu = use_tpu
    self._masked_grads = []
    self._weight2masked_grads = {}
    self._ema = {}

  def set_masked_grads(self, grads, weights):
    """reference :176-181 (cross-replica SUM of the dense grads when `use_tpu`)."""
    sparse_opt_base.cross_replica_sum_(grads, self._use_tpu)
    self._masked_grads = list(grads)
    self._weight2masked_grads = {w.name: g for w, g in zip(weights, grads)}

  def compute_gradients(self, loss, **kwargs):
    grads_and_vars = super(SparseMomentumOptimizer, self).compute_gradients(loss, **kwargs)
    self.collect_masked_grads()
    return grads_and_vars

  def collect_masked_grads(self):
    dense = [mw.dense_grad for mw in self.get_masked_weights()]
    self.set_masked_grads(dense, self.get_weights())

  def _before_apply_gradients(self, grads_and_vars):
    """Updates the EMA before the weights move (reference :195-197)."""
    import torch
    self.collect_masked_grads()          # every step: the buffers may have been rebound / rewritten
    c = 1.0 - self._momentum
    for w, g in zip(self.get_weights(), self._masked_grads):
      ema = self._ema.get(w.name)
      if ema is None:
        ema = self._ema[w.name] = torch.zeros_like(g)
      ema.sub_((ema - g).mul_(c))

  def ema_average(self, weights):
    """`self._ema_grads.average(masked_grad)` of the reference."""
    return self._ema[weights.name]

  def _score_grow_for(self, mask, weights):
    return self._ema[weights.name]              # |.| is taken in the kernel
