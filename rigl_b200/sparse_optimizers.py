"""Concrete sparse optimizers bound to the PyTorch masked-layer registry.

Mirror of rigl/sparse_optimizers.py:46-123 (PruningGetterTf1Mixin,
SparseSETOptimizer, SparseRigLOptimizer, SparseStaticOptimizer).  The momentum /
SNIP / DNW optimizers of the reference (:126-480) reuse the same select
primitive and are listed as "next" in DESIGN.md.
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
                  reinit_when_same=True, noise=None):
    return super(SparseStaticOptimizer, self)._layer_spec(
        mask, weights, noise_std, score_drop=score_drop, score_grow=score_grow,
        reinit_when_same=True, noise=noise)
