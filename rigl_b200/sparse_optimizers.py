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


class _TopKMaskAssigner(object):
  """`mask <- top-(n - get_n_zeros(n, sparsity)) of a per-position score` for every layer at once -- the
  `mask_fn` SNIP and DNW hand to sparse_utils.get_mask_init_fn (reference :286-316, :436-465: a full top_k sort
  per layer, ties -> lower index) -- on the batched select kernels: the drop phase alone
  (RIGL_LAYER_DROP_ONLY), ranking every position (RIGL_LAYER_ALL_ACTIVE), n_prune = get_n_zeros."""

  def _init_topk(self, default_sparsity, mask_init_method, custom_sparsity_map):
    from .masks import MaskUpdateEngine
    self._default_sparsity = default_sparsity
    self._mask_init_method = mask_init_method
    self._custom_sparsity_map = custom_sparsity_map or {}
    self._topk_engine = MaskUpdateEngine()
    self._sparsities = None

  def layer_sparsities(self):
    from . import sparse_utils
    if self._sparsities is None:      # depends on the (static) mask shapes only
      self._sparsities = sparse_utils.get_sparsities(self.get_masks(), self._mask_init_method,
                                                     self._default_sparsity, self._custom_sparsity_map)
    return self._sparsities

  def _assign_topk(self, scores):
    """scores: per layer an explicit float32 score tensor, or None to rank |weights| in-kernel."""
    from . import _cabi, sparse_utils
    sp = self.layer_sparsities()
    specs = []
    for m, w, sc in zip(self.get_masks(), self.get_weights(), scores):
      flat_w = w.data.view(-1)
      spec = dict(mask=m, weights=flat_w, score_grow=flat_w,        # (score_grow is not read: nothing grows)
                  n_prune=int(sparse_utils.get_n_zeros(m.size, sp[m.name])),
                  flags=_cabi.LAYER_DROP_ONLY | _cabi.LAYER_ALL_ACTIVE)
      if sc is not None:
        spec['score_drop'] = sc.contiguous().view(-1)
      specs.append(spec)
    if specs:
      self._topk_engine.run(specs, 0.0)


class SparseSnipOptimizer(PruningGetterTorchMixin, _TopKMaskAssigner):
  """SNIP (Lee et al.): at global step 0 -- instead of an optimizer step -- every mask becomes the top
  (1 - sparsity) fraction of |grad * weight|; afterwards a plain wrapper (reference :217-337)."""

  def __init__(self, optimizer, default_sparsity, mask_init_method, custom_sparsity_map=None,
               use_locking=False, use_tpu=False, name='SparseSnipOptimizer'):
    self._optimizer = optimizer
    self._use_tpu = use_tpu
    self._name = name
    self._init_topk(default_sparsity, mask_init_method, custom_sparsity_map)
    self.is_snipped = False                    # the reference's non-trainable `is_snipped` variable

  @property
  def param_groups(self):
    return self._optimizer.param_groups

  def zero_grad(self, set_to_none=False):
    self._optimizer.zero_grad(set_to_none=set_to_none)
    for mw in self.get_masked_weights():
      mw.fresh = False

  def compute_gradients(self, loss, **kwargs):
    self.zero_grad(set_to_none=kwargs.get('set_to_none', False))
    loss.backward()
    return [(p.grad, p) for g in self._optimizer.param_groups for p in g['params']]

  def apply_gradients(self, grads_and_vars, global_step=None, name=None):
    import torch
    gs = global_step if global_step is not None else sparse_opt_base.get_or_create_global_step()
    if int(gs) == 0 and not self.is_snipped:
      by_var = {id(v): g for g, v in (grads_and_vars or [])}
      scores = []
      for w in self.get_weights():
        g = by_var.get(id(w), w.grad)
        if g is None:
          raise ValueError('SNIP needs a gradient for %s' % w.name)
        g = g.detach().to(torch.float32)
        if self._use_tpu and torch.distributed.is_available() and torch.distributed.is_initialized() \
            and torch.distributed.get_world_size() > 1:
          g = g.clone()
          torch.distributed.all_reduce(g, op=torch.distributed.ReduceOp.SUM)      # tpu_ops.cross_replica_sum
        scores.append((g * w.data).abs_())
      self._assign_topk(scores)
      self.is_snipped = True
      return True
    sparse_opt_base.SparseSETOptimizerBase._install_grads(grads_and_vars)
    self._optimizer.step()
    if global_step is not None:
      global_step.increment()
    return False

  def minimize(self, loss, global_step=None, **kwargs):
    return self.apply_gradients(self.compute_gradients(loss, **kwargs), global_step=global_step)

  def state_dict(self):
    return {'optimizer': self._optimizer.state_dict(), 'is_snipped': bool(self.is_snipped)}

  def load_state_dict(self, sd):
    self._optimizer.load_state_dict(sd['optimizer'])
    self.is_snipped = bool(sd['is_snipped'])


class SparseDNWOptimizer(PruningGetterTorchMixin, _TopKMaskAssigner):
  """Discovering Neural Wirings (Wortsman et al.): the weights are updated with the DENSE gradient
  (dL/d(mask*w) applied to w) and after EVERY optimizer step each mask becomes the top (1 - sparsity) fraction
  of |w| (reference :340-480)."""

  def __init__(self, optimizer, default_sparsity, mask_init_method, custom_sparsity_map=None, use_tpu=False,
               use_locking=False, name='SparseDNWOptimizer'):
    self._optimizer = optimizer
    self._use_tpu = use_tpu
    self._name = name
    self._init_topk(default_sparsity, mask_init_method, custom_sparsity_map)

  @property
  def param_groups(self):
    return self._optimizer.param_groups

  def zero_grad(self, set_to_none=False):
    self._optimizer.zero_grad(set_to_none=set_to_none)
    for mw in self.get_masked_weights():
      mw.fresh = False

  def compute_gradients(self, loss, var_list=None, **kwargs):
    """Gradients with every masked variable replaced by its masked_weights tensor (replace_with_masked_weights,
    :388-396), i.e. the DENSE gradient, handed back under the weight variable (:398-408)."""
    self.zero_grad(set_to_none=kwargs.get('set_to_none', False))
    loss.backward()
    for w, mw in zip(self.get_weights(), self.get_masked_weights()):
      dense = mw.dense_grad.view(w.shape)
      if w.grad is None:
        w.grad = dense.clone()
      else:
        w.grad.copy_(dense)
    params = var_list if var_list is not None else [p for g in self._optimizer.param_groups for p in g['params']]
    return [(p.grad, p) for p in params]

  def apply_gradients(self, grads_and_vars, global_step=None, name=None):
    sparse_opt_base.SparseSETOptimizerBase._install_grads(grads_and_vars)
    self._optimizer.step()
    if global_step is not None:
      global_step.increment()
    self._assign_topk([None] * len(self.get_masks()))       # |w| is ranked in-kernel
    for mw in self.get_masked_weights():
      mw.fresh = False
    return True

  def minimize(self, loss, global_step=None, **kwargs):
    return self.apply_gradients(self.compute_gradients(loss, **kwargs), global_step=global_step)

  def state_dict(self):
    return {'optimizer': self._optimizer.state_dict()}

  def load_state_dict(self, sd):
    self._optimizer.load_state_dict(sd['optimizer'])
