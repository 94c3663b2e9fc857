"""Dynamic-sparse-training optimizers (SET, RigL) over PyTorch + the CUDA mask update.

Host-side mirror of the reference's wrapper-optimizer interface
(google-research/rigl, rigl/sparse_optimizers_base.py):

  extract_number :45              SparseSETOptimizerBase :62
    compute_gradients :113          apply_gradients :118     cond_mask_update_op :152
    get_weights/get_masks/get_masked_weights :189-196 (abstract getter triple)
    is_mask_update_iter :198        get_drop_fraction :232   generic_mask_update :260
    _get_update_op :276             reset_momentum :345      get_grow_tensor :355
    _random_uniform/_random_normal :402-418
  SparseRigLOptimizerBase :421
    set_masked_grads :471  compute_gradients :478  apply_gradients :487
    generic_mask_update :523  get_grow_tensor :540  reset_momentum :555

Differences that are inherent to eager PyTorch (documented in DESIGN.md):
  * the wrapped optimizer is a `torch.optim.Optimizer`; "variables" are
    Parameters carrying `.name` ('<scope>/weights:0'), masks are
    `rigl_b200.masks.MaskVariable`;
  * `global_step` is a `GlobalStep` host counter (the TF variable analogue);
  * all masked layers are updated by ONE batched CUDA launch sequence
    (MaskUpdateEngine) instead of a per-layer op graph; `generic_mask_update` /
    `_get_update_op` remain available per layer and run the same kernels;
  * the drop-score noise is drawn from a counter-keyed torch generator seeded by
    a process-independent hash (crc32) of the variable name, the seed offset and
    the global step -- the reference seeds with Python's per-process salted
    `hash()` (base.py:270,534), which is not reproducible across processes.
"""
import math
import re
import zlib

import numpy as np
import torch

from . import _cabi
from .masks import MaskUpdateEngine


def extract_number(token):
  """'foo_0.5' -> 0.5, 'foo_4' -> 4.0, no numeric suffix -> 1.0."""
  found = re.search(r'.*_(\d*\.?\d*)$', token)
  return float(found.group(1)) if found else 1.


class GlobalStep(object):
  """Host-side training-step counter (stands in for the TF global_step variable)."""

  def __init__(self, value=0):
    self.value = int(value)

  def __int__(self):
    return self.value

  def increment(self):
    self.value += 1

  def state_dict(self):
    return {'value': self.value}

  def load_state_dict(self, sd):
    self.value = int(sd['value'])


_DEFAULT_GLOBAL_STEP = GlobalStep(0)


def get_or_create_global_step():
  return _DEFAULT_GLOBAL_STEP


def stable_hash(text):
  """Process-independent replacement for hash(str) used in RNG seeding."""
  return zlib.crc32(text.encode('utf-8')) & 0x7fffffff


def host_drop_fraction(anneal, initial_value, global_step, begin_step, end_step):
  """float32 drop fraction for `global_step` (before the is-update-iter gate).

  constant | cosine (tf.train.cosine_decay over decay_steps = end-begin, fed the
  RAW global step, alpha=0) | exponential_<k>.  Every op rounds to float32;
  cos/pow are taken in float64 on the float32 argument and rounded once, which
  fixes a machine-independent value (TF's Eigen cosf is not bit-portable).
  """
  f32 = np.float32
  init = f32(float(initial_value))
  if anneal == 'constant':
    return init
  if anneal == 'cosine':
    span = f32(end_step - begin_step)
    progress = f32(f32(min(f32(global_step), span)) / span)
    angle = f32(f32(math.pi) * progress)
    return f32(init * f32(f32(0.5) * f32(f32(1.0) + f32(math.cos(float(angle))))))
  if anneal.startswith('exponential'):
    k = f32(extract_number(anneal))
    frac_done = f32(f32(global_step - begin_step) / f32(end_step - begin_step))
    return f32(init * f32(math.pow(float(f32(f32(1.0) - frac_done)), float(k))))
  raise ValueError('drop_fraction_anneal: %s is not valid' % anneal)


def cross_replica_sum_(grads, enabled):
  """tpu_ops.cross_replica_sum of base.py:471-476: SUM every dense-gradient buffer over the replicas, once
  per backward -- a buffer carries `rigl_reduced` from the moment it is summed until the masked layer's
  next backward rewrites it (layers._MaskedConvFn.backward clears the flag)."""
  if not (enabled and torch.distributed.is_available() and torch.distributed.is_initialized()
          and torch.distributed.get_world_size() > 1):
    return
  for g in grads:
    if not getattr(g, 'rigl_reduced', False):
      torch.distributed.all_reduce(g, op=torch.distributed.ReduceOp.SUM)
      g.rigl_reduced = True


class SparseSETOptimizerBase(object):
  """Wraps a torch optimizer; periodically drops by magnitude and regrows at random."""

  # SET / Static redraw (or rebuild) their grow-score tensors at every update: the layer specs are rebuilt too.
  # Optimizers whose specs only point at persistent buffers (RigL: the dense gradients) set this to True.
  _specs_cacheable = False

  def __init__(self, optimizer, begin_step, end_step, frequency, drop_fraction=0.1,
               drop_fraction_anneal='constant', use_locking=False, grow_init='zeros',
               name='SparseSETOptimizer', use_stateless=True, stateless_seed_offset=0):
    self._optimizer = optimizer
    self._name = name
    self._use_locking = use_locking
    self._grow_init = grow_init
    self._drop_fraction_anneal = drop_fraction_anneal
    self._drop_fraction_initial_value = float(drop_fraction)
    self._begin_step = int(begin_step)
    self._end_step = int(end_step)
    self._frequency = int(frequency)
    self._frequency_val = int(frequency)
    self._use_stateless = use_stateless
    self._stateless_seed_offset = int(stateless_seed_offset)
    self._global_step = None
    self._last_update_step = None          # created lazily: -frequency (base.py:164-171)
    self._engine = MaskUpdateEngine()
    self._noise_bufs = {}
    self._score_bufs = {}
    self.drop_fraction = np.float32(0.)
    self.noise_std = 1e-5                   # default of generic_mask_update (base.py:260,523)
    self.last_update_was_mask_update = False
    # the drop-score noise is drawn inside the select kernels from a counter-based generator keyed like
    # `_generator` (seed offset + crc32('drop'), global step) and per layer by crc32(name + 'drop'); no noise
    # tensor exists.  RIGL_INKERNEL_NOISE=0 (or use_stateless=False): a torch.Generator draw into a flat buffer.
    import os
    self._inkernel_noise = bool(use_stateless) and os.environ.get('RIGL_INKERNEL_NOISE', '1') != '0'
    self._last_noise = None                 # (noise_std, seed) of the last in-kernel draw

  # ---- getter triple: supplied by a mixin (sparse_optimizers.PruningGetterTorchMixin)
  def get_weights(self):
    raise NotImplementedError

  def get_masks(self):
    raise NotImplementedError

  def get_masked_weights(self):
    raise NotImplementedError

  # ---- torch.optim.Optimizer passthroughs
  @property
  def param_groups(self):
    return self._optimizer.param_groups

  def zero_grad(self, set_to_none=False):
    self._optimizer.zero_grad(set_to_none=set_to_none)
    self._mark_dense_grads_stale()

  def _mark_dense_grads_stale(self):
    """The next backward overwrites (rather than accumulates into) the dense-grad buffers."""
    try:
      handles = self.get_masked_weights()
    except NotImplementedError:
      return
    for mw in handles:
      mw.fresh = False

  def get_slot_names(self):
    names = []
    for st in self._optimizer.state.values():
      for k, v in st.items():
        if torch.is_tensor(v) and v.dim() > 0 and k not in names:
          names.append(k)
    return names

  def get_slot(self, weights, name):
    st = self._optimizer.state.get(weights, {})
    v = st.get(name)
    return v if torch.is_tensor(v) and v.numel() == weights.numel() else None

  def state_dict(self):
    gs = self._global_step if self._global_step is not None else get_or_create_global_step()
    return {'optimizer': self._optimizer.state_dict(), 'global_step': int(gs),
            'last_mask_update_step': self._last_update_value()}

  def load_state_dict(self, sd):
    self._optimizer.load_state_dict(sd['optimizer'])
    gs = self._global_step if self._global_step is not None else get_or_create_global_step()
    gs.value = int(sd['global_step'])
    self._last_update_step = int(sd['last_mask_update_step'])

  # ---- gradients
  def _all_params(self):
    return [p for g in self._optimizer.param_groups for p in g['params']]

  def compute_gradients(self, loss, **kwargs):
    """Fresh backward of `loss`; returns [(grad, param)] like tf compute_gradients."""
    for p in self._all_params():
      if p.grad is not None:
        p.grad = None if kwargs.get('set_to_none', False) else p.grad.zero_()
    self._mark_dense_grads_stale()
    loss.backward()
    return [(p.grad, p) for p in self._all_params()]

  @staticmethod
  def _install_grads(grads_and_vars):
    for g, p in grads_and_vars or []:
      if g is not None and p.grad is not g:
        p.grad = g

  def _before_apply_gradients(self, grads_and_vars):
    return None

  def _inner_step(self, grads_and_vars, global_step):
    self._install_grads(grads_and_vars)
    self._optimizer.step()
    if global_step is not None:
      global_step.increment()

  def apply_gradients(self, grads_and_vars, global_step=None, name=None):
    """SET: optimizer step first, then (maybe) the mask update on the new step."""
    self._before_apply_gradients(grads_and_vars)
    self._inner_step(grads_and_vars, global_step)
    gs = global_step if global_step is not None else get_or_create_global_step()
    self._global_step = gs
    try:
      return self.cond_mask_update_op(gs, lambda: None)
    finally:
      # the gradients of this step are consumed: the next backward OVERWRITES the dense-grad buffers
      # whichever zero_grad() the caller uses (the reference recomputes them every step)
      self._mark_dense_grads_stale()

  def minimize(self, loss, global_step=None, **kwargs):
    return self.apply_gradients(self.compute_gradients(loss, **kwargs), global_step=global_step)

  def step(self, global_step=None):
    """torch-style entry: uses the .grad fields already populated by backward()."""
    if global_step is None:
      global_step = self._global_step if self._global_step is not None else get_or_create_global_step()
    return self.apply_gradients(None, global_step=global_step)

  # ---- schedule
  def _last_update_value(self):
    if self._last_update_step is None:
      self._last_update_step = -self._frequency_val
    return self._last_update_step

  def cond_mask_update_op(self, global_step, false_branch):
    last = self._last_update_value()
    if self.is_mask_update_iter(global_step, last):
      self.mask_update_op()
      self._last_update_step = int(global_step)
      self.last_update_was_mask_update = True
      return True
    self.last_update_was_mask_update = False
    false_branch()
    return False

  def mask_update_op(self):
    """Updates every (mask, weights) pair -- one batched launch sequence.  The drop-score
    noise of ALL layers is one keyed draw over a flat buffer (per-layer views), so the host
    cost does not scale with the layer count."""
    pairs = list(zip(self.get_masks(), self.get_weights()))
    if not pairs:
      return
    self._slot_names_cache = self.get_slot_names()
    try:
      noise = {} if self._inkernel_noise else self._batched_noise([w for _, w in pairs], self.noise_std)
      # The per-layer specs (54 dicts of tensor views for ResNet-50) only change when a buffer is reallocated or
      # an optimizer slot appears: they are cached under a key of the storages involved, and the same key lets
      # the engine skip re-validating its launch plan.
      key = (self._grow_init, float(self.noise_std), tuple(self._slot_names_cache)) + tuple(
          (id(m), w.data_ptr(), self._spec_cache_extra(w)) for m, w in pairs)
      cacheable = self._specs_cacheable and (self._grow_init == 'zeros' or self._grow_init.startswith('grad_'))
      cached = getattr(self, '_spec_cache', None)
      if not cacheable or cached is None or cached[0] != key:
        specs = [self._layer_spec(m, w, self.noise_std, noise=noise.get(w.name)) for m, w in pairs]
        self._spec_cache = (key, specs)
      self._run_update(self._spec_cache[1], plan_key=key if cacheable else None, noise_std=self.noise_std)
    finally:
      self._slot_names_cache = None

  def _spec_cache_extra(self, weights):
    """Whatever else a cached layer spec points at (subclasses: the dense-gradient / EMA buffer)."""
    st = self._optimizer.state.get(weights) if hasattr(self._optimizer, 'state') else None
    return tuple(v.data_ptr() for v in st.values() if torch.is_tensor(v) and v.dim() > 0) if st else ()

  def _batched_noise(self, weights, noise_std):
    if not noise_std:
      return {}
    sizes = [(w.numel() + 127) // 128 * 128 for w in weights]
    total = sum(sizes)
    flat = getattr(self, '_noise_flat', None)
    if flat is None or flat.numel() != total or flat.device != weights[0].device:
      flat = torch.empty(total, dtype=torch.float32, device=weights[0].device)
      self._noise_flat = flat
    self._random_normal(flat.shape, stddev=noise_std, dtype=torch.float32, seed=stable_hash('drop'),
                        out=flat, device=flat.device)
    views, off = {}, 0
    for w, sz in zip(weights, sizes):
      views[w.name] = flat[off:off + w.numel()]
      off += sz
    self._noise_bufs = dict(views)
    return views

  def is_mask_update_iter(self, global_step, last_update_step):
    gs = int(global_step)
    in_range = gs >= self._begin_step and (gs <= self._end_step or self._end_step < 0)
    due = int(last_update_step) + self._frequency <= gs
    is_update = bool(in_range and due)
    self.drop_fraction = self.get_drop_fraction(global_step, is_update)
    return is_update

  def get_drop_fraction(self, global_step, is_mask_update_iter_op):
    frac = host_drop_fraction(self._drop_fraction_anneal, self._drop_fraction_initial_value,
                              int(global_step), self._begin_step, self._end_step)
    return frac if is_mask_update_iter_op else np.float32(0.)

  # ---- per-layer pieces
  def _noise_for(self, weights, noise_std):
    if not noise_std:
      return None
    buf = self._noise_bufs.get(weights.name)
    if buf is None or buf.numel() != weights.numel():
      buf = torch.empty(weights.numel(), dtype=torch.float32, device=weights.device)
      self._noise_bufs[weights.name] = buf
    self._random_normal(buf.shape, stddev=noise_std, dtype=torch.float32,
                        seed=stable_hash(weights.name + 'drop'), out=buf, device=weights.device)
    return buf

  def _score_grow_for(self, mask, weights):
    """SET: a uniform draw per position (base.py:271-273)."""
    buf = self._score_bufs.get(weights.name)
    if buf is None or buf.numel() != weights.numel():
      buf = torch.empty(weights.numel(), dtype=torch.float32, device=weights.device)
      self._score_bufs[weights.name] = buf
    self._random_uniform(buf.shape, seed=stable_hash(weights.name + 'grow'), out=buf,
                         device=weights.device)
    return buf

  def _slots_of(self, weights):
    names = getattr(self, '_slot_names_cache', None)
    slots = [self.get_slot(weights, n) for n in (names if names is not None else self.get_slot_names())]
    return [s.view(-1) for s in slots if s is not None]

  def _grow_spec(self, weights, method):
    """(grow_mode, divisor, grow_values) for the kernel."""
    if not isinstance(method, str):
      raise ValueError('Grow-Init: %s is not a string' % method)
    if method == 'zeros':
      return _cabi.GROW_ZEROS, 1.0, None
    return _cabi.GROW_TENSOR, 1.0, self.get_grow_tensor(weights, method).contiguous().view(-1)

  def _acc_scale(self):
    return 0.0

  def _grad_for(self, weights):
    """The gradient the grad_* grow inits / slot reset read when the grow score is NOT it (explicit scores)."""
    return None

  def _layer_spec(self, mask, weights, noise_std, score_drop=None, score_grow=None,
                  reinit_when_same=False, noise=None, signed_grow=False):
    mode, div, grow_values = self._grow_spec(weights, self._grow_init)
    if score_grow is None:
      score_grow = self._score_grow_for(mask, weights)
    if noise is None and score_drop is None and not self._inkernel_noise:
      noise = self._noise_for(weights, noise_std)
    return dict(mask=mask, weights=weights.data.view(-1), score_grow=score_grow.contiguous().view(-1),
                noise=None if score_drop is not None else noise,
                score_drop=None if score_drop is None else score_drop.contiguous().view(-1),
                slots=self._slots_of(weights), grow_values=grow_values, grow_mode=mode,
                grow_divisor=div, reinit_when_same=reinit_when_same,
                noise_key=stable_hash(weights.name + 'drop'),
                flags=_cabi.LAYER_GROW_SCORE_SIGNED if signed_grow else 0,
                grad=self._grad_for(weights) if signed_grow else None)

  def _noise_seed(self):
    gs = int(self._global_step) if self._global_step is not None else 0
    return (((self._stateless_seed_offset + stable_hash('drop')) & 0x7fffffff) << 32) | (gs & 0xffffffff)

  def last_update_noise(self, weights):
    """The noise the last update added in-kernel to the drop scores of `weights` (float32, flat), re-materialised
    for inspection / the CPU oracle; None when that update drew none."""
    if self._last_noise is None or not self._last_noise[0]:
      return None
    from .masks import noise_fill
    std, seed = self._last_noise
    return noise_fill(weights.numel(), stable_hash(weights.name + 'drop'), std, seed, weights.device)

  def _run_update(self, specs, plan_key=None, noise_std=0.0):
    first = specs[0]
    mode, div, reinit = first['grow_mode'], first['grow_divisor'], first['reinit_when_same']
    if plan_key is None or plan_key != getattr(self, '_checked_modes_key', None):
      if any((s['grow_mode'], s['grow_divisor'], s['reinit_when_same']) != (mode, div, reinit) for s in specs):
        raise ValueError('all layers of one update must share grow_init / reinit mode')
      self._checked_modes_key = plan_key
    std = float(noise_std) if (self._inkernel_noise and noise_std) else 0.0
    seed = self._noise_seed() if std else 0
    self._last_noise = (std, seed)
    self._engine.run(specs, np.float32(self.drop_fraction), grow_mode=mode, grow_divisor=div,
                     acc_scale=self._acc_scale(), reinit_when_same=reinit, plan_key=plan_key,
                     noise_std=std, noise_seed=seed)

  def generic_mask_update(self, mask, weights, noise_std=1e-5):
    """Drop/grow of ONE layer with the optimizer's scores (uses self.drop_fraction)."""
    self._run_update([self._layer_spec(mask, weights, noise_std)], noise_std=noise_std)
    return mask

  def _get_update_op(self, score_drop, score_grow, mask, weights, reinit_when_same=False):
    """Prune + grow one layer from explicit score tensors (all of `mask.shape`), both used VERBATIM
    (signed) like base.py:276-343: top-k of score_drop keeps, top-k of score_grow among the rest grows.
    The grad_* grow inits and the RigL slot reset read the stored dense gradient, not score_grow."""
    self._run_update([self._layer_spec(mask, weights, 0., score_drop=score_drop.float(),
                                       score_grow=score_grow.float(),
                                       reinit_when_same=reinit_when_same, signed_grow=True)])
    return mask

  def reset_momentum(self, weights, new_connections):
    """Zeroes every optimizer slot of `weights` where `new_connections` (bool tensor)."""
    for s_name in self.get_slot_names():
      slot = self.get_slot(weights, s_name)
      if slot is not None:
        slot.masked_fill_(new_connections.view(slot.shape), 0.)

  def get_grow_tensor(self, weights, method):
    """Initial values for grown connections: 'zeros', 'initial_dist[_d]',
    'random_normal[_d]', 'random_uniform[_d]'; ValueError otherwise."""
    if not isinstance(method, str):
      raise ValueError('Grow-Init: %s is not a string' % method)
    w = weights.data if isinstance(weights, torch.nn.Parameter) else weights
    name = getattr(weights, 'name', None) or 'weights'
    if method == 'zeros':
      return torch.zeros_like(w)
    if method.startswith('initial_dist'):
      init = getattr(weights, 'initial_value', None)
      if init is None:
        raise ValueError('Grow-Init: %s needs weights.initial_value' % method)
      perm = torch.randperm(init.numel(), device=init.device)
      return (init.reshape(-1)[perm].reshape(init.shape) / extract_number(method)).to(w.dtype)
    if method.startswith('random_normal'):
      std = float(w.double().std(unbiased=False))
      return self._random_normal(w.shape, stddev=std, dtype=w.dtype,
                                 seed=stable_hash(name + 'grow_init_n'),
                                 device=w.device) / extract_number(method)
    if method.startswith('random_uniform'):
      mean = float(w.double().abs().mean())
      return self._random_uniform(w.shape, minval=-mean, maxval=mean, dtype=w.dtype,
                                  seed=stable_hash(name + 'grow_init_u'),
                                  device=w.device) / extract_number(method)
    raise ValueError('Grow-Init: %s is not a valid option.' % method)

  # ---- RNG (stateless: keyed by (seed_offset + seed, global_step), replica-identical)
  def _generator(self, seed, device):
    if not self._use_stateless:
      return None
    gs = int(self._global_step) if self._global_step is not None else 0
    key = (((self._stateless_seed_offset + int(seed)) & 0x7fffffff) << 32) | (gs & 0xffffffff)
    gen = torch.Generator(device=device)
    gen.manual_seed(key)
    return gen

  def _random_uniform(self, shape, minval=0., maxval=1., dtype=torch.float32, seed=0, out=None,
                      device='cpu'):
    if out is None:
      out = torch.empty(tuple(shape), dtype=dtype, device=device)
    out.uniform_(minval, maxval, generator=self._generator(seed, out.device))
    return out

  def _random_normal(self, shape, stddev=1., dtype=torch.float32, seed=0, out=None, device='cpu'):
    if out is None:
      out = torch.empty(tuple(shape), dtype=dtype, device=device)
    out.normal_(0., float(stddev), generator=self._generator(seed, out.device))
    return out


class SparseRigLOptimizerBase(SparseSETOptimizerBase):
  """Grows where the DENSE gradient magnitude is largest (Evci et al., RigL)."""

  _specs_cacheable = True

  def __init__(self, optimizer, begin_step, end_step, frequency, drop_fraction=0.1,
               drop_fraction_anneal='constant', use_locking=False, grow_init='zeros',
               initial_acc_scale=0., use_tpu=False, name='SparseRigLOptimizer',
               stateless_seed_offset=0):
    super(SparseRigLOptimizerBase, self).__init__(
        optimizer, begin_step, end_step, frequency, drop_fraction=drop_fraction,
        drop_fraction_anneal=drop_fraction_anneal, grow_init=grow_init, use_locking=use_locking,
        name='SparseRigLOptimizer', stateless_seed_offset=stateless_seed_offset)
    self._initial_acc_scale = float(initial_acc_scale)
    self._use_tpu = use_tpu       # "aggregate the dense grads across replicas"
    self._masked_grads = []
    self._weight2masked_grads = {}

  def set_masked_grads(self, grads, weights):
    """Stores dL/d(mask*w) per weight name; cross-replica SUM when `use_tpu`."""
    cross_replica_sum_(grads, self._use_tpu)
    self._masked_grads = list(grads)
    self._weight2masked_grads = {w.name: g for w, g in zip(weights, grads)}

  def compute_gradients(self, loss, **kwargs):
    grads_and_vars = super(SparseRigLOptimizerBase, self).compute_gradients(loss, **kwargs)
    self.collect_masked_grads()
    return grads_and_vars

  def collect_masked_grads(self):
    """Picks up the dense gradients the masked layers' backward left behind."""
    dense = [mw.dense_grad for mw in self.get_masked_weights()]
    self.set_masked_grads(dense, self.get_weights())

  def apply_gradients(self, grads_and_vars, global_step=None, name=None):
    """RigL: EITHER a mask update (no optimizer step, step counter frozen) OR a step."""
    self._before_apply_gradients(grads_and_vars)
    gs = global_step if global_step is not None else get_or_create_global_step()
    self._global_step = gs
    # every step (the reference's compute_gradients runs cross_replica_sum every step, base.py:471-485);
    # buffers a DataParallel wrapper or compute_gradients already reduced carry `rigl_reduced` and are
    # not summed twice
    self.collect_masked_grads()
    try:
      return self.cond_mask_update_op(gs, lambda: self._inner_step(grads_and_vars, global_step))
    finally:
      self._mark_dense_grads_stale()

  def _score_grow_for(self, mask, weights):
    return self._weight2masked_grads[weights.name]       # |.| is taken in the kernel

  def _spec_cache_extra(self, weights):
    g = self._weight2masked_grads.get(weights.name)
    return super(SparseRigLOptimizerBase, self)._spec_cache_extra(weights) + ((g.data_ptr(),) if g is not None else ())

  def _grad_for(self, weights):
    g = self._weight2masked_grads.get(weights.name)
    return None if g is None else g.contiguous().view(-1)

  def _acc_scale(self):
    return self._initial_acc_scale

  def _grow_spec(self, weights, method):
    if isinstance(method, str) and method.startswith('grad_scale'):
      return _cabi.GROW_GRAD_SCALE, extract_number(method), None
    if isinstance(method, str) and method.startswith('grad_sign'):
      return _cabi.GROW_GRAD_SIGN, extract_number(method), None
    return super(SparseRigLOptimizerBase, self)._grow_spec(weights, method)

  def get_grow_tensor(self, weights, method):
    if isinstance(method, str) and method.startswith('grad_scale'):
      return self._weight2masked_grads[weights.name].view(weights.shape) / extract_number(method)
    if isinstance(method, str) and method.startswith('grad_sign'):
      return torch.sign(self._weight2masked_grads[weights.name]).view(weights.shape) / \
          extract_number(method)
    return super(SparseRigLOptimizerBase, self).get_grow_tensor(weights, method)

  def reset_momentum(self, weights, new_connections):
    """slot <- dense_grad * initial_acc_scale where `new_connections`."""
    acc = self._weight2masked_grads[weights.name].view(weights.shape) * self._initial_acc_scale
    for s_name in self.get_slot_names():
      slot = self.get_slot(weights, s_name)
      if slot is not None:
        slot.copy_(torch.where(new_connections.view(slot.shape), acc.view(slot.shape), slot))
