"""TF2-style mask-update schedules (google-research/rigl, rigl/rigl_tf2/mask_updaters.py:251-344).

Host-side control flow only: WHEN to update the masks and with WHICH drop fraction; the update
itself is whatever `mask_updater` does (any object with `update_masks(drop_fraction)` and, for
`prune`, `prune_masks(fraction)` -- e.g. a thin adapter over `SparseSETOptimizerBase.mask_update_op`).
Drop fractions are float32 like the TF tensors they replace.
"""
import math

import numpy as np

F32 = np.float32


class UpdateSchedule(object):
  """mask_updaters.py:251-303.  last_update_step < 0: no last step; == 0: never update."""

  def __init__(self, mask_updater, init_drop_fraction, update_freq, last_update_step):
    self._mask_updater = mask_updater
    self.update_freq = update_freq
    self.last_update_step = last_update_step
    self.init_drop_fraction = F32(init_drop_fraction)
    self.last_drop_fraction = 0

  def get_drop_fraction(self, step):
    raise NotImplementedError

  def is_update_iter(self, step):
    """True if `step` is a valid mask update step (:270-283)."""
    if step < 0:
      raise ValueError('step must be >= 0, got %r' % (step,))       # tf.debugging.Assert(step >= 0)
    if self.last_update_step < 0:
      is_valid_step = True
    elif self.last_update_step == 0:
      is_valid_step = False
    else:
      is_valid_step = step <= self.last_update_step
    return bool(is_valid_step and step % self.update_freq == 0)

  def update(self, step, check_update_iter=True):
    if check_update_iter and not self.is_update_iter(step):
      raise ValueError('step %r is not a mask-update step' % (step,))
    self.last_drop_fraction = self.get_drop_fraction(step)
    if self.last_drop_fraction > 0.:
      self._mask_updater.update_masks(self.last_drop_fraction)

  def prune(self, prune_fraction):
    self.last_drop_fraction = prune_fraction
    self._mask_updater.prune_masks(self.last_drop_fraction)

  def set_validation_data(self, val_x, val_y):
    self._mask_updater.set_validation_data(val_x, val_y)


class ConstantUpdateSchedule(UpdateSchedule):
  """Updates a constant fraction of connections (:306-310)."""

  def get_drop_fraction(self, step):
    return self.init_drop_fraction


class CosineUpdateSchedule(UpdateSchedule):
  """tf.keras.experimental.CosineDecay(init, last_update_step, alpha=0) (:313-326):
  init * 0.5 * (1 + cos(pi * min(step, decay_steps) / decay_steps)), evaluated in float32."""

  def get_drop_fraction(self, step):
    decay_steps = F32(self.last_update_step)
    s = F32(min(F32(step), decay_steps))
    completed = F32(s / decay_steps)
    cosine = F32(F32(0.5) * F32(F32(1.0) + F32(math.cos(F32(F32(math.pi) * completed)))))
    return F32(self.init_drop_fraction * cosine)


class ScaledLRUpdateSchedule(UpdateSchedule):
  """Scales the drop fraction with the learning rate (:329-347).  `optimizer.lr` is either a
  number-like (read every time) or a callable `lr(step)`."""

  def __init__(self, mask_updater, init_drop_fraction, update_freq, last_update_step, optimizer):
    self._optimizer = optimizer
    self._initial_lr = self._get_lr(0)
    super(ScaledLRUpdateSchedule, self).__init__(mask_updater, init_drop_fraction, update_freq, last_update_step)

  def _get_lr(self, step):
    lr = self._optimizer.lr
    return lr(step) if callable(lr) else lr

  def get_drop_fraction(self, step):
    current_lr = self._get_lr(step)
    return F32(F32(self.init_drop_fraction / F32(self._initial_lr)) * F32(current_lr))


class MaskUpdaterAdapter(object):
  """The `mask_updater` the schedules drive, over a rigl_b200 sparse optimizer and its batched CUDA mask update:
  the counterpart of rigl_tf2/mask_updaters.py:33-160 (`MaskUpdater.update_masks` / `prune_masks`), whose per-layer
  `generic_mask_update` is the same drop/grow as sparse_optimizers_base._get_update_op.

    update_masks(f): drop the fraction f of every layer's active weights by |mask * w| (no noise, like the TF2
                     updaters' default noise_std = 0) and grow as many by the optimizer's grow score
                     (SET: uniform draw, RigL: |dense gradient|); grown weights <- 0, their optimizer slots <- 0.
    prune_masks(f):  drop only (score_grow = None in the reference): every layer keeps its top n_ones - int(n_ones*f).
  """

  def __init__(self, sparse_optimizer):
    self._opt = sparse_optimizer
    self.val_x = self.val_y = None

  def update_masks(self, drop_fraction):
    opt = self._opt
    if hasattr(opt, 'collect_masked_grads'):
      opt.collect_masked_grads()
    old_std, opt.noise_std = opt.noise_std, 0.
    try:
      opt.drop_fraction = F32(drop_fraction)
      opt.mask_update_op()
    finally:
      opt.noise_std = old_std

  def prune_masks(self, prune_fraction):
    from . import _cabi
    opt = self._opt
    opt.drop_fraction = F32(prune_fraction)
    specs = []
    for m, w in zip(opt.get_masks(), opt.get_weights()):
      flat = w.data.view(-1)
      specs.append(dict(mask=m, weights=flat, score_grow=flat, flags=_cabi.LAYER_DROP_ONLY))
    if specs:
      opt._engine.run(specs, F32(prune_fraction))

  def set_validation_data(self, val_x, val_y):
    self.val_x, self.val_y = val_x, val_y
