"""TF2-style update schedules (rigl/rigl_tf2/mask_updaters.py:251-344): validity rule of
is_update_iter, the three drop-fraction laws, and that `update` only touches the masks for a positive
fraction."""
import math

import numpy as np
import pytest

from rigl_b200 import update_schedules as us


class FakeUpdater(object):
  def __init__(self):
    self.updates, self.prunes = [], []

  def update_masks(self, f):
    self.updates.append(float(f))

  def prune_masks(self, f):
    self.prunes.append(float(f))


def test_is_update_iter_rule():
  s = us.ConstantUpdateSchedule(FakeUpdater(), 0.3, 100, 1000)
  assert [s.is_update_iter(i) for i in (0, 50, 100, 1000, 1100)] == [True, False, True, True, False]
  assert us.ConstantUpdateSchedule(FakeUpdater(), 0.3, 100, -1).is_update_iter(10 ** 6)      # no last step
  assert not us.ConstantUpdateSchedule(FakeUpdater(), 0.3, 100, 0).is_update_iter(0)         # never update
  with pytest.raises(ValueError):
    s.is_update_iter(-1)


def test_constant_and_update_side_effects():
  up = FakeUpdater()
  s = us.ConstantUpdateSchedule(up, 0.3, 10, 100)
  s.update(20)
  assert up.updates == [pytest.approx(0.3)] and s.last_drop_fraction == np.float32(0.3)
  with pytest.raises(ValueError):
    s.update(25)
  s.update(25, check_update_iter=False)
  assert len(up.updates) == 2
  z = us.ConstantUpdateSchedule(up, 0.0, 10, 100)
  z.update(10)
  assert len(up.updates) == 2                                   # zero fraction: masks untouched
  s.prune(0.5)
  assert up.prunes == [0.5] and s.last_drop_fraction == 0.5


def test_cosine_matches_keras_cosine_decay():
  s = us.CosineUpdateSchedule(FakeUpdater(), 0.3, 100, 1000)
  for step in (0, 100, 250, 500, 999, 1000, 5000):
    want = 0.3 * 0.5 * (1 + math.cos(math.pi * min(step, 1000) / 1000))
    got = s.get_drop_fraction(step)
    assert isinstance(got, np.float32) and abs(float(got) - want) < 1e-7
  assert s.get_drop_fraction(0) == np.float32(0.3) and s.get_drop_fraction(1000) < 1e-7
  up = FakeUpdater()
  s = us.CosineUpdateSchedule(up, 0.3, 100, 1000)
  s.update(1000)                                                 # fraction ~0 (cos rounding): at most a no-op update
  assert all(f < 1e-7 for f in up.updates)


def test_scaled_lr_schedule():
  class Opt(object):
    pass
  o = Opt()
  o.lr = lambda step: 0.1 * (0.5 ** (step // 100))
  s = us.ScaledLRUpdateSchedule(FakeUpdater(), 0.3, 100, -1, o)
  assert float(s.get_drop_fraction(0)) == pytest.approx(0.3, rel=1e-6)
  assert float(s.get_drop_fraction(100)) == pytest.approx(0.15, rel=1e-6)
  assert float(s.get_drop_fraction(250)) == pytest.approx(0.075, rel=1e-6)
  o2 = Opt()
  o2.lr = 0.2                                                    # variable-like learning rate, read each time
  s2 = us.ScaledLRUpdateSchedule(FakeUpdater(), 0.4, 10, -1, o2)
  o2.lr = 0.05
  assert float(s2.get_drop_fraction(7)) == pytest.approx(0.1, rel=1e-6)
