"""Host logic of the product optimizers vs the oracle (no GPU: the mask update
itself is stubbed out, only schedule / step-skip / drop-fraction logic runs)."""
import numpy as np
import pytest
import torch

from oracle import rigl_oracle as orc
from rigl_b200 import sparse_optimizers_base as base


class _RiglNoKernel(base.SparseRigLOptimizerBase):

  def __init__(self, *a, **k):
    super(_RiglNoKernel, self).__init__(*a, **k)
    self.updates = []

  def get_weights(self):
    return []

  def get_masks(self):
    return []

  def get_masked_weights(self):
    return []

  def mask_update_op(self):
    self.updates.append((int(self._global_step), float(self.drop_fraction)))


class _SetNoKernel(base.SparseSETOptimizerBase):

  def __init__(self, *a, **k):
    super(_SetNoKernel, self).__init__(*a, **k)
    self.updates = []

  def get_weights(self):
    return []

  def get_masks(self):
    return []

  def mask_update_op(self):
    self.updates.append(int(self._global_step))


def _opt():
  p = torch.nn.Parameter(torch.zeros(3))
  p.grad = torch.ones(3)
  return torch.optim.SGD([p], lr=0.1), p


@pytest.mark.parametrize('sched,expect', [
    ((3, 7, 2), [1, 1, 1, 0, 1, 1, 0, 1, 1, 0, 1]),
    ((1, 5, 3), [1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 1]),
    ((0, 4, 1), [0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1])])
def test_rigl_apply_gradients_step_skip(sched, expect):
  inner, p = _opt()
  opt = _RiglNoKernel(inner, sched[0], sched[1], sched[2], drop_fraction=0.5)
  opt._weight2masked_grads = {'dummy': None}
  gs = base.GlobalStep(0)
  sim = orc.ScheduleSim('rigl', *sched)
  for one_if_incremented in expect:
    before, w_before = gs.value, p.detach().clone()
    opt.apply_gradients(None, global_step=gs)
    upd, stepped = sim.step()
    assert gs.value - before == one_if_incremented
    assert stepped == (one_if_incremented == 1) and upd == (one_if_incremented == 0)
    # optimizer step happened iff the step counter moved
    assert bool((p.detach() != w_before).any()) == (one_if_incremented == 1)
    assert gs.value == sim.global_step


def test_set_updates_on_incremented_step():
  inner, _ = _opt()
  opt = _SetNoKernel(inner, 1, 4, 2, drop_fraction=0.5)
  gs = base.GlobalStep(0)
  runs = []
  for i in range(1, 6):
    if opt.apply_gradients(None, global_step=gs):
      runs.append(i)
  assert runs == [1, 3] and opt.updates == [1, 3]


@pytest.mark.parametrize('anneal', ['constant', 'cosine', 'exponential_3', 'exponential_0.5'])
def test_drop_fraction_bit_exact_vs_oracle(anneal):
  for begin, end in ((0, 25000), (100, 3000), (0, 100)):
    for gs in list(range(begin, min(end, begin + 50))) + [end // 2, end - 1, end]:
      want = orc.get_drop_fraction(anneal, 0.3, gs, begin, end, True)
      got = base.host_drop_fraction(anneal, 0.3, gs, begin, end)
      assert np.float32(got).tobytes() == np.float32(want).tobytes(), (anneal, gs)


def test_drop_fraction_attribute_gated_by_schedule():
  inner, _ = _opt()
  opt = _RiglNoKernel(inner, 0, 100, 10, drop_fraction=0.3, drop_fraction_anneal='cosine')
  assert opt.is_mask_update_iter(0, -10) is True and opt.drop_fraction == np.float32(0.3)
  assert opt.is_mask_update_iter(5, 0) is False and opt.drop_fraction == 0
  assert opt.is_mask_update_iter(200, 0) is False           # beyond end_step
  opt2 = _RiglNoKernel(inner, 0, -1, 10, drop_fraction=0.3)  # negative end: never stops
  assert opt2.is_mask_update_iter(10 ** 6, 0) is True
  with pytest.raises(ValueError):
    _RiglNoKernel(inner, 0, 100, 10, drop_fraction_anneal='bogus').is_mask_update_iter(0, -10)


def test_extract_number_and_hash():
  assert base.extract_number('grad_scale_.5') == 0.5 and base.extract_number('zeros') == 1.0
  assert base.stable_hash('a/weights:0drop') == base.stable_hash('a/weights:0drop')
