import numpy as np
import pytest
import torch

from rigl_b200.masks import MaskVariable
from rigl_b200 import sparse_utils

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('shape', [(1,), (31,), (32,), (33,), (3, 3, 16, 32), (784, 300), (1000003,)])
def test_pack_unpack_popcount_roundtrip(shape):
  rng = np.random.RandomState(0)
  m = (rng.rand(*shape) < 0.37).astype(np.float32)
  mv = MaskVariable('v', shape, DEV)
  assert mv.count_ones() == int(np.prod(shape))          # all-ones at creation
  mv.assign(m)
  assert np.array_equal(mv.numpy(), m)
  assert mv.count_ones() == int(m.sum())
  n = int(np.prod(shape))
  bits = mv.bits.cpu().numpy().view(np.uint32)
  assert bits.size % 4 == 0
  flat = np.unpackbits(bits.view(np.uint8), bitorder='little')
  assert np.array_equal(flat[:n], m.ravel().astype(np.uint8)) and flat[n:].sum() == 0


def test_apply_mask_and_sparsity():
  rng = np.random.RandomState(1)
  m = (rng.rand(1000, 37) < 0.2).astype(np.float32)
  g = rng.standard_normal(m.shape).astype(np.float32)
  mv = MaskVariable('v', m.shape, DEV).assign(m)
  out = mv.apply_to(torch.from_numpy(g).to(DEV).view(-1), scale=0.5).cpu().numpy().reshape(m.shape)
  assert np.array_equal(out, np.where(m == 1, g * np.float32(0.5), np.float32(0)))
  want = 1.0 - m.sum() / m.size
  assert abs(float(sparse_utils.calculate_sparsity([mv])) - want) < 1e-6


def test_mask_init_fn_on_device_masks():
  masks = [MaskVariable('layer1', (784, 300), DEV), MaskVariable('layer2', (300, 100), DEV),
           MaskVariable('layer3', (100, 10), DEV)]
  np.random.seed(0)
  sparse_utils.get_mask_init_fn(masks, 'random', 0.9, {'layer2': 0.81, 'layer3': 0.0})()
  assert [m.count_ones() for m in masks] == [23520, 5700, 1000]
  with pytest.raises(ValueError):
    masks[0].assign(np.ones((3, 3)))
