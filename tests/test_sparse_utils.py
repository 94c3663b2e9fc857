"""Host-side mask utilities of the product vs fixtures produced by the reference."""
import hashlib

import numpy as np
import pytest

from rigl_b200 import sparse_utils


class M(object):

  def __init__(self, name, shape):
    self.name, self.shape, self.dtype = name + '/mask:0', tuple(shape), np.float32
    self.value = None

  def assign(self, v):
    self.value = np.asarray(v)


def test_erk_and_uniform_bit_exact_vs_reference(golden):
  for case in golden['cases']:
    masks = [M(n, sh) for n, sh in case['layers']]
    sp = sparse_utils.get_sparsities(masks, case['method'], case['default_sparsity'], case['custom'],
                                     erk_power_scale=case['erk_power_scale'])
    assert set(sp) == set(case['sparsities_hex'])
    for name, hx in case['sparsities_hex'].items():
      assert float(sp[name]).hex() == hx, (case['tag'], name)
    for m in masks:
      assert m.value is None
      assert int(np.prod(m.shape)) - sparse_utils.get_n_zeros(int(np.prod(m.shape)), sp[m.name]) == \
          case['nnz'][m.name]


def test_random_mask_bit_exact_vs_reference(golden):
  for e in golden['random_masks']:
    m = sparse_utils.get_mask_random_numpy(e['shape'], e['sparsity'], np.random.RandomState(e['seed']))
    assert hashlib.sha256(np.packbits(m.astype(np.uint8).ravel()).tobytes()).hexdigest() == e['sha256']


@pytest.mark.parametrize('shape,sparsity,expected_ones',
                         [((30, 4), 0.5, 60), ((1, 2, 1, 4), 0.8, 2), ((30,), 0.1, 27)])
def test_mask_fraction(shape, sparsity, expected_ones):
  m = sparse_utils.get_mask_random(M('v', shape), sparsity, np.int32)
  assert m.sum() == expected_ones and m.dtype == np.int32


@pytest.mark.parametrize('dtype', [np.int32, np.float32, np.int64, np.float64])
def test_mask_dtype(dtype):
  assert sparse_utils.get_mask_random(M('v', (3, 2)), 0.5, dtype).dtype == dtype


@pytest.mark.parametrize('s', [0., 0.4, 0.9])
def test_sparsity_dict_random(s):
  masks = [M('var1', (2, 3)), M('var2', (2, 3)), M('var3', (1, 1, 3))]
  sp = sparse_utils.get_sparsities(masks, 'random', s, {'var1': 0.8})
  assert sp[masks[0].name] == 0.8 and sp[masks[1].name] == s and sp[masks[2].name] == s


@pytest.mark.parametrize('shape1,shape2,s', [((2, 3), (2, 3), 0.5), ((1, 1, 2, 3), (1, 1, 2, 3), 0.3),
                                             ((8, 6), (4, 3), 0.7), ((80, 4), (20, 20), 0.8),
                                             ((2, 6), (2, 3), 0.8)])
def test_erdos_renyi_scale(shape1, shape2, s):
  # rigl/sparse_utils_test.py:108-143
  masks = [M('var1', shape1), M('var2', shape2)]
  sp = sparse_utils.get_sparsities(masks, 'erdos_renyi', s, {})
  s1, s2 = sp[masks[0].name], sp[masks[1].name]
  n1, n2 = int(np.prod(shape1)), int(np.prod(shape2))
  uni = sparse_utils.get_n_zeros(n1, s) + sparse_utils.get_n_zeros(n2, s)
  cur = sparse_utils.get_n_zeros(n1, s1) + sparse_utils.get_n_zeros(n2, s2)
  assert abs(uni - cur) <= 2
  f1 = (shape1[-1] + shape1[-2]) / float(shape1[-1] * shape1[-2])
  f2 = (shape2[-1] + shape2[-2]) / float(shape2[-1] * shape2[-2])
  assert abs((1 - s1) / f1 - (1 - s2) / f2) < 1e-7


def test_errors():
  masks = [M('a', (2, 3))]
  with pytest.raises(ValueError):
    sparse_utils.get_sparsities(masks, 'random', 0.5, {'zzz': 0.1})
  with pytest.raises(ValueError):
    sparse_utils.get_sparsities(masks, 'nope', 0.5, {})
  with pytest.raises(ValueError):
    sparse_utils.get_sparsities(masks, 'str', 0.8, {})
  sparse_utils.register_str_table(0.8, {'a/mask:0': 0.25})
  assert sparse_utils.get_sparsities(masks, 'str', 0.8, {}) == {'a/mask:0': 0.25}
  assert sparse_utils.mask_extract_name_fn('x/y/mask:0') == 'x/y'


def test_mask_init_fn_assigns_exact_counts():
  np.random.seed(0)
  masks = [M('layer1', (784, 300)), M('layer2', (300, 100)), M('layer3', (100, 10))]
  fn = sparse_utils.get_mask_init_fn(masks, 'random', 0.9, {'layer2': 0.81, 'layer3': 0.0})
  assert all(m.value is None for m in masks)
  fn()
  assert [int(m.value.sum()) for m in masks] == [23520, 5700, 1000]


def test_calculate_sparsity_arrays():
  a = np.array([1, 0, 0, 1.], np.float32)
  b = np.zeros((2, 2), np.float32)
  assert abs(float(sparse_utils.calculate_sparsity([a, b])) - 0.75) < 1e-7
