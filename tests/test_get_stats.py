"""`sparse_utils.get_stats` + the restated MicroNet counting rules (rigl_b200/counting.py) against the
numbers the reference publishes (README.md:31-79: inference FLOPs and model size of sparse
ResNet-50s and MobileNet-v1 produced by its `get_stats` colab) and against the closed forms."""
import numpy as np
import pytest

from oracle import rigl_oracle as orc
from rigl_b200 import counting
from rigl_b200 import sparse_utils as su


def _resnet50_specs():
  specs = []
  for name, shape, stride, out_hw in orc.resnet50_masked_layers():
    if len(shape) == 4:
      specs.append(su.LayerSpec('conv', name, shape, out_hw * stride, stride))
    else:
      specs.append(su.LayerSpec('dense', name, shape))
  return specs


def test_resnet50_dense_matches_readme():
  flops, bits, real = su.get_stats(_resnet50_specs(), 0., 'random')
  assert round(flops / 1e9, 1) == 8.2                    # README: Inference FLOPs 8.2e9
  assert round(bits / 8e6, 3) == 102.122                 # README: Model Size 102.122 (MB)
  assert real == 0.


FIRST = 'resnet_model/initial_conv'


@pytest.mark.parametrize('method,sparsity,custom,flops_x,size_mb', [
    ('erdos_renyi_kernel', 0.9, {}, 0.24, 13.499),       # README rows "ERK 0.9", "ERK 0.95"
    ('erdos_renyi_kernel', 0.95, {}, 0.12, 8.399),
    ('random', 0.9, {FIRST: 0.}, 0.13, 13.532),          # "Uniform": first layer kept dense
    ('random', 0.95, {FIRST: 0.}, 0.08, 8.433),
    ('erdos_renyi_kernel', 0.99, {FIRST: 0.}, 0.05, 4.354),
])
def test_resnet50_sparse_rows_match_readme(method, sparsity, custom, flops_x, size_mb):
  specs = _resnet50_specs()
  f0, _, _ = su.get_stats(specs, 0., 'random')
  f, bits, real = su.get_stats(specs, sparsity, method, custom)
  assert round(f / f0, 2) == flops_x
  assert round(bits / 8e6, 3) == size_mb
  assert abs(real - sparsity) < 2e-3


def test_resnet50_erk80_size():
  _, bits, _ = su.get_stats(_resnet50_specs(), 0.8, 'erdos_renyi_kernel')
  assert abs(bits / 8e6 - 23.683) < 2e-3                 # README 23.683


def test_mobilenet_v1_dense_flops():
  cfg = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1), (512, 1), (512, 1),
         (512, 1), (1024, 2), (1024, 1)]
  specs = [su.LayerSpec('conv', 'conv1', (3, 3, 3, 32), 224, 2)]
  cin, size = 32, 112
  for i, (f, s) in enumerate(cfg, 1):
    specs.append(su.LayerSpec('depthwise', 'conv_dw_%d' % i, (3, 3, cin, 1), size, s))
    size //= s
    specs.append(su.LayerSpec('conv', 'conv_pw_%d' % i, (1, 1, cin, f), size, 1))
    cin = f
  specs.append(su.LayerSpec('conv', 'conv_preds', (1, 1, 1024, 1000), 1, 1))
  flops, _, _ = su.get_stats(specs, 0., 'random')
  assert round(flops / 1e9, 2) == 1.14                   # README: 1.14e9


def test_count_ops_closed_forms():
  bits, m, a = counting.count_ops(counting.Conv2D(8, [3, 3, 4, 5], (2, 2), 'same', True, 'relu'), 0., 32)
  assert (bits, m, a) == ((3 * 3 * 4 * 5 + 5) * 32, 36 * 4 * 4 * 5, 35 * 80 + 80)
  bits, m, a = counting.count_ops(counting.FullyConnected([10, 4], True, 'relu'), 0.5, 32)
  assert bits == 10 * 4 * 32 * 0.5 + 40 + 4 * 32 and m == 20 and a == 4 * 4 + 4
  bits, m, a = counting.count_ops(counting.DepthWiseConv2D(8, [3, 3, 6, 1], (1, 1), 'same', False, None), 0., 16)
  assert (bits, m, a) == (54 * 16, 9 * 64 * 6, 8 * 64 * 6)
  assert counting.get_conv_output_size(224, 7, 'same', 2) == 112
  assert counting.get_conv_output_size(7, 3, 'valid', 1) == 5
  with pytest.raises(ValueError):
    counting.count_ops('conv', 0., 32)


def test_width_scaling_and_layer_spec_from_shapes():
  a = su.LayerSpec('conv', 'conv1/kernel', (3, 3, 3, 16), 32, 1)
  b = su.LayerSpec('conv', 'mid/kernel', (3, 3, 16, 16), 32, 1)
  c = su.LayerSpec('dense', 'conv_preds/kernel', (16, 10))
  f1, _, _ = su.get_stats([a, b, c], 0., 'random', width=1.)
  f2, _, _ = su.get_stats([a, b, c], 0., 'random', width=2.)
  # first layer: only d_out doubles; middle: both; last: only d_in (sparse_utils.py:425-431)
  c1 = lambda cin, cout, k=3, hw=32: (2 * k * k * cin) * hw * hw * cout
  assert f1 == c1(3, 16) + c1(16, 16) + 2 * 16 * 10
  assert f2 == c1(3, 32) + c1(32, 32) + 2 * 32 * 10
  with pytest.raises(ValueError):
    su.LayerSpec('pool', 'x', (1,))


def test_get_compressed_fc():
  """mnist_train_eval.py:165-189 on a hand-checkable 4-3-2 network."""
  m0 = np.array([[1, 0, 0],      # input 0 -> unit 0
                 [0, 0, 0],      # input 1 dead
                 [0, 0, 1],      # input 2 -> unit 2
                 [1, 0, 1]])     # input 3 -> units 0, 2     (unit 1 has no incoming edge)
  m1 = np.array([[1, 0],         # unit 0 -> out 0
                 [1, 1],         # unit 1 (dead upstream) -> both outputs
                 [0, 0]])        # unit 2 has no outgoing edge
  keep0, keep1 = m0.copy(), m1.copy()
  sp, sizes = su.get_compressed_fc([m0, m1])
  assert sizes == [3, 1, 1]                                  # 3 live inputs, only unit 0 survives, out 1 loses its only edge
  assert sp == [1.0 / 3.0, 0.0]
  assert (m0 == keep0).all() and (m1 == keep1).all()         # inputs untouched
  # a dense network is returned unchanged
  sp, sizes = su.get_compressed_fc([np.ones((5, 4)), np.ones((4, 3))])
  assert sp == [0.0, 0.0] and sizes == [5, 4, 3]
