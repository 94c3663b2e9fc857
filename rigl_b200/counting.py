"""FLOP / model-size counting used by `sparse_utils.get_stats`.

The reference imports `google_research.micronet_challenge.counting` (rigl/sparse_utils.py:26), a
third-party module that is not vendored in google-research/rigl.  This file restates the published
MicroNet-challenge counting rules it implements (count_ops for Conv2D / DepthWiseConv2D /
FullyConnected): a dot product of length n costs n multiplications and n-1 additions, a bias one more
addition per output, a sparse tensor stores its non-zeros at `param_bits` each plus a 1-bit mask per
element, ReLU is free.  Pinned against the reference's published numbers: with these rules `get_stats`
reproduces the README's ResNet-50 model sizes to all printed digits (dense 102.122 MB; ERK 0.9
13.499 MB; ERK 0.95 8.399 MB; uniform 0.9 / 0.95 with a dense first layer 13.532 / 8.433 MB; ...) and
its inference-FLOP multiples (tests/test_get_stats.py).
"""
import collections

import numpy as np

Conv2D = collections.namedtuple('Conv2D', ['input_size', 'kernel_shape', 'strides', 'padding', 'use_bias',
                                           'activation'])
DepthWiseConv2D = collections.namedtuple('DepthWiseConv2D', ['input_size', 'kernel_shape', 'strides', 'padding',
                                                             'use_bias', 'activation'])
FullyConnected = collections.namedtuple('FullyConnected', ['kernel_shape', 'use_bias', 'activation'])


def get_sparse_size(tensor_shape, param_bits, sparsity):
  """Bits needed for a tensor stored with a binary mask (no mask for dense tensors)."""
  n_elements = np.prod(tensor_shape)
  c_size = n_elements * param_bits * (1 - sparsity)
  if sparsity > 0:
    c_size += n_elements          # 1 bit per element
  return c_size


def get_conv_output_size(image_size, filter_size, padding, stride):
  if padding == 'same':
    pad = filter_size // 2
  elif padding == 'valid':
    pad = 0
  else:
    raise NotImplementedError('Padding: %s should be `same` or `valid`.' % padding)
  return int(np.ceil((image_size - filter_size + 1. + 2 * pad) / stride))


def _activation_ops(activation, n_output_elements):
  if activation in (None, 'relu', 'linear'):
    return 0, 0                     # comparisons are not counted
  if activation in ('swish', 'sigmoid'):
    # one exp ~ 1 mult/add pair for the challenge's purposes plus the scaling multiply
    extra = 2 if activation == 'swish' else 1
    return n_output_elements * extra, n_output_elements
  raise ValueError('activation %r not supported' % (activation,))


def count_ops(op, sparsity, param_bits):
  """-> (param_count in bits, n_mults, n_adds) of one layer at the given kernel sparsity."""
  flop_mults = flop_adds = param_count = 0
  if isinstance(op, Conv2D):
    k_size, _, c_in, c_out = op.kernel_shape
    param_count += get_sparse_size([k_size, k_size, c_in, c_out], param_bits, sparsity)
    stride = op.strides[0] if isinstance(op.strides, (tuple, list)) else op.strides
    vector_length = (k_size * k_size * c_in) * (1 - sparsity)
    n_output_elements = get_conv_output_size(op.input_size, k_size, op.padding, stride) ** 2 * c_out
    flop_mults += vector_length * n_output_elements
    flop_adds += (vector_length - 1) * n_output_elements
    if op.use_bias:
      param_count += c_out * param_bits
      flop_adds += n_output_elements
    m, a = _activation_ops(op.activation, n_output_elements)
    flop_mults, flop_adds = flop_mults + m, flop_adds + a
  elif isinstance(op, DepthWiseConv2D):
    k_size, _, channels, _ = op.kernel_shape
    param_count += get_sparse_size([k_size, k_size, channels], param_bits, sparsity)
    stride = op.strides[0] if isinstance(op.strides, (tuple, list)) else op.strides
    vector_length = (k_size * k_size) * (1 - sparsity)
    n_output_elements = get_conv_output_size(op.input_size, k_size, op.padding, stride) ** 2 * channels
    flop_mults += vector_length * n_output_elements
    flop_adds += (vector_length - 1) * n_output_elements
    if op.use_bias:
      param_count += channels * param_bits
      flop_adds += n_output_elements
    m, a = _activation_ops(op.activation, n_output_elements)
    flop_mults, flop_adds = flop_mults + m, flop_adds + a
  elif isinstance(op, FullyConnected):
    c_in, c_out = op.kernel_shape
    param_count += get_sparse_size([c_in, c_out], param_bits, sparsity)
    flop_mults += c_in * (1 - sparsity) * c_out
    flop_adds += (c_in * (1 - sparsity) - 1) * c_out
    if op.use_bias:
      param_count += c_out * param_bits
      flop_adds += c_out
    m, a = _activation_ops(op.activation, c_out)
    flop_mults, flop_adds = flop_mults + m, flop_adds + a
  else:
    raise ValueError('Encountered unknown operation %s.' % str(op))
  return param_count, flop_mults, flop_adds
