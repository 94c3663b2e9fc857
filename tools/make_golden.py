"""Generates tests/golden/sparse_utils_golden.json from the REFERENCE ITSELF.

Runs only in the build container (needs /root/reference).  Imports the
reference's own `rigl/sparse_utils.py` unmodified, with `tensorflow.compat.v1`
(only `.logging.info` is touched on this path) and the un-vendored
`google_research.micronet_challenge.counting` stubbed in sys.modules, and calls
it on fake mask objects carrying the variable names / shapes the reference
models create.  Nothing here is product code; the GPU box only reads the JSON.

  python tools/make_golden.py
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/root/reference'


def _install_stubs():
  tf = types.ModuleType('tensorflow')
  compat = types.ModuleType('tensorflow.compat')
  v1 = types.ModuleType('tensorflow.compat.v1')
  v1.logging = types.SimpleNamespace(info=lambda *a, **k: None)
  tf.compat = compat
  compat.v1 = v1
  sys.modules['tensorflow'] = tf
  sys.modules['tensorflow.compat'] = compat
  sys.modules['tensorflow.compat.v1'] = v1
  gr = types.ModuleType('google_research')
  mc = types.ModuleType('google_research.micronet_challenge')
  cnt = types.ModuleType('google_research.micronet_challenge.counting')
  gr.micronet_challenge = mc
  mc.counting = cnt
  sys.modules['google_research'] = gr
  sys.modules['google_research.micronet_challenge'] = mc
  sys.modules['google_research.micronet_challenge.counting'] = cnt


class _Shape(object):

  def __init__(self, dims):
    self._d = list(dims)

  def as_list(self):
    return list(self._d)

  def __str__(self):
    return str(tuple(self._d))


class RefMask(object):

  def __init__(self, scope, shape):
    self.name = scope + '/mask:0'
    self.shape = _Shape(shape)


def _hex(d):
  return {k: float(v).hex() for k, v in d.items()}


def main():
  _install_stubs()
  sys.path.insert(0, REF)
  from rigl import sparse_utils as ref  # the reference, unmodified
  from oracle import rigl_oracle as orc

  out = {'generator': 'tools/make_golden.py', 'reference': 'google-research/rigl d39fc7d',
         'cases': []}

  def add_case(tag, layers, method, s, custom, erk_power_scale=1.0):
    masks = [RefMask(n, sh) for n, sh in layers]
    sp = ref.get_sparsities(masks, method, s, custom, erk_power_scale=erk_power_scale)
    nnz = {}
    for n, sh in layers:
      size = int(np.prod(sh))
      nnz[n + '/mask:0'] = size - ref.get_n_zeros(size, sp[n + '/mask:0'])
    out['cases'].append({'tag': tag, 'layers': [[n, list(sh)] for n, sh in layers],
                         'method': method, 'default_sparsity': s, 'custom': custom,
                         'erk_power_scale': erk_power_scale,
                         'sparsities_hex': _hex(sp), 'nnz': nnz})

  r50 = [(n, sh) for n, sh, _, _ in orc.resnet50_masked_layers()]
  add_case('r50_erk80', r50, 'erdos_renyi_kernel', 0.8, {})
  add_case('r50_erk90', r50, 'erdos_renyi_kernel', 0.9, {})
  add_case('r50_er80', r50, 'erdos_renyi', 0.8, {})
  add_case('r50_uniform80_firstdense', r50, 'random', 0.8, {'resnet_model/initial_conv': 0.0})
  add_case('r50_erk80_scale05', r50, 'erdos_renyi_kernel', 0.8, {}, erk_power_scale=0.5)
  add_case('r50_erk80_custom_fc', r50, 'erdos_renyi_kernel', 0.8, {'resnet_model/final_dense': 0.5})

  # WRN-22-2 (cifar_resnet/resnet_model.py:70-235), conv_1 unmasked by default.
  wrn = [('resnet_model/skip_conv_2', (1, 1, 16, 32)), ('resnet_model/conv_2_0_1', (3, 3, 16, 32))]
  wrn += [('resnet_model/conv_2_%s' % t, (3, 3, 32, 32)) for t in ('0_2', '1_1', '1_2', '2_1', '2_2')]
  wrn += [('resnet_model/skip_conv_3', (1, 1, 32, 64)), ('resnet_model/conv_3_0_1', (3, 3, 32, 64))]
  wrn += [('resnet_model/conv_3_%s' % t, (3, 3, 64, 64)) for t in ('0_2', '1_1', '1_2', '2_1', '2_2')]
  wrn += [('resnet_model/skip_conv_4', (1, 1, 64, 128)), ('resnet_model/conv_4_0_1', (3, 3, 64, 128))]
  wrn += [('resnet_model/conv_4_%s' % t, (3, 3, 128, 128)) for t in ('0_2', '1_1', '1_2', '2_1', '2_2')]
  wrn += [('resnet_model/logits', (128, 10))]
  add_case('wrn22_2_erk95', wrn, 'erdos_renyi_kernel', 0.95, {})

  mbv1_pw = [(32, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 512)] + \
            [(512, 512)] * 5 + [(512, 1024), (1024, 1024)]
  mb = [('resnet_model/contraction_1x1_%d' % i, (1, 1, a, b)) for i, (a, b) in enumerate(mbv1_pw)]
  mb += [('resnet_model/final_dense', (1024, 1000))]
  add_case('mobilenetv1_uniform90', mb, 'random', 0.9, {})

  mnist = [('layer1', (784, 300)), ('layer2', (300, 100)), ('layer3', (100, 10))]
  add_case('mnist_random90', mnist, 'random', 0.9, {'layer2': 0.9 * 0.9, 'layer3': 0.0})

  # The reference's own unit-test fixtures (rigl/sparse_utils_test.py:73-143).
  t3 = [('var1', (2, 4)), ('var2', (2, 3)), ('var3', (1, 1, 3))]
  for s in (0.1, 0.4, 0.9):
    add_case('test_er_custom_%g' % s, t3, 'erdos_renyi', s, {'var3': 0.8})
  for sh1, sh2, s in (((2, 3), (2, 3), 0.5), ((1, 1, 2, 3), (1, 1, 2, 3), 0.3),
                      ((8, 6), (4, 3), 0.7), ((80, 4), (20, 20), 0.8), ((2, 6), (2, 3), 0.8)):
    add_case('test_er_scale_%s_%s_%g' % (sh1, sh2, s), [('var1', sh1), ('var2', sh2)],
             'erdos_renyi', s, {})

  # get_mask_random_numpy with an explicit RandomState (sparse_utils.py:48-68).
  out['random_masks'] = []
  for shape, s, seed in (((30, 4), 0.5, 0), ((1, 2, 1, 4), 0.8, 1), ((30,), 0.1, 2),
                         ((3, 3, 16, 32), 0.795075, 3), ((784, 300), 0.9, 4),
                         ((3, 3, 512, 512), 0.956534, 5)):
    m = ref.get_mask_random_numpy(list(shape), s, np.random.RandomState(seed))
    entry = {'shape': list(shape), 'sparsity': s, 'seed': seed, 'n_ones': int(m.sum()),
             'sha256': hashlib.sha256(np.packbits(m.astype(np.uint8).ravel()).tobytes()).hexdigest()}
    if m.size <= 200:
      entry['mask'] = m.astype(int).ravel().tolist()
    out['random_masks'].append(entry)

  out['get_n_zeros'] = [[size, s, ref.get_n_zeros(size, s)]
                        for size, s in ((120, 0.5), (8, 0.8), (30, 0.1), (2359296, 0.956534),
                                        (9408, 0.142805), (235200, 0.9), (30000, 0.81))]

  path = os.path.join(ROOT, 'tests', 'golden', 'sparse_utils_golden.json')
  with open(path, 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
  print('wrote', path, len(out['cases']), 'cases')


if __name__ == '__main__':
  main()
