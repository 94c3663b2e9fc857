"""Generates tests/golden/snip_dnw_golden.json by EXECUTING the reference's SparseSnipOptimizer and
SparseDNWOptimizer (rigl/sparse_optimizers.py:217-480) -- `apply_gradients`, the nested snip_fn /
dnw_fn and the reference's own `sparse_utils.get_mask_init_fn` / `get_sparsities` -- over the
numpy-backed TF op stand-ins of tools/make_golden_update_op.py (plus eager `cond`, `get_variable`,
`tf.assign` / `tf.group`).  Build container only (needs /root/reference).

  python tools/make_golden_snip_dnw.py
"""
import json
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden_update_op as base  # noqa: E402

ROOT, REF, F32 = base.ROOT, base.REF, np.float32


class Shape(tuple):
  def as_list(self):
    return list(self)


class TArr(np.ndarray):
  """ndarray whose .shape offers TensorShape.as_list()."""
  @property
  def shape(self):
    return Shape(np.ndarray.shape.__get__(self))


class Var(base.Var):
  @property
  def shape(self):
    return Shape(self.value.shape)

  def __mul__(self, o):
    return np.asarray(self.value) * base.A(o)
  __rmul__ = __mul__


def _install_more_stubs(state):
  base._install_tf_stubs()
  m = sys.modules
  math_ops, array_ops, cf = m['tensorflow.python.ops.math_ops'], m['tensorflow.python.ops.array_ops'], m['tensorflow.python.ops.control_flow_ops']
  math_ops.abs = lambda a: np.abs(base.A(a)).view(TArr)
  array_ops.reshape = lambda x, shape: base.A(x).reshape(tuple(shape))
  cf.cond = lambda pred, true_fn, false_fn: (true_fn() if bool(base.A(pred)) else false_fn())
  vs = m['tensorflow.python.ops.variable_scope']
  vs.get_variable = lambda name, initializer=None, trainable=True, **kw: Var(name + ':0', initializer())

  def mod(name):
    mm = types.ModuleType(name)
    m[name] = mm
    return mm
  variables = mod('tensorflow.python.ops.variables')
  variables.trainable_variables = lambda: []
  m['tensorflow.python.ops'].variables = variables
  ma = mod('tensorflow.python.training.moving_averages')
  ma.ExponentialMovingAverage = lambda decay: types.SimpleNamespace(decay=decay)
  m['tensorflow.python.training'].moving_averages = ma

  class Optimizer(object):
    def __init__(self, use_locking=False, name=None):
      self._name = name
  m['tensorflow.python.training.optimizer'].Optimizer = Optimizer
  # tensorflow.contrib.model_pruning.python.pruning: the getter seam
  contrib = mod('tensorflow.contrib')
  mp = mod('tensorflow.contrib.model_pruning')
  mpp = mod('tensorflow.contrib.model_pruning.python')
  pr = mod('tensorflow.contrib.model_pruning.python.pruning')
  pr.get_masks = lambda: state['masks']
  pr.get_weights = lambda: state['weights']
  pr.get_masked_weights = lambda: state['masked_weights']
  m['tensorflow'].contrib = contrib
  contrib.model_pruning, mp.python, mpp.pruning = mp, mpp, pr
  # tensorflow.compat.v1 as used by rigl/sparse_utils.py
  compat, v1 = mod('tensorflow.compat'), mod('tensorflow.compat.v1')
  v1.logging = types.SimpleNamespace(info=lambda *a, **k: None)
  v1.assign = m['tensorflow.python.ops.state_ops'].assign
  v1.group = lambda xs, **kw: list(xs)
  m['tensorflow'].compat = compat
  compat.v1 = v1
  gr, mc, cnt = mod('google_research'), mod('google_research.micronet_challenge'), mod('google_research.micronet_challenge.counting')
  gr.micronet_challenge, mc.counting = mc, cnt


class Inner(object):
  """Stand-in for the wrapped tf.train.Optimizer: records that a real step was asked for."""
  def __init__(self):
    self.steps = 0

  def apply_gradients(self, grads_and_vars, global_step=None, name=None):
    self.steps += 1
    return 'step'


def main():
  state = {'masks': [], 'weights': [], 'masked_weights': []}
  _install_more_stubs(state)
  sys.path.insert(0, REF)
  from rigl import sparse_optimizers as ref       # the reference, unmodified

  rng = np.random.RandomState(20260924)
  enc = base._enc
  cases = []

  def layers(shapes, ties=False):
    ws, ms, gs = [], [], []
    for i, sh in enumerate(shapes, 1):
      scope = 'layer%d' % i
      w = rng.randint(-3, 4, size=sh).astype(F32) if ties else rng.standard_normal(sh).astype(F32)
      g = rng.randint(-2, 3, size=sh).astype(F32) if ties else rng.standard_normal(sh).astype(F32)
      ws.append(Var(scope + '/weights:0', w))
      ms.append(Var(scope + '/mask:0', np.ones(sh, F32)))
      gs.append(g)
    return ws, ms, gs

  def snip(tag, shapes, method, sparsity, custom=None, ties=False):
    ws, ms, gs = layers(shapes, ties)
    state['masks'], state['weights'] = ms, ws
    inner = Inner()
    opt = ref.SparseSnipOptimizer(inner, sparsity, method, custom_sparsity_map=custom or {})
    w0 = [w.value.copy() for w in ws]
    opt.apply_gradients(list(zip(gs, ws)), global_step=0)
    assert bool(opt.is_snipped.value) and inner.steps == 0          # the snip iteration takes no optimizer step
    snipped = [m.value.copy() for m in ms]
    opt.apply_gradients(list(zip(gs, ws)), global_step=0)           # already snipped: plain step, masks unchanged
    assert inner.steps == 1 and all(np.array_equal(a, m.value) for a, m in zip(snipped, ms))
    cases.append({'tag': tag, 'kind': 'snip', 'method': method, 'sparsity': sparsity, 'custom': custom or {},
                  'shapes': [list(s) for s in shapes], 'weights': [enc(w) for w in w0], 'grads': [enc(g) for g in gs],
                  'masks': [enc(m) for m in snipped]})

  def dnw(tag, shapes, method, sparsity, custom=None, ties=False):
    ws, ms, gs = layers(shapes, ties)
    state['masks'], state['weights'] = ms, ws
    inner = Inner()
    opt = ref.SparseDNWOptimizer(inner, sparsity, method, custom_sparsity_map=custom or {})
    opt.apply_gradients(list(zip(gs, ws)), global_step=3)
    assert inner.steps == 1                                         # DNW: optimizer step, THEN the masks follow |w|
    cases.append({'tag': tag, 'kind': 'dnw', 'method': method, 'sparsity': sparsity, 'custom': custom or {},
                  'shapes': [list(s) for s in shapes], 'weights': [enc(w.value) for w in ws],
                  'masks': [enc(m.value) for m in ms]})

  mlp = [(20, 12), (12, 8), (8, 4)]
  conv = [(3, 3, 4, 8), (3, 3, 8, 8), (1, 1, 8, 16), (16, 10)]
  snip('snip_random_mlp', mlp, 'random', 0.5)
  snip('snip_random_80', mlp, 'random', 0.8)
  snip('snip_erk_conv', conv, 'erdos_renyi_kernel', 0.7)
  snip('snip_er_conv_custom', conv, 'erdos_renyi', 0.6, {'layer1': 0.25})
  snip('snip_ties', [(8, 8), (8, 6)], 'random', 0.5, ties=True)
  dnw('dnw_random_mlp', mlp, 'random', 0.5)
  dnw('dnw_erk_conv', conv, 'erdos_renyi_kernel', 0.8)
  dnw('dnw_ties', [(8, 8), (8, 6)], 'random', 0.75, ties=True)
  dnw('dnw_custom', mlp, 'random', 0.9, {'layer3': 0.0})

  out = {'generator': 'tools/make_golden_snip_dnw.py',
         'reference': 'google-research/rigl rigl/sparse_optimizers.py + rigl/sparse_utils.py, executed over numpy-backed TF op stubs',
         'cases': cases}
  path = os.path.join(ROOT, 'tests', 'golden', 'snip_dnw_golden.json')
  with open(path, 'w') as f:
    json.dump(out, f)
  print('wrote', path, len(cases), 'cases', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
