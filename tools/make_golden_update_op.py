"""Generates tests/golden/update_op_golden.json by EXECUTING THE REFERENCE'S OWN CODE for the mask update.

Runs only in the build container (needs /root/reference).  TensorFlow cannot be installed here, so
`rigl/sparse_optimizers_base.py` is imported unmodified on top of a tiny eager, numpy-backed stand-in
for the handful of `tensorflow.python.*` ops that `SparseSETOptimizerBase._get_update_op`,
`reset_momentum`, `get_grow_tensor` and `SparseRigLOptimizerBase.generic_mask_update / get_grow_tensor /
reset_momentum` touch (cast, reduce_sum/min, top_k, where, scatter_nd, reshape, assign, ...).  The
dataflow -- which scores are lifted, how n_prune is rounded, which positions count as new connections,
what the slots are reset to -- is therefore the reference's, not a restatement.  The only semantics
supplied here are those of the TF primitives themselves, taken from their documentation:
`tf.nn.top_k` returns equal elements lower index first; `tf.cast(float -> int32)` truncates toward
zero; arithmetic on float32 tensors is float32.  The random noise of `generic_mask_update` is
replaced by a recorded tensor (the reference seeds it with a per-process salted `hash(name)`).

  python tools/make_golden_update_op.py
"""
import contextlib
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
F32 = np.float32


class Var(object):
  """Stand-in for a tf.Variable / resource variable: a named numpy array that can be assigned."""

  def __init__(self, name, value):
    self.name = name
    self.value = np.array(value)
    self.initial_value = self.value.copy()

  @property
  def dtype(self):
    return self.value.dtype

  @property
  def shape(self):
    return self.value.shape

  def __array__(self, dtype=None, copy=None):
    return self.value if dtype is None else self.value.astype(dtype)


def A(x):
  return x.value if isinstance(x, Var) else np.asarray(x)


def _install_tf_stubs():
  def mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m
  tf = mod('tensorflow')
  py = mod('tensorflow.python')
  fw = mod('tensorflow.python.framework')
  opsm = mod('tensorflow.python.ops')
  tpu = mod('tensorflow.python.tpu')
  tpu_o = mod('tensorflow.python.tpu.ops')
  tr = mod('tensorflow.python.training')
  tf.python = py
  py.framework, py.ops, py.tpu, py.training = fw, opsm, tpu, tr
  tpu.ops = tpu_o

  dtypes = mod('tensorflow.python.framework.dtypes')
  dtypes.float32, dtypes.int32, dtypes.int64, dtypes.bool = np.dtype('float32'), np.dtype('int32'), np.dtype('int64'), np.dtype('bool')
  fw.dtypes = dtypes
  ops = mod('tensorflow.python.framework.ops')
  ops.control_dependencies = lambda deps: contextlib.nullcontext()
  fw.ops = ops

  array_ops = mod('tensorflow.python.ops.array_ops')
  array_ops.size = lambda x: int(A(x).size)
  array_ops.reshape = lambda x, shape: A(x).reshape(tuple(shape))
  array_ops.expand_dims = lambda x, axis: np.expand_dims(A(x), axis)
  array_ops.where = lambda c, a, b: np.where(A(c), A(a), A(b))
  array_ops.ones_like = lambda x, dtype=None: np.ones_like(A(x), dtype=dtype or A(x).dtype)
  array_ops.zeros_like = lambda x, dtype=None: np.zeros_like(A(x), dtype=dtype or A(x).dtype)
  array_ops.stack = lambda xs: np.stack([A(x) for x in xs])

  def scatter_nd(indices, updates, shape):
    out = np.zeros(tuple(shape), A(updates).dtype)
    idx = A(indices)[:, 0]
    assert len(np.unique(idx)) == len(idx)           # (duplicates would be summed by TF; never happens here)
    out[idx] = A(updates)
    return out
  array_ops.scatter_nd = scatter_nd

  math_ops = mod('tensorflow.python.ops.math_ops')

  def cast(x, dtype=None, **kw):
    a, dt = A(x), np.dtype(dtype)
    if a.dtype.kind == 'f' and dt.kind in 'iu':
      return np.trunc(a).astype(dt)                  # tf.cast truncates toward zero
    return a.astype(dt)
  math_ops.cast = cast
  math_ops.reduce_sum = lambda x: A(x).sum(dtype=A(x).dtype)
  math_ops.reduce_min = lambda x: A(x).min()
  math_ops.reduce_mean = lambda x: A(x).mean(dtype=A(x).dtype)
  math_ops.reduce_std = lambda x: A(x).std(dtype=A(x).dtype)
  math_ops.range = lambda n: np.arange(int(n), dtype=np.int32)
  math_ops.equal = lambda a, b: A(a) == A(b)
  math_ops.logical_and = lambda a, b: np.logical_and(A(a), A(b))
  math_ops.logical_not = lambda a: np.logical_not(A(a))
  math_ops.abs = lambda a: np.abs(A(a))
  math_ops.sign = lambda a: np.sign(A(a))

  nn_ops = mod('tensorflow.python.ops.nn_ops')

  def top_k(x, k):
    a = A(x)
    order = np.argsort(-a, kind='stable')[:int(k)]   # descending, equal elements: lower index first
    return a[order], order.astype(np.int32)
  nn_ops.top_k = top_k

  state_ops = mod('tensorflow.python.ops.state_ops')

  def assign(var, value, **kw):
    var.value = np.array(A(value), dtype=var.value.dtype).reshape(var.value.shape)
    return var.value
  state_ops.assign = assign

  cf = mod('tensorflow.python.ops.control_flow_ops')

  def Assert(cond, data, **kw):
    assert bool(A(cond)), 'tf Assert failed'
    return None
  cf.Assert = Assert
  cf.group = lambda xs, **kw: list(xs)
  for name in ('init_ops', 'random_ops', 'stateless_random_ops', 'variable_scope'):
    setattr(opsm, name, mod('tensorflow.python.ops.' + name))
  opsm.array_ops, opsm.math_ops, opsm.nn_ops, opsm.state_ops, opsm.control_flow_ops = array_ops, math_ops, nn_ops, state_ops, cf
  tpu_ops = mod('tensorflow.python.tpu.ops.tpu_ops')
  tpu_o.tpu_ops = tpu_ops
  for name in ('learning_rate_decay', 'training_util'):
    setattr(tr, name, mod('tensorflow.python.training.' + name))
  opt = mod('tensorflow.python.training.optimizer')
  opt.Optimizer = object
  tr.optimizer = opt


class FakeInnerOptimizer(object):
  def __init__(self, slots):
    self._slots = slots                              # {slot_name: {weights.name: Var}}

  def get_slot_names(self):
    return sorted(self._slots)

  def get_slot(self, var, name):
    return self._slots[name][var.name]


def _enc(a):
  a = np.asarray(a)
  return {'shape': list(a.shape), 'dtype': str(a.dtype), 'hex': a.astype(a.dtype).tobytes().hex()}


def main():
  _install_tf_stubs()
  sys.path.insert(0, REF)
  from rigl import sparse_optimizers_base as ref     # the reference, unmodified

  cases = []
  rng = np.random.RandomState(20260923)

  def run(tag, cls, shape, sparsity, drop_fraction, grow_init='zeros', reinit=False, ties=False, acc_scale=0.,
          n_slots=1, via_generic=False, static=False):
    n = int(np.prod(shape))
    mask0 = (rng.rand(*shape) >= sparsity).astype(F32)
    if ties:                                         # many equal scores: exercises the top_k tie rule
      w0 = rng.randint(-2, 3, size=shape).astype(F32)
      g0 = rng.randint(-2, 3, size=shape).astype(F32)
    else:
      w0 = rng.standard_normal(shape).astype(F32)
      g0 = rng.standard_normal(shape).astype(F32)
    noise = (rng.standard_normal(shape) * 1e-5).astype(F32) if via_generic else None
    slots0 = [rng.standard_normal(shape).astype(F32) for _ in range(n_slots)]
    mask, weights = Var('layer/mask:0', mask0), Var('layer/weights:0', w0)
    slot_vars = {'slot%d' % i: {weights.name: Var('layer/weights/slot%d:0' % i, s)} for i, s in enumerate(slots0)}
    me = types.SimpleNamespace()
    me.drop_fraction = F32(drop_fraction)
    me._grow_init = grow_init
    me._optimizer = FakeInnerOptimizer(slot_vars)
    me._weight2masked_grads = {weights.name: g0}
    me._initial_acc_scale = F32(acc_scale)
    me._random_normal = lambda *a, **k: noise
    # bind the REFERENCE's methods of `cls` to the bare namespace object
    for fn in ('_get_update_op', 'reset_momentum', 'get_grow_tensor', 'generic_mask_update'):
      setattr(me, fn, types.MethodType(getattr(cls, fn), me))
    # (super(...) calls inside the RigL subclass need a real instance: route them explicitly)
    if cls is ref.SparseRigLOptimizerBase:
      base_grow = ref.SparseSETOptimizerBase.get_grow_tensor

      def rigl_grow(self, weights_, method):
        if method.startswith('grad_scale') or method.startswith('grad_sign'):
          return cls.get_grow_tensor(self, weights_, method)
        return base_grow(self, weights_, method)
      me.get_grow_tensor = types.MethodType(rigl_grow, me)
    if via_generic:
      cls.generic_mask_update(me, mask, weights)
      sd = sg = None
    else:
      sd = (np.abs(mask0 * w0)).astype(F32)
      sg = np.abs(g0).astype(F32) if cls is ref.SparseRigLOptimizerBase else rng.rand(*shape).astype(F32)
      if static:                                     # SparseStaticOptimizer.generic_mask_update: score_grow = mask
        sg = mask0.copy()
      cls._get_update_op(me, sd, sg, mask, weights, reinit_when_same=reinit)
    cases.append({
        'tag': tag, 'optimizer': cls.__name__, 'shape': list(shape), 'drop_fraction': float(F32(drop_fraction)).hex(),
        'grow_init': grow_init, 'reinit_when_same': bool(reinit), 'initial_acc_scale': float(F32(acc_scale)).hex(),
        'via_generic_mask_update': bool(via_generic),
        'in': {'mask': _enc(mask0), 'weights': _enc(w0), 'dense_grad': _enc(g0),
               'score_drop': None if sd is None else _enc(sd), 'score_grow': None if sg is None else _enc(sg),
               'noise': None if noise is None else _enc(noise), 'slots': [_enc(s) for s in slots0]},
        'out': {'mask': _enc(mask.value), 'weights': _enc(weights.value),
                'slots': [_enc(slot_vars['slot%d' % i][weights.name].value) for i in range(n_slots)]},
    })
    assert mask.value.sum() == mask0.sum(), tag      # sparsity is preserved by construction

  SET, RIGL = ref.SparseSETOptimizerBase, ref.SparseRigLOptimizerBase
  run('set_basic', SET, (6, 7), 0.5, 0.3)
  run('set_reinit', SET, (5, 9), 0.6, 0.5, reinit=True)
  run('set_ties', SET, (8, 8), 0.5, 0.4, ties=True)
  run('set_conv_shape', SET, (3, 3, 4, 8), 0.8, 0.3, n_slots=2)
  run('set_drop_all', SET, (4, 5), 0.5, 1.0)
  run('set_drop_none', SET, (4, 5), 0.5, 0.0)
  run('static_basic', SET, (6, 7), 0.5, 0.4, reinit=True, static=True)
  run('static_ties', SET, (8, 8), 0.6, 0.5, reinit=True, ties=True, static=True, n_slots=2)
  run('rigl_basic', RIGL, (6, 7), 0.5, 0.3)
  run('rigl_ties', RIGL, (8, 8), 0.7, 0.5, ties=True)
  run('rigl_grad_scale', RIGL, (3, 3, 4, 8), 0.8, 0.3, grow_init='grad_scale_2', acc_scale=0.5)
  run('rigl_grad_sign', RIGL, (5, 11), 0.9, 0.5, grow_init='grad_sign_10', acc_scale=1.0, n_slots=2)
  run('rigl_generic', RIGL, (7, 9), 0.6, 0.3, via_generic=True)
  run('rigl_generic_conv', RIGL, (3, 3, 8, 8), 0.9, 0.1, via_generic=True, grow_init='grad_scale_1', acc_scale=0.1)
  run('rigl_dense_mask', RIGL, (4, 6), 0.0, 0.25)

  out = {'generator': 'tools/make_golden_update_op.py',
         'reference': 'google-research/rigl rigl/sparse_optimizers_base.py, executed over numpy-backed TF op stubs',
         'cases': cases}
  path = os.path.join(ROOT, 'tests', 'golden', 'update_op_golden.json')
  with open(path, 'w') as f:
    json.dump(out, f)
  print('wrote', path, len(cases), 'cases', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
