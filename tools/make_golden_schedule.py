"""Generates tests/golden/schedule_golden.json by EXECUTING the reference's
`SparseSETOptimizerBase.is_mask_update_iter` / `get_drop_fraction`
(rigl/sparse_optimizers_base.py:198-258) over numpy-backed TF op stubs: which steps are update
steps and which drop fraction they use, for constant / cosine / exponential anneals.  The decision
structure (range test incl. the negative end_step rule, last_update + frequency <= step, the RAW
global step fed to cosine_decay with decay_steps = end - begin, the where() gate) is the reference's;
`tf.train.cosine_decay` itself is supplied from its documentation (float32 ops, cos / pow evaluated in
float64 on the float32 argument and rounded once -- TF's Eigen cosf is not bit-portable anyway).
Build container only.

  python tools/make_golden_schedule.py
"""
import json
import math
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden_update_op as base  # noqa: E402

F32, A = np.float32, base.A


def main():
  base._install_tf_stubs()
  m = sys.modules
  mo, ao = m['tensorflow.python.ops.math_ops'], m['tensorflow.python.ops.array_ops']
  mo.greater_equal = lambda a, b: A(a) >= A(b)
  mo.less_equal = lambda a, b: A(a) <= A(b)
  mo.less = lambda a, b: A(a) < A(b)
  mo.logical_or = lambda a, b: np.logical_or(A(a), A(b))
  mo.add = lambda a, b: A(a) + A(b)
  mo.divide = lambda a, b: (A(a) / A(b)).astype(np.result_type(A(a), A(b)))
  mo.multiply = lambda a, b, name=None: (A(a) * A(b)).astype(F32)
  mo.pow = lambda a, b: F32(math.pow(float(F32(A(a))), float(F32(b))))
  m['tensorflow.python.framework.ops'].convert_to_tensor = \
      lambda v, name=None, dtype=None: np.asarray(v, dtype=(F32 if isinstance(v, float) else np.int64))

  def cosine_decay(learning_rate, global_step, decay_steps, alpha=0.0, name=None):
    lr = F32(A(learning_rate))
    ds = F32(A(decay_steps))
    gs = F32(min(F32(A(global_step)), ds))
    completed = F32(gs / ds)
    cosine_decayed = F32(F32(0.5) * F32(F32(1.0) + F32(math.cos(float(F32(F32(math.pi) * completed))))))
    decayed = F32(F32(F32(1.0) - F32(alpha)) * cosine_decayed + F32(alpha))
    return F32(lr * decayed)
  m['tensorflow.python.training.learning_rate_decay'].cosine_decay = cosine_decay

  class Optimizer(object):
    def __init__(self, use_locking=False, name=None):
      pass
  m['tensorflow.python.training.optimizer'].Optimizer = Optimizer
  sys.path.insert(0, base.REF)
  from rigl import sparse_optimizers_base as ref       # the reference, unmodified

  cases = []
  grid = [(0, 100, 10, 'constant', 0.3), (5, 50, 7, 'cosine', 0.3), (0, 1000, 100, 'cosine', 0.5),
          (10, -1, 25, 'constant', 0.1), (3, 40, 4, 'exponential_3', 0.3), (0, 32000, 100, 'cosine', 0.3),
          (20, 60, 20, 'exponential_1', 0.9)]
  for begin, end, freq, anneal, init in grid:
    opt = ref.SparseSETOptimizerBase(None, begin, end, freq, drop_fraction=init, drop_fraction_anneal=anneal)
    last = -freq
    steps = sorted(set(list(range(0, 70)) + list(range(90, 130)) + [999, 1000, 1001, 31900, 32000, 32100]))
    rows = []
    for gs in steps:
      g = np.asarray(gs, np.int64)
      try:
        is_upd = bool(opt.is_mask_update_iter(g, np.asarray(last, np.int64)))
      except ZeroDivisionError:
        continue
      frac = F32(A(opt.drop_fraction))
      rows.append([gs, last, int(is_upd), float(frac).hex()])
      if is_upd:
        last = gs                                        # cond_mask_update_op assigns last_update_step
    cases.append({'begin': begin, 'end': end, 'frequency': freq, 'anneal': anneal, 'initial': init, 'rows': rows})
  out = {'generator': 'tools/make_golden_schedule.py',
         'reference': 'rigl/sparse_optimizers_base.py:198-258 executed over numpy-backed TF op stubs', 'cases': cases}
  path = os.path.join(base.ROOT, 'tests', 'golden', 'schedule_golden.json')
  with open(path, 'w') as f:
    json.dump(out, f)
  print('wrote', path, sum(len(c['rows']) for c in cases), 'rows')


if __name__ == '__main__':
  main()
