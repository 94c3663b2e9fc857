"""The oracle's mask update against golden vectors produced by EXECUTING the reference's own
`_get_update_op` / `generic_mask_update` / `reset_momentum` / `get_grow_tensor`
(rigl/sparse_optimizers_base.py:276-353, 523-564) over numpy-backed TF op stubs
(tools/make_golden_update_op.py).  Bit-for-bit: masks, weights and optimizer slots."""
import json
import os

import numpy as np
import pytest

from oracle import rigl_oracle as orc

PATH = os.path.join(os.path.dirname(__file__), 'golden', 'update_op_golden.json')
with open(PATH) as f:
  GOLD = json.load(f)


def _dec(e):
  if e is None:
    return None
  return np.frombuffer(bytes.fromhex(e['hex']), dtype=np.dtype(e['dtype'])).reshape(e['shape']).copy()


@pytest.mark.parametrize('case', GOLD['cases'], ids=[c['tag'] for c in GOLD['cases']])
def test_oracle_update_op_matches_reference_execution(case):
  i = {k: (_dec(v) if not isinstance(v, list) else [_dec(x) for x in v]) for k, v in case['in'].items()}
  df = np.float32(float.fromhex(case['drop_fraction']))
  acc = np.float32(float.fromhex(case['initial_acc_scale']))
  rigl = case['optimizer'] == 'SparseRigLOptimizerBase'
  if case['via_generic_mask_update']:
    r = orc.rigl_mask_update(i['mask'], i['weights'], i['dense_grad'], df, noise=i['noise'],
                             grow_init=case['grow_init'], initial_acc_scale=acc, slots=i['slots'])
  elif rigl:
    r = orc.get_update_op(i['score_drop'], i['score_grow'], i['mask'], i['weights'], df,
                          grow_tensor=orc.rigl_grow_tensor(case['grow_init'], i['weights'], i['dense_grad']),
                          reinit_when_same=case['reinit_when_same'], slots=i['slots'],
                          slot_reset=(i['dense_grad'] * acc).astype(np.float32))
  else:
    r = orc.get_update_op(i['score_drop'], i['score_grow'], i['mask'], i['weights'], df,
                          reinit_when_same=case['reinit_when_same'], slots=i['slots'])
  want_mask, want_w = _dec(case['out']['mask']), _dec(case['out']['weights'])
  assert np.array_equal(r['mask'].astype(np.float32), want_mask)
  assert r['weights'].dtype == np.float32 and r['weights'].tobytes() == want_w.tobytes()
  for got, want in zip(r['slots'], [_dec(s) for s in case['out']['slots']]):
    assert got.astype(np.float32).tobytes() == want.tobytes()
  assert want_mask.sum() == i['mask'].sum()


def test_golden_covers_both_optimizers_and_the_tie_rule():
  tags = {c['tag'] for c in GOLD['cases']}
  assert {'set_ties', 'rigl_ties', 'rigl_generic', 'rigl_grad_scale', 'rigl_grad_sign', 'set_reinit'} <= tags
  assert GOLD['generator'] == 'tools/make_golden_update_op.py'


def test_static_cases_keep_the_mask_and_reinitialise_dropped_weights():
  """SparseStaticOptimizer (sparse_optimizers.py:109-123): grow score = mask, reinit_when_same=True --
  the connectivity never changes, the dropped (weakest) connections restart from the grow tensor."""
  for case in GOLD['cases']:
    if not case['tag'].startswith('static'):
      continue
    m0, w0 = _dec(case['in']['mask']), _dec(case['in']['weights'])
    m1, w1 = _dec(case['out']['mask']), _dec(case['out']['weights'])
    assert np.array_equal(m0, m1)
    changed = w0 != w1
    assert changed.any() and (w1[changed] == 0).all() and (m0[changed] == 1).all()
