"""The oracle's SNIP / DNW masks against golden vectors produced by EXECUTING the reference's
SparseSnipOptimizer / SparseDNWOptimizer.apply_gradients (rigl/sparse_optimizers.py:258-337, 408-470),
including its own get_mask_init_fn / get_sparsities, over numpy-backed TF op stubs
(tools/make_golden_snip_dnw.py).  Bit-for-bit."""
import json
import os

import numpy as np
import pytest

from oracle import rigl_oracle as orc

with open(os.path.join(os.path.dirname(__file__), 'golden', 'snip_dnw_golden.json')) as f:
  GOLD = json.load(f)


def _dec(e):
  return np.frombuffer(bytes.fromhex(e['hex']), dtype=np.dtype(e['dtype'])).reshape(e['shape']).copy()


@pytest.mark.parametrize('case', GOLD['cases'], ids=[c['tag'] for c in GOLD['cases']])
def test_snip_dnw_masks_match_reference_execution(case):
  shapes = [tuple(s) for s in case['shapes']]
  fake = [orc.FakeMask('layer%d/mask:0' % (i + 1), sh) for i, sh in enumerate(shapes)]
  sp = orc.get_sparsities(fake, case['method'], case['sparsity'], case['custom'])
  for i, sh in enumerate(shapes):
    s = sp['layer%d/mask:0' % (i + 1)]
    w = _dec(case['weights'][i])
    if case['kind'] == 'snip':
      got = orc.snip_mask(_dec(case['grads'][i]), w, s)
    else:
      got = orc.dnw_mask(w, s)
    want = _dec(case['masks'][i])
    assert np.array_equal(got, want), (case['tag'], i)
    assert want.size - want.sum() == orc.get_n_zeros(want.size, s)
