"""Update-step decision and drop fraction: the oracle AND the product's host schedule
(rigl_b200.sparse_optimizers_base) against vectors produced by executing the reference's
`is_mask_update_iter` / `get_drop_fraction` (rigl/sparse_optimizers_base.py:198-258) over
numpy-backed TF op stubs (tools/make_golden_schedule.py).  Bit-for-bit float32."""
import json
import os

import numpy as np
import pytest

from oracle import rigl_oracle as orc
from rigl_b200 import sparse_optimizers_base as host

with open(os.path.join(os.path.dirname(__file__), 'golden', 'schedule_golden.json')) as f:
  GOLD = json.load(f)


@pytest.mark.parametrize('case', GOLD['cases'], ids=['%s_%d_%d_%d' % (c['anneal'], c['begin'], c['end'], c['frequency'])
                                                     for c in GOLD['cases']])
def test_schedule_matches_reference_execution(case):
  b, e, fr, anneal, init = case['begin'], case['end'], case['frequency'], case['anneal'], case['initial']
  n_updates = 0
  for gs, last, is_upd, frac_hex in case['rows']:
    want = np.float32(float.fromhex(frac_hex))
    got_upd = orc.is_mask_update_iter(gs, last, b, e, fr)
    assert got_upd == bool(is_upd), (gs, last)
    got = orc.get_drop_fraction(anneal, init, gs, b, e, got_upd)
    assert np.float32(got).tobytes() == want.tobytes(), (gs, float(got), float(want))
    # the product's host-side schedule: same gate, same float32 value
    prod = host.host_drop_fraction(anneal, init, gs, b, e) if is_upd else np.float32(0.)
    assert np.float32(prod).tobytes() == want.tobytes(), ('host', gs, float(prod), float(want))
    n_updates += is_upd
  assert n_updates > 0
