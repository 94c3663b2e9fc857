"""Parity of the batched CUDA mask update with the CPU oracle: mask indices,
re-initialised weights and optimizer slots must be BIT-IDENTICAL (integer /
index work; tolerance = 0)."""
import numpy as np
import pytest
import torch

from oracle import rigl_oracle as orc
from rigl_b200 import _cabi
from rigl_b200.masks import MaskUpdateEngine, MaskVariable

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _t(a):
  return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)


def _run_case(layers_np, drop_fraction, grow_mode=_cabi.GROW_ZEROS, grow_divisor=1.0, acc_scale=0.0,
              reinit=False, check_stats=True):
  """layers_np: list of dicts(mask, w, g, noise?, slots?, grow_values?, score_drop?, n_prune?)."""
  eng = MaskUpdateEngine()
  specs, keep = [], []
  for i, ly in enumerate(layers_np):
    mv = MaskVariable('layer%d' % i, ly['mask'].shape, DEV)
    mv.assign(ly['mask'])
    w, g = _t(ly['w']).view(-1), _t(ly['g']).view(-1)
    spec = dict(mask=mv, weights=w, score_grow=g)
    if ly.get('noise') is not None:
      spec['noise'] = _t(ly['noise']).view(-1)
    if ly.get('slots'):
      spec['slots'] = [_t(s).view(-1) for s in ly['slots']]
    if ly.get('grow_values') is not None:
      spec['grow_values'] = _t(ly['grow_values']).view(-1)
    if ly.get('score_drop') is not None:
      spec['score_drop'] = _t(ly['score_drop']).view(-1)
    if 'n_prune' in ly:
      spec['n_prune'] = ly['n_prune']
    specs.append(spec)
  eng.run(specs, np.float32(drop_fraction), grow_mode=grow_mode, grow_divisor=grow_divisor,
          acc_scale=acc_scale, reinit_when_same=reinit)
  torch.cuda.synchronize()
  stats = eng.stats()
  for i, (ly, spec) in enumerate(zip(layers_np, specs)):
    m, w, g = ly['mask'], ly['w'], ly['g']
    if ly.get('score_drop') is not None:
      sd = ly['score_drop'].astype(np.float32)
    else:
      sd = np.abs(m.astype(np.float32) * w.astype(np.float32))
      if ly.get('noise') is not None:
        sd = (sd + ly['noise'].astype(np.float32)).astype(np.float32)
    sg = np.abs(g.astype(np.float32))
    if grow_mode == _cabi.GROW_ZEROS:
      grow = None
    elif grow_mode == _cabi.GROW_TENSOR:
      grow = ly['grow_values']
    elif grow_mode == _cabi.GROW_GRAD_SCALE:
      grow = (g.astype(np.float32) / np.float32(grow_divisor)).astype(np.float32)
    else:
      grow = (np.sign(g.astype(np.float32)) / np.float32(grow_divisor)).astype(np.float32)
    frac = drop_fraction
    if 'n_prune' in ly:      # emulate the override through an exact fraction-free path
      n_ones = int(m.sum())
      frac = None
    if frac is None:
      want = _oracle_with_n_prune(sd, sg, m, w, ly['n_prune'], grow, reinit, ly.get('slots') or [],
                                  (g.astype(np.float32) * np.float32(acc_scale)).astype(np.float32))
    else:
      want = orc.get_update_op(sd, sg, m, w, frac, grow_tensor=grow, reinit_when_same=reinit,
                               slots=ly.get('slots') or [],
                               slot_reset=(g.astype(np.float32) * np.float32(acc_scale)).astype(np.float32))
    got_mask = spec['mask'].numpy()
    assert np.array_equal(got_mask, want['mask']), 'layer %d mask differs at %d positions' % (
        i, int((got_mask != want['mask']).sum()))
    got_w = spec['weights'].cpu().numpy().reshape(w.shape)
    assert got_w.tobytes() == want['weights'].astype(np.float32).tobytes(), 'layer %d weights' % i
    for s_got, s_want in zip(spec.get('slots', []), want['slots']):
      assert s_got.cpu().numpy().reshape(w.shape).tobytes() == s_want.astype(np.float32).tobytes()
    if check_stats:
      assert stats[i][0] == int(m.sum()) and stats[i][1] == want['n_prune'] and stats[i][2] == want['n_keep']
  return stats


def _oracle_with_n_prune(sd, sg, m, w, n_prune, grow, reinit, slots, slot_reset):
  n_ones = int(m.sum())
  if n_ones == 0:
    frac = 0.0
  else:
    # find a float32 fraction reproducing n_prune exactly
    frac = None
    for cand in (np.float32(n_prune / n_ones), np.nextafter(np.float32(n_prune / n_ones), np.float32(2)),
                 np.float32((n_prune + 0.5) / n_ones)):
      if int(np.int32(np.float32(n_ones) * np.float32(cand))) == n_prune:
        frac = cand
        break
    assert frac is not None
  return orc.get_update_op(sd, sg, m, w, frac, grow_tensor=grow, reinit_when_same=reinit, slots=slots,
                           slot_reset=slot_reset)


def _layer(rng, shape, sparsity, noise_std=0., slots=0):
  w = rng.standard_normal(shape).astype(np.float32) * np.float32(0.05)
  g = rng.standard_normal(shape).astype(np.float32) * np.float32(1e-3)
  m = orc.get_mask_random_numpy(shape, sparsity, rng).astype(np.float32)
  ly = dict(mask=m, w=w, g=g)
  if noise_std:
    ly['noise'] = (rng.standard_normal(shape) * noise_std).astype(np.float32)
  if slots:
    ly['slots'] = [rng.standard_normal(shape).astype(np.float32) for _ in range(slots)]
  return ly


def test_mnist_layers_no_noise():
  rng = np.random.RandomState(0)
  layers = [_layer(rng, (784, 300), 0.9, slots=1), _layer(rng, (300, 100), 0.81, slots=1),
            _layer(rng, (100, 10), 0.0, slots=1)]
  _run_case(layers, 0.3)


def test_mnist_layers_injected_noise_and_acc_scale():
  rng = np.random.RandomState(1)
  layers = [_layer(rng, (784, 300), 0.9, noise_std=1e-5, slots=1),
            _layer(rng, (300, 100), 0.81, noise_std=1e-5, slots=2),
            _layer(rng, (100, 10), 0.0, noise_std=1e-5)]
  _run_case(layers, 0.2999, acc_scale=0.5)


@pytest.mark.parametrize('n', [1, 2, 5, 31, 32, 33, 127, 128, 129, 1000, 4097, 32768, 32769, 70001])
def test_ragged_sizes(n):
  rng = np.random.RandomState(n)
  _run_case([_layer(rng, (n,), 0.5, slots=1)], 0.37)


@pytest.mark.parametrize('frac', [0.0, 1.0, 0.5, 1e-9, 0.999999])
def test_extreme_fractions(frac):
  rng = np.random.RandomState(7)
  _run_case([_layer(rng, (3, 3, 16, 32), 0.8, slots=1), _layer(rng, (257,), 0.3)], frac)


def test_all_zero_scores_tie_break_by_index():
  # every drop score is exactly 0 and every grow score equal: pure index tie-break,
  # inactive positions outrank higher-index active ones (SURVEY hard part 4).
  n = 50000
  rng = np.random.RandomState(3)
  m = (rng.rand(n) < 0.3).astype(np.float32)
  ly = dict(mask=m, w=np.zeros(n, np.float32), g=np.full(n, 0.25, np.float32),
            slots=[np.ones(n, np.float32)])
  _run_case([ly], 0.4)


def test_heavy_duplicates():
  rng = np.random.RandomState(4)
  n = 200000
  m = (rng.rand(n) < 0.2).astype(np.float32)
  w = (rng.randint(-3, 4, n) * 0.125).astype(np.float32)        # 7 distinct values incl. 0
  g = (rng.randint(-2, 3, n) * 0.5).astype(np.float32)
  _run_case([dict(mask=m, w=w, g=g, slots=[rng.standard_normal(n).astype(np.float32)])], 0.3)


def test_negative_zero_and_negative_noise():
  n = 4096
  rng = np.random.RandomState(5)
  m = (rng.rand(n) < 0.5).astype(np.float32)
  w = rng.standard_normal(n).astype(np.float32) * 1e-6
  w[::7] = -0.0
  g = rng.standard_normal(n).astype(np.float32)
  g[::5] = -0.0
  noise = (rng.standard_normal(n) * 1e-5).astype(np.float32)     # scores go negative
  _run_case([dict(mask=m, w=w, g=g, noise=noise)], 0.3)


def test_all_ones_and_all_zeros_masks():
  rng = np.random.RandomState(6)
  a = _layer(rng, (100, 10), 0.0, slots=1)
  b = _layer(rng, (64, 64), 0.5)
  b['mask'][:] = 0
  _run_case([a, b], 0.3)


@pytest.mark.parametrize('mode,div', [(_cabi.GROW_GRAD_SCALE, 2.0), (_cabi.GROW_GRAD_SIGN, 4.0),
                                      (_cabi.GROW_TENSOR, 1.0)])
def test_grow_init_modes(mode, div):
  rng = np.random.RandomState(8)
  ly = _layer(rng, (3, 3, 32, 32), 0.85, slots=1)
  ly['g'][0, 0, 0, :5] = 0.0
  if mode == _cabi.GROW_TENSOR:
    ly['grow_values'] = rng.standard_normal(ly['w'].shape).astype(np.float32)
  _run_case([ly], 0.3, grow_mode=mode, grow_divisor=div, acc_scale=0.1)


def test_reinit_when_same_static_flavour():
  # SparseStaticOptimizer: grow score = mask, reinit_when_same (sparse_optimizers.py:109-123)
  rng = np.random.RandomState(9)
  ly = _layer(rng, (128, 64), 0.7, slots=1)
  ly['g'] = ly['mask'].copy()
  ly['grow_values'] = rng.standard_normal(ly['w'].shape).astype(np.float32)
  before = ly['mask'].copy()
  _run_case([ly], 0.3, grow_mode=_cabi.GROW_TENSOR, reinit=True)
  # (mask unchanged is asserted through oracle equality; double-check the invariant itself)
  want = orc.get_update_op(np.abs(before * ly['w']), before, before, ly['w'], 0.3,
                           grow_tensor=ly['grow_values'], reinit_when_same=True)
  assert np.array_equal(want['mask'], before)


def test_explicit_score_drop_entry():
  rng = np.random.RandomState(10)
  ly = _layer(rng, (500, 40), 0.6)
  ly['score_drop'] = rng.standard_normal(ly['w'].shape).astype(np.float32)     # arbitrary, incl. negative
  _run_case([ly], 0.25)


def test_n_prune_override():
  rng = np.random.RandomState(11)
  ly = _layer(rng, (1000,), 0.5)
  ly['n_prune'] = 123
  _run_case([ly], 0.9, check_stats=False)


def test_set_flavour_uniform_scores():
  rng = np.random.RandomState(12)
  ly = _layer(rng, (15, 25), 0.5, slots=1)
  ly['g'] = rng.rand(15, 25).astype(np.float32)
  _run_case([ly], 0.5)


def test_wrn22_2_layer_set(golden):
  case = [c for c in golden['cases'] if c['tag'] == 'wrn22_2_erk95'][0]
  rng = np.random.RandomState(13)
  layers = [_layer(rng, tuple(sh), float.fromhex(case['sparsities_hex'][n + '/mask:0']),
                   noise_std=1e-5, slots=1) for n, sh in case['layers']]
  _run_case(layers, 0.3)


def test_resnet50_erk80_full_layer_set(golden):
  """BASELINE config C2's 54 masked tensors (25.5 M weights) in one batched update."""
  case = [c for c in golden['cases'] if c['tag'] == 'r50_erk80'][0]
  rng = np.random.RandomState(14)
  layers = [_layer(rng, tuple(sh), float.fromhex(case['sparsities_hex'][n + '/mask:0']), slots=1)
            for n, sh in case['layers']]
  stats = _run_case(layers, 0.3)
  assert sum(s[0] for s in stats) == 5100630


def test_idempotent_count_and_repeatability():
  """Size-independent properties at full size: #ones conserved; two identical runs agree."""
  rng = np.random.RandomState(15)
  ly = _layer(rng, (3, 3, 512, 512), 0.956534)
  outs = []
  for _ in range(2):
    mv = MaskVariable('big', ly['mask'].shape, DEV)
    mv.assign(ly['mask'])
    w, g = _t(ly['w']).view(-1), _t(ly['g']).view(-1)
    eng = MaskUpdateEngine()
    eng.run([dict(mask=mv, weights=w, score_grow=g)], np.float32(0.3))
    outs.append(mv.bits.clone())
    assert mv.count_ones() == int(ly['mask'].sum())
  assert torch.equal(outs[0], outs[1])


# ---- the CUDA update on the exact inputs of the golden vectors produced by executing the reference's
# `_get_update_op` / `generic_mask_update` (tools/make_golden_update_op.py).  The oracle already matches
# those vectors bit-for-bit on the CPU (tests/test_update_op_golden.py); this closes the loop on the
# device.  Written after the round's GPU budget was spent: gated until its first validated run.
import json as _json
import os as _os

_GOLD_PATH = _os.path.join(_os.path.dirname(__file__), 'golden', 'update_op_golden.json')
with open(_GOLD_PATH) as _f:
  _GOLD = _json.load(_f)


def _gdec(e):
  return None if e is None else np.frombuffer(bytes.fromhex(e['hex']), dtype=np.dtype(e['dtype'])).reshape(e['shape']).copy()


@pytest.mark.parametrize('case', _GOLD['cases'], ids=[c['tag'] for c in _GOLD['cases']])
def test_cuda_update_on_reference_executed_golden_inputs(case):
  i = case['in']
  rigl = case['optimizer'] == 'SparseRigLOptimizerBase'
  g = _gdec(i['dense_grad']) if rigl else _gdec(i['score_grow'])
  layer = dict(mask=_gdec(i['mask']), w=_gdec(i['weights']), g=g, slots=[_gdec(s) for s in i['slots']])
  if i['noise'] is not None:
    layer['noise'] = _gdec(i['noise'])
  mode, div = _cabi.GROW_ZEROS, 1.0
  if case['grow_init'].startswith('grad_scale'):
    mode, div = _cabi.GROW_GRAD_SCALE, orc.extract_number(case['grow_init'])
  elif case['grow_init'].startswith('grad_sign'):
    mode, div = _cabi.GROW_GRAD_SIGN, orc.extract_number(case['grow_init'])
  _run_case([layer], np.float32(float.fromhex(case['drop_fraction'])), grow_mode=mode, grow_divisor=div,
            acc_scale=float(np.float32(float.fromhex(case['initial_acc_scale']))), reinit=case['reinit_when_same'],
            check_stats=False)


@pytest.mark.parametrize('mode', ['zeros', 'grad_scale', 'grad_sign'])
def test_explicit_signed_grow_scores_with_separate_gradient(mode):
  """`_get_update_op(score_drop, score_grow, ...)` ranks score_grow VERBATIM (base.py:305-318): the rigl_tf2
  updaters pass -|g|, i.e. all-negative scores whose order is the reverse of |g|'s.  The grad_* grow inits and
  the slot reset then read the stored dense gradient (base.py:540-564), not the score."""
  rng = np.random.RandomState(11)
  shape = (96, 130)
  m = orc.get_mask_random_numpy(shape, 0.7, rng).astype(np.float32)
  w = rng.standard_normal(shape).astype(np.float32)
  grad = rng.standard_normal(shape).astype(np.float32)
  sg = (-np.abs(grad) * (1 + rng.randint(0, 3, shape))).astype(np.float32)      # negative, ties, not |grad| order
  sg[rng.rand(*shape) < 0.05] = -0.0
  sd = (np.abs(w) * m + rng.standard_normal(shape) * 0.1).astype(np.float32)   # signed drop scores too
  slot = rng.standard_normal(shape).astype(np.float32)
  gmode, div, grow = _cabi.GROW_ZEROS, 1.0, None
  if mode == 'grad_scale':
    gmode, div, grow = _cabi.GROW_GRAD_SCALE, 4.0, (grad / np.float32(4.0)).astype(np.float32)
  elif mode == 'grad_sign':
    gmode, div, grow = _cabi.GROW_GRAD_SIGN, 8.0, (np.sign(grad) / np.float32(8.0)).astype(np.float32)
  acc = 0.25
  want = orc.get_update_op(sd, sg, m, w, np.float32(0.4), grow_tensor=grow, slots=[slot],
                           slot_reset=(grad * np.float32(acc)).astype(np.float32))
  mv = MaskVariable('signed', shape, DEV).assign(m)
  wd, sl = _t(w).view(-1), _t(slot).view(-1)
  spec = dict(mask=mv, weights=wd, score_grow=_t(sg).view(-1), score_drop=_t(sd).view(-1), slots=[sl],
              grad=_t(grad).view(-1), flags=_cabi.LAYER_GROW_SCORE_SIGNED)
  MaskUpdateEngine().run([spec], np.float32(0.4), grow_mode=gmode, grow_divisor=div, acc_scale=acc)
  assert np.array_equal(mv.numpy(), want['mask'])
  assert wd.cpu().numpy().reshape(shape).tobytes() == want['weights'].tobytes()
  assert sl.cpu().numpy().reshape(shape).tobytes() == want['slots'][0].tobytes()
  # and the default (|score|) ranking really differs on this input
  mv2 = MaskVariable('abs', shape, DEV).assign(m)
  MaskUpdateEngine().run([dict(mask=mv2, weights=_t(w).view(-1), score_grow=_t(sg).view(-1),
                               score_drop=_t(sd).view(-1))], np.float32(0.4))
  assert not np.array_equal(mv2.numpy(), want['mask'])


def test_in_kernel_noise_equals_materialised_noise():
  """rigl_mask_update_run_noise draws the drop-score noise inside the two scans that need it (no noise tensor);
  rigl_mask_noise_fill materialises the same values: an update with the in-kernel draw is bit-identical to the
  oracle (and to the tensor-noise path) fed that tensor.  Also checks the draw is a plausible N(0, std)."""
  from rigl_b200.masks import noise_fill
  rng = np.random.RandomState(21)
  shapes = [(3, 3, 64, 64), (784, 300), (70001,), (5,)]
  std, seed = 1e-3, (1234567 << 32) | 4100
  layers = [_layer(rng, sh, 0.8, slots=1) for sh in shapes]
  for ly in layers:
    ly['w'] = (ly['w'] * np.float32(0.02)).astype(np.float32)         # |w| comparable to the noise: it decides ranks
  keys = [101, 0xdeadbeef, 7, 0]
  noises = [noise_fill(int(np.prod(sh)), k, std, seed, DEV).cpu().numpy().reshape(sh) for sh, k in zip(shapes, keys)]
  big = noises[2].astype(np.float64)
  assert abs(big.mean()) < 5 * std / np.sqrt(big.size) and abs(big.std() - std) < 0.02 * std
  assert np.unique(big).size > 0.99 * big.size
  assert not np.array_equal(noises[0].ravel()[:5], noises[1].ravel()[:5])           # layer key matters
  other = noise_fill(5, 0, std, seed + 1, DEV).cpu().numpy()
  assert not np.array_equal(other, noises[3].ravel())                               # seed (global step) matters
  specs = []
  for i, (ly, k) in enumerate(zip(layers, keys)):
    mv = MaskVariable('n%d' % i, ly['mask'].shape, DEV).assign(ly['mask'])
    specs.append(dict(mask=mv, weights=_t(ly['w']).view(-1), score_grow=_t(ly['g']).view(-1),
                      slots=[_t(ly['slots'][0]).view(-1)], noise_key=k))
  MaskUpdateEngine().run(specs, np.float32(0.3), noise_std=std, noise_seed=seed)
  for ly, spec, nz in zip(layers, specs, noises):
    want = orc.rigl_mask_update(ly['mask'], ly['w'], ly['g'], np.float32(0.3), noise=nz, slots=ly['slots'])
    assert np.array_equal(spec['mask'].numpy(), want['mask'])
    assert spec['weights'].cpu().numpy().reshape(ly['w'].shape).tobytes() == want['weights'].tobytes()
    assert spec['slots'][0].cpu().numpy().reshape(ly['w'].shape).tobytes() == want['slots'][0].tobytes()
    # and the noise really changed the outcome relative to a noiseless update
  quiet = orc.rigl_mask_update(layers[1]['mask'], layers[1]['w'], layers[1]['g'], np.float32(0.3))
  assert not np.array_equal(quiet['mask'], specs[1]['mask'].numpy())


@pytest.mark.parametrize('acc_scale', [0.0, 0.5])
def test_three_optimizer_slots(acc_scale):
  """The reference resets EVERY optimizer slot at new connections (base.py:345-353, 555-564).  The kernels carry
  two slot pointers per layer (momentum; Adam's moments); a third one (amsgrad's max_exp_avg_sq) is reset from the
  old / new bitmaps after the update -- bit-identical to the oracle."""
  rng = np.random.RandomState(17)
  layers = [_layer(rng, (300, 100), 0.8, slots=3), _layer(rng, (3, 3, 16, 32), 0.6, slots=3)]
  _run_case(layers, 0.3, acc_scale=acc_scale)
