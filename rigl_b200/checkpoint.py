"""Checkpoints with the reference's partial-restore semantics.

The reference trains from TF checkpoints and supports loading only a subset of variables from a
previous run -- masks only (lottery-ticket / "scratch" experiments) or parameters only --
selected by NAME SUFFIX (`utils.initialize_parameters_from_ckpt`,
rigl/imagenet_resnet/utils.py:93-125; flags `--initial_value_checkpoint`, `--load_mask_dir` of
imagenet_train_eval.py).  Variable names follow the reference's scopes: `<scope>/mask`,
`<scope>/weights`, optimizer slots `<scope>/weights/<slot>`, everything else by module path.

Format: one `.npz` per step (`model.ckpt-<step>.npz`), float32 arrays in the reference layouts (masks
as 0/1 float32, HWIO / [in,out] weights) + `global_step`.  Host-side only; the arrays are staged
through numpy.
"""
import collections
import glob
import os
import re

import numpy as np


class Handle(object):
  """get() -> ndarray, set(ndarray) for one variable."""

  def __init__(self, get, set_):
    self.get, self.set = get, set_


def _tensor_handle(t):
  def get():
    return t.detach().float().cpu().numpy()

  def set_(a):
    import torch
    with torch.no_grad():
      t.copy_(torch.from_numpy(np.ascontiguousarray(a)).to(t.device).to(t.dtype).reshape(t.shape))
  return Handle(get, set_)


def variables_of(model, optimizer=None):
  """OrderedDict name -> Handle for a rigl_b200 model: masks, masked weights, the remaining
  parameters / buffers and (optionally) the inner optimizer's per-parameter state tensors."""
  out = collections.OrderedDict()
  masked = {}
  for l in model.registry.layers():
    out[l.scope + '/mask'] = Handle(l.mask.numpy, l.mask.assign)
    out[l.scope + '/weights'] = _tensor_handle(l.weight)
    masked[id(l.weight)] = l.scope + '/weights'
  for name, p in list(model.named_parameters()) + list(model.named_buffers()):
    if id(p) not in masked:
      out[name.replace('.', '/')] = _tensor_handle(p)
  if optimizer is not None:
    names = {id(p): (masked.get(id(p)) or n.replace('.', '/')) for n, p in model.named_parameters()}
    for p, st in optimizer.state.items():
      for k, v in st.items():
        if hasattr(v, 'shape') and tuple(v.shape) == tuple(p.shape):
          out['%s/%s' % (names.get(id(p), 'param%d' % id(p)), k)] = _tensor_handle(v)
  return out


def save(model_dir, variables, global_step):
  """Writes model.ckpt-<step>.npz; returns its path."""
  os.makedirs(model_dir, exist_ok=True)
  path = os.path.join(model_dir, 'model.ckpt-%d.npz' % int(global_step))
  arrays = {k: np.asarray(h.get(), dtype=np.float32) for k, h in variables.items()}
  arrays['global_step'] = np.asarray(int(global_step), dtype=np.int64)
  tmp = path + '.tmp.npz'
  np.savez(tmp, **arrays)
  os.replace(tmp, path)
  return path


def latest_checkpoint(model_dir):
  """Path of the checkpoint with the highest step in `model_dir`, or None (tf.train.latest_checkpoint)."""
  if not model_dir or not os.path.isdir(model_dir):
    return None
  best, best_step = None, -1
  for p in glob.glob(os.path.join(model_dir, 'model.ckpt-*.npz')):
    m = re.search(r'model\.ckpt-(\d+)\.npz$', p)
    if m and int(m.group(1)) > best_step:
      best, best_step = p, int(m.group(1))
  return best


def restore(path, variables, strict=True):
  """Full restore; returns the stored global step.  strict: every variable must be present with
  the right number of elements."""
  with np.load(path) as z:
    for k, h in variables.items():
      if k not in z.files:
        if strict:
          raise KeyError('variable %s not in checkpoint %s' % (k, path))
        continue
      a = z[k]
      if strict and a.size != np.asarray(h.get()).size:
        raise ValueError('variable %s: checkpoint has %d elements, model %d' % (k, a.size, np.asarray(h.get()).size))
      h.set(a)
    return int(z['global_step']) if 'global_step' in z.files else 0


def initialize_parameters_from_ckpt(ckpt_path, model_dir, param_suffixes, variables, log=None):
  """utils.py:93-125: load from `ckpt_path` ONLY the variables whose name ends with one of
  `param_suffixes` (str or tuple, e.g. 'mask' or ('weights', 'gamma')), and only if training has
  not already started in `model_dir`.  Variables with a matching suffix that the checkpoint lacks
  are skipped (logged).  Returns the list of loaded names."""
  log = log or (lambda *a: None)
  if latest_checkpoint(model_dir) is not None:
    log('Training already started on this model, not loading from previously trained model')
    return []
  suffixes = (param_suffixes,) if isinstance(param_suffixes, str) else tuple(param_suffixes)
  loaded = []
  with np.load(ckpt_path) as z:
    present = {n for n in z.files if n.endswith(suffixes)}
    for name, h in variables.items():
      if name in present:
        log('Loading parameter variable from checkpoint: %s' % name)
        h.set(z[name])
        loaded.append(name)
      elif name.endswith(suffixes):
        log('Cannot find parameter variable in checkpoint, skipping: %s' % name)
  return loaded
