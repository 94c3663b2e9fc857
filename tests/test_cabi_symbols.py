"""The C-ABI library loads without a GPU and exports exactly what include/rigl_b200.h declares."""
import os
import re
import subprocess

from rigl_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  with open(os.path.join(ROOT, 'include', 'rigl_b200.h')) as f:
    text = f.read()
  return sorted(set(re.findall(r'RIGL_API[^;(]*?\b(rigl_\w+)\s*\(', text)))


def test_header_library_and_binding_agree():
  declared = _declared()
  assert len(declared) >= 20
  lib = _cabi.lib()
  for name in declared:
    assert hasattr(lib, name), 'library does not export %s' % name
  assert sorted(_cabi.SIGNATURES) == declared
  out = subprocess.check_output(['nm', '-D', '--defined-only', _cabi.LIB_PATH], text=True)
  exported = sorted(set(re.findall(r' T (rigl_\w+)', out)))
  assert exported == declared


def test_host_only_entry_points():
  lib = _cabi.lib()
  assert lib.rigl_version() >= 100
  assert lib.rigl_mask_words(1) == 4 and lib.rigl_mask_words(128) == 4 and lib.rigl_mask_words(129) == 8
  assert lib.rigl_packed_weights_bytes(1, 8, 8) > 0 and lib.rigl_packed_weights_bytes(0, 8, 8) == 0
  # argument validation happens before any CUDA call
  assert lib.rigl_mask_pack_f32(None, 10, None, None) == -1
  assert b'bad arguments' in lib.rigl_last_error()


def test_product_never_touches_the_oracle():
  pkg = os.path.join(ROOT, 'rigl_b200')
  for dirpath, _, files in os.walk(pkg):
    for fn in files:
      if fn.endswith(('.py', '.cu', '.cuh', '.h')):
        with open(os.path.join(dirpath, fn)) as f:
          src = f.read()
        assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), fn
        assert 'rigl_oracle' not in src, fn
