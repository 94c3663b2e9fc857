"""Generates rigl_b200/data/str_sparsities_resnet50.json by IMPORTING the reference's rigl/str_sparsities.py
and calling its read_all() (the per-layer ResNet-50 sparsities reported by the STR paper, keyed by the
reference's mask names through its _name_map_str).  Run in the build container, where /root/reference exists:

  python tools/make_str_table.py

The product loads the JSON for get_sparsities(method='str') (rigl_b200/sparse_utils.get_sparsities_str);
`register_str_table` still lets a caller supply another table."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/rigl/str_sparsities.py'


def main():
  spec = importlib.util.spec_from_file_location('ref_str_sparsities', REF)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  table = mod.read_all()
  out = {'generator': 'tools/make_str_table.py',
         'source': 'google-research/rigl rigl/str_sparsities.py read_all() (STR paper, ResNet-50)',
         'tables': [{'overall_sparsity': float(s), 'overall_sparsity_hex': float(s).hex(),
                     'per_mask': {k: float(v) for k, v in sorted(t.items())}} for s, t in sorted(table.items())]}
  path = os.path.join(ROOT, 'rigl_b200', 'data', 'str_sparsities_resnet50.json')
  with open(path, 'w') as f:
    json.dump(out, f, indent=0)
  print('wrote', path, len(out['tables']), 'tables', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
