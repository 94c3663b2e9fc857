"""Blackwell-instruction evidence from the SHIPPED library: per kernel, how many tcgen05 / TMEM / TMA SASS
instructions `cuobjdump -sass rigl_b200/librigl_b200.so` shows (mnemonics as listed in B200_PROFILING.md:
UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store,
UTMAPF = tensormap prefetch, SYNCS = mbarrier ops).

  python tools/sass_evidence.py > profiles/r02_sass_evidence.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'rigl_b200', 'librigl_b200.so')
PATTERNS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'SYNCS', 'RED', 'ATOM']


def main():
  out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
  if out.returncode != 0:
    sys.exit(out.stderr)
  dem = {}
  counts = collections.OrderedDict()
  variants = collections.defaultdict(set)
  cur = None
  for line in out.stdout.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
      cur = m.group(1)
      counts[cur] = collections.Counter()
      continue
    if cur is None:
      continue
    m = re.search(r'^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
      op = m.group(1)
      for p in PATTERNS:
        if op.startswith(p):
          counts[cur][p] += 1
          if p in ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'UTCBAR'):
            variants[cur].add(op)
  names = list(counts)
  d = subprocess.run(['c++filt'] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines()
  for n, x in zip(names, d):
    dem[n] = x
  arch = subprocess.run(['cuobjdump', '-lelf', LIB], stdout=subprocess.PIPE, text=True).stdout.strip().splitlines()
  print('# r02: SASS evidence for the shipped `rigl_b200/librigl_b200.so`')
  print()
  print('`python tools/sass_evidence.py` (cuobjdump -sass, CUDA 12.9).  ELF images: ' + ', '.join(a.split()[-1] for a in arch))
  print()
  print('| kernel | ' + ' | '.join(PATTERNS) + ' | tensor / TMA instruction variants |')
  print('|---|' + '---|' * len(PATTERNS) + '---|')
  for n in names:
    c = counts[n]
    if not any(c[p] for p in PATTERNS):
      continue
    short = re.sub(r'^(void )?rigl::', '', dem[n])
    short = re.sub(r'\(.*$', '', short)
    print('| `%s` | %s | %s |' % (short[:70], ' | '.join(str(c[p]) for p in PATTERNS),
                                  ' '.join(sorted(variants[n]))[:150]))
  tc = [n for n in names if counts[n]['UTCHMMA']]
  print()
  print('%d kernels issue tcgen05.mma (UTCHMMA); %d use TMA tensor loads; %d read accumulators from TMEM (LDTM).' % (
      len(tc), sum(1 for n in names if counts[n]['UTMALDG']), sum(1 for n in names if counts[n]['LDTM'])))


if __name__ == '__main__':
  main()
