"""Puts an importable copy of the reference (twke18/SPML) under baseline/_ref/ so that it
travels to the GPU box with the snapshot (baseline/_ref is git-ignored, not gpurun-ignored).

    python scripts/install_reference.py            # build container only: needs /root/reference

The reference has no setup.py / pyproject (nothing for pip to install), so this is the
recipe of SURVEY.md Appendix A: copy the `spml` and `lib` packages and apply the one-line CPU
shim (`tensor.device.index` is None on the CPU, spml/utils/segsort/common.py:376).  Nothing
under baseline/_ref is part of the product or of the git history; it is used by
  * tests/test_gpu_reference_dropin.py: the reference's OWN code on cuda:0, with and without
    spml_b200.install(), same inputs, results compared;
  * bench.py --impl reference: the reference's own CPU implementation as the timed baseline
    (cpu_baseline.kind = "reference").
"""

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')


def main():
  if not os.path.isdir(os.path.join(SRC, 'spml')):
    sys.exit('no reference checkout at %s' % SRC)
  if os.path.isdir(DST):
    shutil.rmtree(DST)
  os.makedirs(DST)
  for pkg in ('spml', 'lib'):
    shutil.copytree(os.path.join(SRC, pkg), os.path.join(DST, pkg),
                    ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
  path = os.path.join(DST, 'spml', 'utils', 'segsort', 'common.py')
  src = open(path).read()
  needle = 'gpu_id = cur_cluster_indices.device.index\n'
  assert src.count(needle) == 1
  open(path, 'w').write(src.replace(needle, 'gpu_id = cur_cluster_indices.device.index or 0\n'))
  n = sum(len(files) for _, _, files in os.walk(DST))
  print('installed %d files under %s' % (n, os.path.relpath(DST, ROOT)))


if __name__ == '__main__':
  main()
