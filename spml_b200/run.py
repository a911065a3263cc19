"""Launcher that runs an UNCHANGED reference script on libspml_b200.so:

    PYTHONPATH=/path/to/SPML python -m spml_b200.run pyscripts/train/train.py --cfg_path ...

It rebinds the hot-path symbols of the imported `spml.*` modules (spml_b200.install) and
then executes the script as `__main__` with the remaining command line, so the reference
file stays byte-for-byte what it is.  `--spml-b200-lenient` (before the script path)
skips reference modules that cannot be imported instead of failing.
"""

from __future__ import annotations

import runpy
import sys


def main(argv=None):
  argv = list(sys.argv[1:] if argv is None else argv)
  strict = True
  if argv and argv[0] == '--spml-b200-lenient':
    strict = False
    argv = argv[1:]
  if not argv:
    raise SystemExit('usage: python -m spml_b200.run [--spml-b200-lenient] SCRIPT.py [args...]')
  import importlib
  rebound = importlib.import_module('spml_b200.install').install(strict=strict)
  sys.stderr.write('spml_b200: rebound %d reference symbols\n' % len(rebound))
  script = argv[0]
  sys.argv = argv
  runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
  main()
