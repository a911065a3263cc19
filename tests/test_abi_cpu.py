"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares, the ctypes table covers the header, and the product refuses CPU tensors
(no fallback).  No kernel runs here."""

import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT
from spml_b200 import _lib


def header_functions():
  text = open(os.path.join(ROOT, 'include', 'spml_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(spml_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported_and_bound():
  names = header_functions()
  assert len(names) >= 20
  lib = _lib.load()
  for n in names:
    assert hasattr(lib, n), 'not exported: ' + n
  assert sorted(_lib.SIGNATURES) == names
  out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True,
                       text=True).stdout
  exported = set(re.findall(r' T (spml_[a-z0-9_]+)', out))
  assert set(names) <= exported


def test_abi_version_and_error_text():
  lib = _lib.load()
  assert lib.spml_abi_version() == _lib.ABI_VERSION
  assert lib.spml_unique_inverse(None, None, -1, None, 0, None, None, None, None, None, None,
                                 0, None) == -1
  assert b'unique_inverse' in lib.spml_last_error()


def test_struct_layouts_match_c():
  """sizeof the three argument structs as a C compiler sees them == the ctypes mirrors ==
  what the library reports (spml_sizeof_struct)."""
  src = ('#include "spml_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu",'
         'sizeof(spml_segsort_desc),sizeof(spml_cluster_args),sizeof(spml_head_args));}')
  exe = '/tmp/spml_struct_sizes'
  r = subprocess.run(['gcc', '-x', 'c', '-', '-I', os.path.join(ROOT, 'include'), '-o', exe],
                     input=src, text=True, capture_output=True)
  assert r.returncode == 0, r.stderr
  import ctypes
  sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
  lib = _lib.load()
  for which, (size, struct) in enumerate(zip(sizes, (_lib.SegsortDesc, _lib.ClusterArgs,
                                                     _lib.HeadArgs))):
    assert size == ctypes.sizeof(struct) == lib.spml_sizeof_struct(which), struct.__name__


def test_no_cpu_fallback():
  from spml_b200 import general_common, segsort_common, segsort_loss
  with pytest.raises(RuntimeError, match='CUDA'):
    general_common.normalize_embedding(torch.zeros(2, 3))
  with pytest.raises(RuntimeError, match='CUDA'):
    segsort_common.segment_by_kmeans(torch.zeros(1, 4, 8, 8))
  with pytest.raises(RuntimeError, match='CUDA'):
    segsort_loss.SegSortLoss(10)(torch.zeros(4, 3), torch.zeros(4, dtype=torch.long),
                                 torch.zeros(4, dtype=torch.long), torch.zeros(2, 3),
                                 torch.zeros(2, dtype=torch.long))


def test_product_does_not_import_oracle():
  pkg = os.path.join(ROOT, 'spml_b200')
  for fn in os.listdir(pkg):
    if fn.endswith('.py'):
      assert 'oracle' not in open(os.path.join(pkg, fn)).read(), fn


def integration_snippet():
  text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
  blocks = re.findall(r'```python\n(.*?)```', text, flags=re.S)
  return next(b for b in blocks if 'spml_segment_prototypes_fwd' in b)


def test_integration_snippet_matches_the_header():
  """The binding INTEGRATION.md shows a maintainer has as many arguments, in the same order of
  pointer / integer / float kinds, as include/spml_b200.h declares."""
  code = integration_snippet()
  argtypes = re.search(r'spml_segment_prototypes_fwd\.argtypes = \[(.*?)\]', code).group(1)
  kinds = [a.strip() for a in argtypes.split(',')]
  header = open(os.path.join(ROOT, 'include', 'spml_b200.h')).read()
  decl = re.search(r'int spml_segment_prototypes_fwd\((.*?)\);', header, flags=re.S).group(1)
  params = [p.strip() for p in decl.replace('\n', ' ').split(',')]
  assert len(kinds) == len(params) == 12

  def kind(p):
    if '*' in p:
      return '_vp'
    return {'int64_t': '_i64', 'int': '_i32', 'float': '_f32', 'size_t': '_sz'}[p.split()[0]]
  assert kinds == [kind(p) for p in params]
  assert [c_.__name__ for c_ in _lib.SIGNATURES['spml_segment_prototypes_fwd'][1]] == [
      {'_vp': 'c_void_p', '_i64': 'c_long', '_i32': 'c_int', '_f32': 'c_float',
       '_sz': 'c_ulong'}[k] for k in kinds]
