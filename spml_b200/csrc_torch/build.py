"""Builds spml_b200/_C.so: the ATen-side binding of the stage-group C ABI (binding.cpp).

    python spml_b200/csrc_torch/build.py

A plain g++ invocation with the include / library paths of the installed torch (no ninja, no
JIT cache: the result lives in-tree so that it travels to the GPU box with the snapshot).
Host code only; the CUDA kernels are in libspml_b200.so, which this links against.
"""

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)


def build(force=False):
  import torch
  from torch.utils import cpp_extension as ext
  src = os.path.join(HERE, 'binding.cpp')
  out = os.path.join(PKG, '_C.so')
  lib = os.path.join(PKG, 'libspml_b200.so')
  header = os.path.join(os.path.dirname(PKG), 'include', 'spml_b200.h')
  newest = max(os.path.getmtime(p) for p in (src, header))
  if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
    return out
  if not os.path.exists(lib):
    raise RuntimeError('build libspml_b200.so first (make -C spml_b200/csrc)')
  torch_lib = os.path.join(os.path.dirname(torch.__file__), 'lib')
  cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-Wno-deprecated-declarations',
         '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI),
         '-DTORCH_EXTENSION_NAME=_C', '-DTORCH_API_INCLUDE_EXTENSION_H']
  for inc in ext.include_paths() + [sysconfig.get_paths()['include'], '/usr/local/cuda/include']:
    cmd += ['-isystem', inc]
  cmd += [src, '-o', out, '-L', torch_lib, '-L', PKG, '-lspml_b200', '-lc10', '-lc10_cuda',
          '-ltorch_cpu', '-ltorch_cuda', '-ltorch', '-ltorch_python',
          '-Wl,-rpath,$ORIGIN', '-Wl,-rpath,' + torch_lib]
  subprocess.run(cmd, check=True)
  return out


if __name__ == '__main__':
  print(build(force='--force' in sys.argv))
