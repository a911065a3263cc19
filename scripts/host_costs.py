"""Host-side cost of the driver calls the stage-group entries make (ns per call)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spml_b200 import _lib  # noqa: E402

torch.zeros(1, device='cuda')
lib = ctypes.CDLL(_lib.library_path()) if hasattr(_lib, 'library_path') else _lib.load()
fn = lib.spml_debug_host_costs
fn.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_void_p]
out = (ctypes.c_double * 5)()
for _ in range(2):
  rc = fn(out, torch.cuda.current_stream().cuda_stream)
assert rc == 0
names = ('cuTensorMapEncodeTiled', 'cudaFuncSetAttribute', 'kernel launch (empty, queue not full)',
         'cudaMemsetAsync 16 B', 'cudaEventRecord + cudaStreamWaitEvent')
for n, v in zip(names, out):
  print('%-45s %8.0f ns' % (n, v))
