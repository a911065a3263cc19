"""ctypes binding of libspml_b200.so (the C ABI declared in include/spml_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared
object is missing, or a tensor is not a CUDA tensor, the call fails loudly.
"""

from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libspml_b200.so')
ABI_VERSION = 2

c_i32, c_i64, c_f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
c_vp, c_sz = ctypes.c_void_p, ctypes.c_size_t

MODE_CLASS, MODE_TAGS = 0, 1
REDUCE_MEAN, REDUCE_GROUP_MEAN, REDUCE_SUM = 0, 1, 2
MAX_DIM, MAX_TOPK = 136, 32


class SegsortDesc(ctypes.Structure):
  """struct spml_segsort_desc (include/spml_b200.h)."""
  _fields_ = [
      ('emb', c_vp), ('ld_emb', c_i64), ('dim', c_i32), ('num_groups', c_i32),
      ('row_index', c_vp), ('group_off', c_vp), ('col_off', c_vp),
      ('n_rows', c_i64), ('max_rows_per_group', c_i64),
      ('pix_code', c_vp), ('seg', c_vp),
      ('protos', c_vp), ('ld_protos', c_i64), ('m', c_i64),
      ('proto_code', c_vp), ('proto_valid', c_vp),
      ('kappa', c_f32), ('mode', c_i32), ('reduction', c_i32), ('reserved', c_i32),
  ]


# name -> (restype, argtypes); every symbol include/spml_b200.h declares.
SIGNATURES = {
    'spml_last_error': (ctypes.c_char_p, []),
    'spml_abi_version': (ctypes.c_int, []),
    'spml_debug_launch_count': (ctypes.c_uint64, []),
    'spml_normalize_rows_fwd': (ctypes.c_int, [c_vp, c_i64, c_i32, c_f32, c_vp, c_vp, c_vp]),
    'spml_normalize_rows_bwd': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    'spml_valid_scan_workspace_bytes': (c_sz, [c_i32, c_i32]),
    'spml_valid_scan': (ctypes.c_int, [c_vp, c_i32, c_i64, c_vp, c_i32, c_i32, c_vp, c_vp,
                                       c_vp, c_vp, c_sz, c_vp]),
    'spml_normalize_pack_fwd': (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_i64,
                                               c_vp, c_i32, c_i32, c_i32, c_i64, c_f32,
                                               c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'spml_normalize_pack_bwd': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                               c_i32, c_i32, c_i32, c_i32, c_f32, c_vp, c_vp]),
    'spml_kmeans_workspace_bytes': (c_sz, [c_i32, c_i32, c_i32, c_i32]),
    'spml_kmeans': (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp,
                                   c_vp, c_vp, c_vp, c_sz, c_vp]),
    'spml_nearest_prototype': (ctypes.c_int, [c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_vp]),
    'spml_unique_workspace_bytes': (c_sz, [c_i64]),
    'spml_unique_inverse': (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp,
                                           c_vp, c_vp, c_vp, c_sz, c_vp]),
    'spml_segment_prototypes_workspace_bytes': (c_sz, [c_i64, c_i32]),
    'spml_segment_prototypes_fwd': (ctypes.c_int, [c_vp, c_i64, c_vp, c_i32, c_vp, c_i64, c_f32,
                                                   c_vp, c_vp, c_vp, c_sz, c_vp]),
    'spml_segment_prototypes_bwd': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i32,
                                                   c_i64, c_f32, c_f32, c_vp, c_vp]),
    'spml_segment_labels': (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64,
                                           c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                           c_vp]),
    'spml_segsort_workspace_bytes': (c_sz, [ctypes.POINTER(SegsortDesc)]),
    'spml_segsort_fwd': (ctypes.c_int, [ctypes.POINTER(SegsortDesc), c_vp, c_vp, c_vp, c_vp,
                                        c_sz, c_vp]),
    'spml_segsort_bwd': (ctypes.c_int, [ctypes.POINTER(SegsortDesc), c_vp, c_vp, c_f32, c_vp,
                                        c_i64, c_vp, c_vp, c_sz, c_vp]),
    'spml_pack_tags': (ctypes.c_int, [c_vp, c_i64, c_i32, c_i64, c_vp, c_vp]),
    'spml_topk_ranking': (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp,
                                         c_i32, c_vp, c_vp, c_vp, c_vp]),
}

_lock = threading.Lock()
_lib = None


def load():
  """Loads the shared library once and checks every declared symbol."""
  global _lib
  if _lib is not None:
    return _lib
  with _lock:
    if _lib is not None:
      return _lib
    if not os.path.exists(LIB_PATH):
      raise RuntimeError(
          'spml_b200: %s is missing. Build it with `make -C spml_b200/csrc` (or '
          '`python -c "import __graft_entry__ as g; g.build()"`). There is no '
          'CPU / PyTorch fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
      fn = getattr(lib, name)      # AttributeError if the symbol is not exported
      fn.restype, fn.argtypes = restype, argtypes
    got = lib.spml_abi_version()
    if got != ABI_VERSION:
      raise RuntimeError('spml_b200: ABI version %d, expected %d (stale build?)'
                         % (got, ABI_VERSION))
    _lib = lib
  return _lib


# bench.py sets this to a list to collect (entry point, start event, end event,
# kernels launched) for every call; None (the default) costs nothing.
PROFILE = None
PROFILE_TAG = ''     # appended to the entry-point name of profiled calls (e.g. ':sem_occ')


def call(name, *args):
  """Invokes an int-returning entry point and raises on a non-zero status."""
  lib = load()
  if PROFILE is not None:
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    before = lib.spml_debug_launch_count()
    start.record()
    rc = getattr(lib, name)(*args)
    end.record()
    PROFILE.append((name + PROFILE_TAG, start, end, lib.spml_debug_launch_count() - before))
  else:
    rc = getattr(lib, name)(*args)
  if rc != 0:
    raise RuntimeError('%s failed (%d): %s'
                       % (name, rc, lib.spml_last_error().decode('utf-8', 'replace')))


def launch_count():
  return int(load().spml_debug_launch_count())


def ptr(t):
  """Device pointer of a CUDA tensor (None -> NULL)."""
  if t is None:
    return None
  if not t.is_cuda:
    raise RuntimeError('spml_b200 only runs on CUDA tensors (got a %s tensor); '
                       'there is no CPU path' % t.device.type)
  return ctypes.c_void_p(t.data_ptr())


def stream_of(t):
  return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
