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
# SPML_B200_LIB: another build of the same library (e.g. a -DSPML_KM_TRACE build for scripts/)
LIB_PATH = os.environ.get('SPML_B200_LIB') or os.path.join(_HERE, 'libspml_b200.so')
ABI_VERSION = 4

c_i32, c_i64, c_f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
c_vp, c_sz = ctypes.c_void_p, ctypes.c_size_t

MODE_CLASS, MODE_TAGS = 0, 1
REDUCE_MEAN, REDUCE_GROUP_MEAN, REDUCE_SUM = 0, 1, 2
MAX_DIM, MAX_TOPK, MAX_BANK = 136, 32, 8


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


class ClusterArgs(ctypes.Structure):
  """struct spml_cluster_args (include/spml_b200.h)."""
  _fields_ = [
      ('emb', c_vp), ('loc', c_vp), ('loc_batch_stride', c_i64),
      ('labels', c_vp), ('sem', c_vp), ('inst', c_vp),
      ('label_divisor', c_i64), ('semantic_ignore', c_i64),
      ('ignore_index_dev', c_vp), ('ignore_index', c_i64),
      ('seeds', c_vp), ('seed_batch_stride', c_i64), ('k_per_image', c_vp),
      ('batch_index_offset', c_i64),
      ('batch', c_i32), ('dim', c_i32), ('n', c_i32), ('loc_ch', c_i32),
      ('num_clusters', c_i32), ('iterations', c_i32), ('has_ignore', c_i32), ('reserved', c_i32),
      ('eps', c_f32), ('reserved_f', c_f32),
      ('e', c_vp), ('el', c_vp), ('nx', c_vp), ('nc', c_vp),
      ('labels_out', c_vp), ('batch_out', c_vp), ('segment_ids', c_vp),
      ('sem_out', c_vp), ('inst_out', c_vp),
      ('dst', c_vp), ('img_off', c_vp), ('kmeans_labels', c_vp), ('seed_out', c_vp),
      ('num_segments', c_vp),
      ('counts_host', c_vp), ('status', c_vp), ('counts_dev', c_vp),
  ]


class HeadArgs(ctypes.Structure):
  """struct spml_head_args (include/spml_b200.h)."""
  _fields_ = [
      ('e', c_vp), ('el', c_vp), ('seg', c_vp), ('bid', c_vp), ('sem', c_vp), ('inst', c_vp),
      ('protos', c_vp), ('protos_loc', c_vp), ('psem', c_vp), ('pinst', c_vp), ('pbid', c_vp),
      ('img_tags', c_vp), ('ptags', c_vp),
      ('img_tags_ld', c_i64), ('ptags_ld', c_i64), ('tag_rows', c_i64),
      ('n', c_i64), ('m', c_i64), ('num_classes', c_i64), ('max_rows_per_group', c_i64),
      ('dim', c_i32), ('dim_loc', c_i32), ('tag_col0', c_i32), ('tag_col1', c_i32),
      ('num_bank', c_i32), ('max_groups', c_i32), ('enable', ctypes.c_uint32),
      ('nn_tags', c_i32), ('img_sim_on_plain', c_i32), ('wide_tags', c_i32),
      ('kappa_ann', c_f32), ('kappa_occ', c_f32), ('kappa_sim', c_f32),
      ('weight_ann', c_f32), ('weight_occ', c_f32), ('weight_sim', c_f32),
      ('nn_threshold', c_f32), ('eps', c_f32),
      ('status', c_vp),
      ('bank_protos', c_vp * MAX_BANK), ('bank_protos_loc', c_vp * MAX_BANK),
      ('bank_psem', c_vp * MAX_BANK), ('bank_pbid', c_vp * MAX_BANK),
      ('bank_tags', c_vp * MAX_BANK), ('bank_tags_ld', c_i64 * MAX_BANK),
      ('bank_m', c_i64 * MAX_BANK),
  ]


# name -> (restype, argtypes); every symbol include/spml_b200.h declares.
SIGNATURES = {
    'spml_last_error': (ctypes.c_char_p, []),
    'spml_abi_version': (ctypes.c_int, []),
    'spml_debug_launch_count': (ctypes.c_uint64, []),
    'spml_debug_kmeans_path': (ctypes.c_int, [c_i32, c_i32, c_i32, c_i32]),
    'spml_sizeof_struct': (c_sz, [ctypes.c_int]),
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
    'spml_segsort_bwd_rows': (ctypes.c_int, [ctypes.POINTER(SegsortDesc), c_vp, c_vp, c_f32, c_vp,
                                             c_i64, c_vp, c_i64, c_vp, c_sz, c_vp]),
    'spml_pack_tags': (ctypes.c_int, [c_vp, c_i64, c_i32, c_i64, c_vp, c_vp]),
    'spml_topk_ranking': (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp,
                                         c_i32, c_vp, c_vp, c_vp, c_vp]),
    'spml_nn_multiset_labels_workspace_bytes': (c_sz, [c_i64, c_i32]),
    'spml_nn_multiset_labels': (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp,
                                               c_i32, c_i32, c_f32, c_vp, c_vp, c_vp, c_sz, c_vp]),
    'spml_segment_by_kmeans_workspace_bytes': (c_sz, [c_i32, c_i32, c_i32, c_i32, c_i32]),
    'spml_segment_by_kmeans': (ctypes.c_int, [ctypes.POINTER(ClusterArgs), c_vp, c_sz, c_vp]),
    'spml_gather_prototypes_workspace_bytes': (c_sz, [c_i64, c_i32, c_i32]),
    'spml_gather_prototypes_fwd': (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp,
                                                  c_vp, c_vp, c_i64, c_f32, c_vp, c_vp, c_vp,
                                                  c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz,
                                                  c_vp]),
    'spml_gather_prototypes_bwd': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64,
                                                  c_i32, c_i32, c_i64, c_f32, c_vp, c_vp, c_vp]),
    'spml_head_workspace_bytes': (c_sz, [ctypes.POINTER(HeadArgs)]),
    'spml_head_fwd': (ctypes.c_int, [ctypes.POINTER(HeadArgs), c_vp, c_sz, c_vp, c_vp]),
    'spml_head_bwd': (ctypes.c_int, [ctypes.POINTER(HeadArgs), c_vp, c_sz, c_vp, c_vp, c_vp, c_vp,
                                     c_vp, c_vp, c_vp, c_vp]),
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
    for which, struct in enumerate((SegsortDesc, ClusterArgs, HeadArgs)):
      if lib.spml_sizeof_struct(which) != ctypes.sizeof(struct):
        raise RuntimeError('spml_b200: layout of %s differs from the library (%d vs %d bytes)'
                           % (struct.__name__, ctypes.sizeof(struct),
                              lib.spml_sizeof_struct(which)))
    _lib = lib
  return _lib


# bench.py sets this to a list to collect (entry point, start event, end event,
# kernels launched) for every call; None (the default) costs nothing.
PROFILE = None
# label appended to the entry-point name of profiled calls (e.g. ':sem_occ'); per host thread,
# because the reference drives every GPU from its own Python thread
_tls = threading.local()


def set_profile_tag(tag):
  _tls.tag = tag


def call(name, *args):
  """Invokes an int-returning entry point and raises on a non-zero status.  The last
  argument is the stream (`stream_of(tensor)`), which remembers the tensor's device: the
  call runs with that device current (kernels launch on the current device; the stream and
  the pointers belong to the tensor's), whatever the calling thread's current device is."""
  lib = load()
  device = getattr(args[-1], 'device', None) if args else None
  if device is not None and device.index != torch.cuda.current_device():
    with torch.cuda.device(device):
      return _call(lib, name, args)
  return _call(lib, name, args)


_fn_cache = {}


def _call(lib, name, args):
  if PROFILE is not None:
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    before = lib.spml_debug_launch_count()
    start.record()
    rc = getattr(lib, name)(*args)
    end.record()
    PROFILE.append((name + getattr(_tls, 'tag', ''), start, end,
                    lib.spml_debug_launch_count() - before))
  else:
    rc = getattr(lib, name)(*args)
  if rc != 0:
    raise RuntimeError('%s failed (%d): %s'
                       % (name, rc, lib.spml_last_error().decode('utf-8', 'replace')))


def launch_count():
  return int(load().spml_debug_launch_count())


def ptr(t):
  """Device address of a CUDA tensor as an int (None -> NULL); ctypes converts it."""
  if t is None:
    return None
  if not t.is_cuda:
    raise RuntimeError('spml_b200 only runs on CUDA tensors (got a %s tensor); '
                       'there is no CPU path' % t.device.type)
  return t.data_ptr()


class _Stream(ctypes.c_void_p):
  """cudaStream_t that remembers its device (see call)."""
  device = None


def stream_of(t):
  dev = t.device
  s = _Stream(torch._C._cuda_getCurrentRawStream(dev.index))
  s.device = dev
  return s


def addr(t, offset_bytes=0):
  """Raw device address (int) of a CUDA tensor plus a byte offset, for struct fields."""
  return t.data_ptr() + offset_bytes
