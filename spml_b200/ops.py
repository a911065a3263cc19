"""Tensor-level wrappers and autograd Functions over the C ABI (spml_b200/_lib.py).

PyTorch is used here for device memory (the caching allocator owns every buffer
the library reads or writes), the current CUDA stream and autograd bookkeeping.
All arithmetic of the path happens inside libspml_b200.so.
"""

from __future__ import annotations

import ctypes
import os

import torch

from . import _lib
from ._lib import call, ptr, stream_of

EPS = 1e-12


def _f32c(t, name):
  if not t.is_cuda:
    raise RuntimeError('%s: spml_b200 needs CUDA tensors, got %s (no CPU path)'
                       % (name, t.device))
  if t.dtype != torch.float32:
    raise TypeError('%s: expected float32, got %s' % (name, t.dtype))
  return t if t.is_contiguous() else t.contiguous()


def _i64c(t, name):
  if not t.is_cuda:
    raise RuntimeError('%s: spml_b200 needs CUDA tensors, got %s (no CPU path)'
                       % (name, t.device))
  if t.dtype != torch.int64:
    t = t.long()
  return t if t.is_contiguous() else t.contiguous()


_size_cache = {}


def _cached_size(fn_name, *key):
  """Workspace sizes only depend on the shape: one library call per distinct shape."""
  k = (fn_name,) + key
  v = _size_cache.get(k)
  if v is None:
    v = _size_cache[k] = int(getattr(_lib.load(), fn_name)(*key))
  return v


def _workspace(nbytes, device):
  return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------ A1


class NormalizeRows(torch.autograd.Function):
  """normalize_embedding (spml/utils/general/common.py:101-120)."""

  @staticmethod
  def forward(ctx, x, eps):
    shape = x.shape
    x2 = _f32c(x, 'normalize_embedding').view(-1, shape[-1])
    y = torch.empty_like(x2)
    norms = torch.empty(x2.shape[0], dtype=torch.float32, device=x2.device)
    call('spml_normalize_rows_fwd', ptr(x2), x2.shape[0], x2.shape[1], eps, ptr(y), ptr(norms),
         stream_of(x2))
    ctx.save_for_backward(y, norms)
    ctx.shape = shape
    return y.view(shape)

  @staticmethod
  def backward(ctx, dy):
    y, norms = ctx.saved_tensors
    dy2 = _f32c(dy, 'normalize_embedding.backward').view(-1, ctx.shape[-1])
    dx = torch.empty_like(dy2)
    call('spml_normalize_rows_bwd', ptr(dy2), ptr(y), ptr(norms), y.shape[0], y.shape[1],
         ptr(dx), stream_of(y))
    return dx.view(ctx.shape), None


# ------------------------------------------------------------------------------ A8 front


def valid_scan(labels, ignore_index, batch, n, want_src=False):
  """labels: int64 [batch, n] (or None when nothing is ignored).
  Returns dst [batch*n] int32, src (or None), img_off [batch+1] int32."""
  device = labels.device if labels is not None else torch.device('cuda')
  dst = torch.empty(batch * n, dtype=torch.int32, device=device)
  src = torch.empty(batch * n, dtype=torch.int32, device=device) if want_src else None
  img_off = torch.empty(batch + 1, dtype=torch.int32, device=device)
  lib = _lib.load()
  ws = _workspace(lib.spml_valid_scan_workspace_bytes(batch, n), device)
  has_ignore = ignore_index is not None
  ignore_dev = None
  if torch.is_tensor(ignore_index):     # stays on the device: no host sync
    ignore_dev = ignore_index.to(device=device, dtype=torch.int64).reshape(1)
    ignore_index = 0
  call('spml_valid_scan', ptr(labels) if has_ignore else None, int(has_ignore),
       int(ignore_index) if has_ignore else 0, ptr(ignore_dev), batch, n, ptr(dst), ptr(src),
       ptr(img_off), ptr(ws), ws.numel(), stream_of(dst))
  return dst, src, img_off


class NormalizePack(torch.autograd.Function):
  """NCHW embeddings -> packed normalised rows (+ location features), with the
  ignored pixels dropped.  Outputs have the CAPACITY batch*n; the caller narrows
  them to the live row count."""

  @staticmethod
  def forward(ctx, emb, loc, labels, seeds, dst, batch_index_offset):
    emb = _f32c(emb, 'segment_by_kmeans(embeddings)')
    B, D, H, W = emb.shape
    n = H * W
    cap = B * n
    dev = emb.device
    if loc is not None:
      if loc.dim() != 4 or loc.shape[1] != H or loc.shape[2] != W:
        raise ValueError('local_features must be [batch, H, W, C]')
      loc_ch = loc.shape[3]
      if loc.stride(0) == 0 or loc.shape[0] == 1:
        loc_c, loc_bs = _f32c(loc[0], 'local_features'), 0
      else:
        loc_c, loc_bs = _f32c(loc, 'local_features'), n * loc_ch
    else:
      loc_c, loc_ch, loc_bs = None, 0, 0
    if seeds.dim() == 2:      # one [H, W] seed map shared by every image
      seeds_c, seed_bs = _i64c(seeds, 'cluster_indices'), 0
    else:
      seeds_c, seed_bs = _i64c(seeds, 'cluster_indices'), n
    e = torch.empty(cap, D, dtype=torch.float32, device=dev)
    el = torch.empty(cap, D + loc_ch, dtype=torch.float32, device=dev)
    nx = torch.empty(cap, dtype=torch.float32, device=dev)
    nc = torch.empty(cap, dtype=torch.float32, device=dev)
    # rows past the live count are padding: integers read as 0 there (fixed-capacity mode)
    # (one zero-filled buffer, three views: one fill kernel instead of three)
    zeros = torch.zeros(cap * 20, dtype=torch.uint8, device=dev)
    labels_out = zeros[:cap * 8].view(torch.int64)
    batch_out = zeros[cap * 8:cap * 16].view(torch.int64)
    seed_out = zeros[cap * 16:].view(torch.int32)
    call('spml_normalize_pack_fwd', ptr(emb), ptr(loc_c), loc_bs, loc_ch, ptr(labels),
         ptr(seeds_c), seed_bs, ptr(dst), B, D, n, int(batch_index_offset), EPS,
         ptr(e), ptr(el), ptr(nx), ptr(nc), ptr(labels_out), ptr(batch_out), ptr(seed_out),
         stream_of(emb))
    ctx.save_for_backward(e, el, nx, nc, dst)
    ctx.dims = (B, D, loc_ch, H, W)
    ctx.mark_non_differentiable(labels_out, batch_out, seed_out)
    return e, el, labels_out, batch_out, seed_out

  @staticmethod
  def backward(ctx, de, del_, *_):
    e, el, nx, nc, dst = ctx.saved_tensors
    B, D, loc_ch, H, W = ctx.dims
    demb = torch.empty(B, D, H, W, dtype=torch.float32, device=e.device)
    de = _f32c(de, 'd(cluster_embedding)') if de is not None else None
    del_ = _f32c(del_, 'd(cluster_embedding_with_loc)') if del_ is not None else None
    call('spml_normalize_pack_bwd', ptr(de), ptr(del_), ptr(e), ptr(el), ptr(nx), ptr(nc),
         ptr(dst), B, D, loc_ch, H * W, EPS, ptr(demb), stream_of(e))
    return demb, None, None, None, None, None


# ------------------------------------------------------------------------------ A4-A6


def kmeans(x, img_off, batch, max_rows_per_image, num_clusters, iterations, init_labels,
           k_per_image=None, want_i64=True):
  """x [cap, dim] fp32, img_off [batch+1] int32, init_labels [cap] int32.
  Returns (labels int32 [cap], labels int64 [cap] or None)."""
  dim = x.shape[1]
  out = torch.empty(x.shape[0], dtype=torch.int32, device=x.device)
  out64 = torch.empty(x.shape[0], dtype=torch.int64, device=x.device) if want_i64 else None
  lib = _lib.load()
  ws = _workspace(lib.spml_kmeans_workspace_bytes(batch, num_clusters, dim, iterations), x.device)
  call('spml_kmeans', ptr(x), ptr(img_off), batch, int(max_rows_per_image), dim,
       int(num_clusters), ptr(k_per_image), int(iterations), ptr(init_labels), ptr(out),
       ptr(out64), ptr(ws), ws.numel(), stream_of(x))
  return out, out64


def nearest_prototype(x, protos):
  x = _f32c(x, 'find_nearest_prototypes(embeddings)')
  protos = _f32c(protos, 'find_nearest_prototypes(prototypes)')
  x2 = x.view(-1, protos.shape[-1])
  out = torch.empty(x2.shape[0], dtype=torch.int64, device=x.device)
  call('spml_nearest_prototype', ptr(x2), x2.shape[0], x2.shape[1], ptr(protos),
       protos.shape[0], ptr(out), stream_of(x2))
  return out


# ------------------------------------------------------------------------------ A7


def unique_inverse(lo, hi=None, bound=0, n_dev=None, want_keys=True):
  """torch.unique(hi * bound + lo, return_inverse=True) on the device.

  Returns (inverse [n] int64, uniq_hi, uniq_lo (capacity n), count int32[1],
  bound_used int64[1]).  Nothing is synchronised; the caller reads `count`."""
  lo = _i64c(lo, 'unique(keys)').view(-1)
  hi = _i64c(hi, 'unique(keys)').view(-1) if hi is not None else None
  n = lo.numel()
  dev = lo.device
  inverse = (torch.zeros if n_dev is not None else torch.empty)(n, dtype=torch.int64, device=dev)
  uniq_hi = torch.empty(n, dtype=torch.int64, device=dev) if (want_keys and hi is not None) else None
  uniq_lo = torch.empty(n, dtype=torch.int64, device=dev) if want_keys else None
  count = torch.empty(1, dtype=torch.int32, device=dev)
  bound_out = torch.empty(1, dtype=torch.int64, device=dev)
  lib = _lib.load()
  ws = _workspace(lib.spml_unique_workspace_bytes(n), dev)
  call('spml_unique_inverse', ptr(hi), ptr(lo), n, ptr(n_dev), int(bound), ptr(inverse),
       ptr(uniq_hi), ptr(uniq_lo), ptr(count), ptr(bound_out), ptr(ws), ws.numel(),
       stream_of(lo))
  return inverse, uniq_hi, uniq_lo, count, bound_out


# ------------------------------------------------------------------------------ A4 / B1


class SegmentPrototypes(torch.autograd.Function):
  """calculate_prototypes_from_labels (spml/utils/segsort/common.py:11-41)."""

  @staticmethod
  def forward(ctx, x, seg, m, rows_dev=None):
    x2 = _f32c(x, 'calculate_prototypes_from_labels(embeddings)').view(-1, x.shape[-1])
    seg = _i64c(seg, 'calculate_prototypes_from_labels(labels)').view(-1)
    if seg.numel() != x2.shape[0]:
      raise ValueError('labels and embeddings disagree: %d vs %d rows'
                       % (seg.numel(), x2.shape[0]))
    m = int(m)
    dim = x2.shape[1]
    protos = torch.empty(m, dim, dtype=torch.float32, device=x2.device)
    norms = torch.empty(m, dtype=torch.float32, device=x2.device)
    lib = _lib.load()
    ws = _workspace(lib.spml_segment_prototypes_workspace_bytes(m, dim), x2.device)
    call('spml_segment_prototypes_fwd', ptr(x2), x2.shape[0], ptr(rows_dev), dim, ptr(seg), m,
         EPS, ptr(protos), ptr(norms), ptr(ws), ws.numel(), stream_of(x2))
    ctx.save_for_backward(protos, norms, seg)
    ctx.rows_dev = rows_dev
    ctx.xshape = x.shape
    return protos

  @staticmethod
  def backward(ctx, dp):
    protos, norms, seg = ctx.saved_tensors
    dp = _f32c(dp, 'd(prototypes)')
    rows, dim = seg.numel(), protos.shape[1]
    alloc = torch.zeros if ctx.rows_dev is not None else torch.empty
    dx = alloc(rows, dim, dtype=torch.float32, device=protos.device)
    call('spml_segment_prototypes_bwd', ptr(dp), ptr(protos), ptr(norms), ptr(seg), rows,
         ptr(ctx.rows_dev), dim, protos.shape[0], EPS, 0.0, ptr(dx), stream_of(protos))
    return dx.view(ctx.xshape), None, None, None


# ------------------------------------------------------------------------------ C1 / C2


def pack_tags(tags):
  """[rows, cols<=64] int64 0/1 matrix (any row stride) -> int64 bit masks."""
  if tags.dim() != 2:
    raise ValueError('tags must be 2-D')
  if not tags.is_cuda:
    raise RuntimeError('pack_tags: needs a CUDA tensor (no CPU path)')
  if tags.dtype != torch.int64:
    tags = tags.long()
  if tags.stride(1) != 1:
    tags = tags.contiguous()
  out = torch.empty(tags.shape[0], dtype=torch.int64, device=tags.device)
  call('spml_pack_tags', ptr(tags), tags.shape[0], tags.shape[1], tags.stride(0), ptr(out),
       stream_of(tags))
  return out


class SegsortProblem:
  """Everything but the two differentiable operands of one SegSort launch."""

  def __init__(self, pix_code, seg, proto_code, kappa, mode, reduction=_lib.REDUCE_MEAN,
               row_index=None, group_off=None, col_off=None, num_groups=1, n_rows=None,
               max_rows_per_group=None, proto_valid=None, path='auto', name='',
               proto_grad_rows=None):
    self.pix_code = _i64c(pix_code, 'segsort(pixel labels)').view(-1)
    self.seg = _i64c(seg, 'segsort(instance labels)').view(-1)
    self.proto_code = _i64c(proto_code, 'segsort(prototype labels)').view(-1)
    self.kappa, self.mode, self.reduction = float(kappa), int(mode), int(reduction)
    self.row_index, self.group_off, self.col_off = row_index, group_off, col_off
    self.num_groups = int(num_groups)
    self.n_rows = int(n_rows) if n_rows is not None else self.pix_code.numel()
    self.max_rows_per_group = (int(max_rows_per_group) if max_rows_per_group is not None
                               else self.n_rows)
    self.proto_valid = proto_valid
    self.name = name      # label of this problem in bench.py's per-call profile
    # only the first `proto_grad_rows` prototypes need a gradient (the rest: a detached bank)
    self.proto_grad_rows = None if proto_grad_rows is None else int(proto_grad_rows)
    # 'fp32' / 'tc' pin the CUDA-core / tcgen05 kernels (tests compare the two)
    self.path_bits = {'auto': 0, 'fp32': 1, 'tc': 2}[path]
    if proto_valid is not None and proto_valid.dtype != torch.uint8:
      self.proto_valid = proto_valid.to(torch.uint8)

  def desc(self, emb, protos):
    d = _lib.SegsortDesc()
    d.emb, d.ld_emb, d.dim = ptr(emb), emb.stride(0), emb.shape[1]
    d.num_groups = self.num_groups
    d.row_index, d.group_off, d.col_off = ptr(self.row_index), ptr(self.group_off), ptr(self.col_off)
    d.n_rows, d.max_rows_per_group = self.n_rows, self.max_rows_per_group
    d.pix_code, d.seg = ptr(self.pix_code), ptr(self.seg)
    d.protos, d.ld_protos, d.m = ptr(protos), protos.stride(0), protos.shape[0]
    d.proto_code, d.proto_valid = ptr(self.proto_code), ptr(self.proto_valid)
    d.kappa, d.mode, d.reduction, d.reserved = self.kappa, self.mode, self.reduction, self.path_bits
    return d


class SegsortLossFn(torch.autograd.Function):
  """SegSortLoss / SetSegSortLoss forward + hand-written backward
  (spml/utils/segsort/loss.py:15-251)."""

  @staticmethod
  def forward(ctx, emb, protos, problem):
    emb = _f32c(emb, 'segsort(embeddings)')
    protos = _f32c(protos, 'segsort(prototypes)')
    emb = emb.view(-1, emb.shape[-1])
    protos = protos.view(-1, protos.shape[-1])
    if emb.shape[1] != protos.shape[1]:
      raise ValueError('embedding dim %d != prototype dim %d' % (emb.shape[1], protos.shape[1]))
    if problem.proto_code.numel() != protos.shape[0]:
      raise ValueError('prototype labels and prototypes disagree')
    dev = emb.device
    d = problem.desc(emb, protos)
    lib = _lib.load()
    ws = _workspace(lib.spml_segsort_workspace_bytes(ctypes.byref(d)), dev)
    stats = torch.empty(max(problem.n_rows, 1), 3, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    _lib.set_profile_tag(':' + problem.name if problem.name else '')
    call('spml_segsort_fwd', ctypes.byref(d), ptr(stats), None, ptr(loss), ptr(ws), ws.numel(),
         stream_of(emb))
    _lib.set_profile_tag('')
    ctx.save_for_backward(emb, protos, stats)
    ctx.problem = problem
    ctx.workspace = ws          # the bf16 operands prepared by the forward are re-used
    return loss

  @staticmethod
  def backward(ctx, grad_loss):
    emb, protos, stats = ctx.saved_tensors
    problem = ctx.problem
    dev = emb.device
    d = problem.desc(emb, protos)
    grad_loss = grad_loss.to(torch.float32).contiguous()
    need_e, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    demb = dprotos = None
    if need_e:
      # rows outside the problem (row_index subsets) must read as zero
      demb = (torch.zeros_like(emb) if problem.row_index is not None or
              problem.group_off is not None else torch.empty_like(emb))
    rows = protos.shape[0]
    if need_p:
      if problem.proto_grad_rows is not None and problem.proto_grad_rows < rows:
        rows = problem.proto_grad_rows
        dprotos = torch.zeros_like(protos)       # the rows behind stay zero
      else:
        dprotos = torch.empty_like(protos)
    if need_e or need_p:
      ws = ctx.workspace
      d.reserved |= 4
      _lib.set_profile_tag(':' + problem.name if problem.name else '')
      call('spml_segsort_bwd_rows', ctypes.byref(d), ptr(stats), ptr(grad_loss), 0.0, ptr(demb),
           emb.shape[1], ptr(dprotos), rows, ptr(ws), ws.numel(), stream_of(emb))
      _lib.set_profile_tag('')
    return demb, dprotos, None


# ------------------------------------------------------------------------------ C3


def topk_ranking(q, qlab, p, plab, k, qvalid=None, pvalid=None):
  q = _f32c(q, 'top_k_ranking(embeddings)')
  p = _f32c(p, 'top_k_ranking(prototypes)')
  q2, p2 = q.view(-1, q.shape[-1]), p.view(-1, p.shape[-1])
  qlab, plab = _i64c(qlab, 'top_k_ranking(labels)').view(-1), _i64c(plab, 'top_k_ranking').view(-1)
  nq = q2.shape[0]
  labels = torch.zeros(nq, k, dtype=torch.int64, device=q.device)
  hits = torch.empty(2, dtype=torch.int32, device=q.device)
  call('spml_topk_ranking', ptr(q2), nq, ptr(p2), p2.shape[0], q2.shape[1], ptr(qlab),
       ptr(plab), ptr(qvalid), ptr(pvalid), int(k), ptr(labels), None, ptr(hits), stream_of(q2))
  # mean over (queries that took part) x k; 0 / 0 = nan like torch.mean of an empty tensor
  acc = hits[0].to(torch.float32) / (hits[1].to(torch.float32) * float(k))
  return acc, labels


def segment_labels(labels, batch, seg, rows_dev, divisor, num_classes, m_cap, dead_label,
                   overflow):
  """spml_segment_labels: per-pixel (sem, inst, keep) and per-segment (sem, inst, batch, live)
  for fixed-capacity buffers.  Returns 7 tensors."""
  cap, dev = labels.shape[0], labels.device
  sem = torch.empty(cap, dtype=torch.int64, device=dev)
  inst = torch.empty(cap, dtype=torch.int64, device=dev)
  keep = torch.empty(cap, dtype=torch.int64, device=dev)
  p_sem = torch.empty(m_cap, dtype=torch.int64, device=dev)
  p_inst = torch.empty(m_cap, dtype=torch.int64, device=dev)
  p_batch = torch.empty(m_cap, dtype=torch.int64, device=dev)
  p_live = torch.empty(m_cap, dtype=torch.uint8, device=dev)
  call('spml_segment_labels', ptr(labels), ptr(batch), ptr(seg), cap, ptr(rows_dev), int(divisor),
       int(num_classes), int(m_cap), int(dead_label), ptr(sem), ptr(inst), ptr(keep), ptr(p_sem),
       ptr(p_inst), ptr(p_batch), ptr(p_live), ptr(overflow), stream_of(labels))
  return sem, inst, keep, p_sem, p_inst, p_batch, p_live


# ------------------------------------------------------------------------------ f3


def nn_multiset_labels(q, p, plab, qgroup, pgroup, num_classes, top_k, threshold):
  """spml_nn_multiset_labels -> [rows, num_classes] int64 multi-hot tags."""
  q = _f32c(q, 'nearest-neighbour tags(embeddings)')
  p = _f32c(p, 'nearest-neighbour tags(prototypes)')
  q2 = q.view(-1, q.shape[-1])
  p2 = p.view(-1, q2.shape[1])
  plab = _i64c(plab, 'nearest-neighbour tags(labels)').view(-1)
  qgroup = _i64c(qgroup, 'nearest-neighbour tags').view(-1)
  pgroup = _i64c(pgroup, 'nearest-neighbour tags').view(-1)
  nq = q2.shape[0]
  tags = torch.empty(nq, int(num_classes), dtype=torch.int64, device=q.device)
  lib = _lib.load()
  ws = _workspace(lib.spml_nn_multiset_labels_workspace_bytes(nq, int(top_k)), q.device)
  call('spml_nn_multiset_labels', ptr(q2), nq, ptr(p2), p2.shape[0], q2.shape[1], ptr(plab),
       ptr(qgroup), ptr(pgroup), int(num_classes), int(top_k), float(threshold), ptr(tags), None,
       ptr(ws), ws.numel(), stream_of(q2))
  return tags


# ------------------------------------------------------------------------------ stage groups
#
# One C-ABI call per stage of the training step (csrc/pipeline.cu).  Each device has a small
# status word the kernels OR error bits into; it is read together with the next step's row /
# segment counts (the one host synchronisation of a step), so a violated precondition raises
# one step late instead of costing a synchronisation of its own.

STATUS_IMPURE_SEGMENT, STATUS_BAD_SEGMENT_ID, STATUS_TOO_MANY_IMAGES = 1, 2, 4
_STATUS_TEXT = {
    STATUS_IMPURE_SEGMENT: 'gather_clustering_and_update_prototypes: pixels of one segment carry '
                           'different semantic / instance / batch labels (the ids were not '
                           'produced from these labels)',
    STATUS_BAD_SEGMENT_ID: 'a cluster index lies outside [0, number of prototypes)',
    STATUS_TOO_MANY_IMAGES: 'img_sim: the pixels span more images than the batch size given',
}
_status_words = {}


def status_word(device):
  """int32[1] device tensor the stage kernels report violated preconditions into."""
  key = device.index if device.index is not None else torch.cuda.current_device()
  if key not in _status_words:
    _status_words[key] = torch.zeros(1, dtype=torch.int32, device=device)
  return _status_words[key]


def raise_on_status(bits, device):
  if bits:
    status_word(device).zero_()
    raise RuntimeError('spml_b200 (reported by an earlier, asynchronous call): ' + '; '.join(
        text for bit, text in _STATUS_TEXT.items() if bits & bit))


def check_status(device=None):
  """Host-synchronising check of the status word (tests, end of a run)."""
  device = torch.device('cuda', torch.cuda.current_device()) if device is None else device
  raise_on_status(int(status_word(device)), device)


# The ATen-side binding (csrc_torch/binding.cpp -> spml_b200/_C.so) does the same bookkeeping
# as the ctypes code below in C++; it is used when it has been built (the build entry point
# does) unless SPML_B200_BINDING=ctypes.  Both paths end in the same library calls.
_C = None
if os.environ.get('SPML_B200_BINDING', '') != 'ctypes':
  try:
    _lib.load()                      # libspml_b200.so first: _C.so links against it
    from . import _C                 # noqa: F811
    if _C.abi_version() != _lib.ABI_VERSION:
      _C = None
  except (ImportError, RuntimeError, OSError):
    _C = None


def binding():
  """'aten' (C++ binding) or 'ctypes'."""
  return 'aten' if _C is not None else 'ctypes'


def _status_error(bits):
  return RuntimeError('spml_b200 (reported by an earlier, asynchronous call): ' + '; '.join(
      text for bit, text in _STATUS_TEXT.items() if bits & bit))


class SegmentByKmeansFn(torch.autograd.Function):
  """segment_by_kmeans (spml/utils/segsort/common.py:270-408) as ONE library call and one
  host read-back (rows kept, segments, status).  Label packing of resnet_deeplab.py:112-117
  happens inside the call when (sem, inst, divisor, semantic_ignore) are given instead of
  `labels`.  `box` (a list) receives the sizes for the caller."""

  @staticmethod
  def forward(ctx, emb, loc, labels, sem, inst, divisor, semantic_ignore, ignore_index, seeds,
              k_per_image, num_k, iterations, batch_index_offset, box):
    if _C is not None:
      ignore_dev = ignore_index if torch.is_tensor(ignore_index) else None
      outs, rows, segments, bits = _C.segment_fwd(
          emb, loc, labels, sem, inst, int(divisor) if divisor else 0,
          int(semantic_ignore) if semantic_ignore is not None else 0,
          ignore_index is not None, 0 if (ignore_dev is not None or ignore_index is None)
          else int(ignore_index), ignore_dev, seeds, k_per_image, int(num_k), int(iterations),
          int(batch_index_offset), status_word(emb.device))
      if bits:
        raise _status_error(bits)
      fbuf, ibuf = outs[-2], outs[-1]
      outs = outs[:-2]
      B, D, H, W = emb.shape
      loc_ch = loc.shape[3] if loc is not None else 0
      ctx.save_for_backward(fbuf, ibuf)
      ctx.dims = (B, D, loc_ch, H, W, (B + 2 + 4 + 3) // 4 * 4)
      ctx.mark_non_differentiable(*outs[2:])
      ctx.set_materialize_grads(False)
      box.append((rows, segments, ibuf))
      return tuple(outs)
    emb = _f32c(emb, 'segment_by_kmeans(embeddings)')
    B, D, H, W = emb.shape
    n = H * W
    cap = B * n
    dev = emb.device
    loc_ch, loc_bs, loc_c = 0, 0, None
    if loc is not None:
      if loc.dim() != 4 or loc.shape[1] != H or loc.shape[2] != W:
        raise ValueError('local_features must be [batch, H, W, C]')
      loc_ch = loc.shape[3]
      if loc.stride(0) == 0 or loc.shape[0] == 1:
        loc_c = _f32c(loc[0], 'local_features')
      else:
        loc_c, loc_bs = _f32c(loc, 'local_features'), n * loc_ch
    DL = D + loc_ch
    seeds_c = _i64c(seeds, 'cluster_indices')
    seed_bs = 0 if seeds.dim() == 2 else n
    fbuf = torch.empty(cap * (D + DL + 2), dtype=torch.float32, device=dev)
    lbuf = torch.empty(cap * 5, dtype=torch.int64, device=dev)
    head = (B + 2 + 4 + 3) // 4 * 4      # img_off [B + 1], num_segments [1], read-back scratch [4]
    ibuf = torch.empty(head + cap * 3, dtype=torch.int32, device=dev)
    ws = _workspace(_cached_size('spml_segment_by_kmeans_workspace_bytes', B, n, DL, int(num_k),
                                 int(iterations)), dev)
    a = _lib.ClusterArgs()
    a.emb, a.loc, a.loc_batch_stride = emb.data_ptr(), (loc_c.data_ptr() if loc_ch else None), loc_bs
    if labels is not None:
      labels_c = _i64c(labels, 'segment_by_kmeans(labels)')
      a.labels = labels_c.data_ptr()
      a.has_ignore = int(ignore_index is not None)
      if torch.is_tensor(ignore_index):      # stays on the device: no host sync
        ignore_dev = ignore_index.to(device=dev, dtype=torch.int64).reshape(1)
        a.ignore_index_dev = ignore_dev.data_ptr()
      elif ignore_index is not None:
        a.ignore_index = int(ignore_index)
    else:
      sem_c, inst_c = _i64c(sem, 'semantic_labels'), _i64c(inst, 'instance_labels')
      a.sem, a.inst = sem_c.data_ptr(), inst_c.data_ptr()
      a.semantic_ignore = int(semantic_ignore)
    a.label_divisor = int(divisor) if divisor else 0
    a.seeds, a.seed_batch_stride = seeds_c.data_ptr(), seed_bs
    a.k_per_image = k_per_image.data_ptr() if k_per_image is not None else None
    a.batch_index_offset = int(batch_index_offset)
    a.batch, a.dim, a.n, a.loc_ch = B, D, n, loc_ch
    a.num_clusters, a.iterations, a.eps = int(num_k), int(iterations), EPS
    fp, lp, ip = fbuf.data_ptr(), lbuf.data_ptr(), ibuf.data_ptr()
    a.e, a.el = fp, fp + 4 * cap * D
    a.nx, a.nc = fp + 4 * cap * (D + DL), fp + 4 * cap * (D + DL + 1)
    a.labels_out, a.batch_out, a.segment_ids = lp, lp + 8 * cap, lp + 16 * cap
    if divisor:
      a.sem_out, a.inst_out = lp + 24 * cap, lp + 32 * cap
    a.img_off, a.num_segments = ip, ip + 4 * (B + 1)
    a.dst = ip + 4 * head
    a.kmeans_labels, a.seed_out = ip + 4 * (head + cap), ip + 4 * (head + 2 * cap)
    # the one host synchronisation of the step happens inside the call: rows kept, segments
    # and the pending error bits come back in one 16-byte copy
    counts = (ctypes.c_int32 * 4)()
    a.counts_host = ctypes.addressof(counts)
    a.counts_dev = ip + 4 * (B + 2)
    a.status = status_word(dev).data_ptr()
    call('spml_segment_by_kmeans', ctypes.byref(a), ptr(ws), ws.numel(), stream_of(emb))
    rows, segments, bits = counts[0], counts[1], counts[2]
    if bits:
      raise _status_error(bits)
    e = torch.as_strided(fbuf, (rows, D), (D, 1), 0)
    el = torch.as_strided(fbuf, (rows, DL), (DL, 1), cap * D)
    lab = torch.as_strided(lbuf, (rows,), (1,), 0)
    bid = torch.as_strided(lbuf, (rows,), (1,), cap)
    cid = torch.as_strided(lbuf, (rows,), (1,), 2 * cap)
    outs = [e, el, lab, cid, bid]
    if divisor:
      outs += [torch.as_strided(lbuf, (rows,), (1,), 3 * cap),
               torch.as_strided(lbuf, (rows,), (1,), 4 * cap)]
    # the backward reads e, el, the two norms and the pixel -> row map: all inside these two
    # buffers (kept whole; the addresses are re-derived from the dims)
    ctx.save_for_backward(fbuf, ibuf)
    ctx.dims = (B, D, loc_ch, H, W, head)
    ctx.mark_non_differentiable(*outs[2:])
    ctx.set_materialize_grads(False)
    box.append((rows, segments, ibuf))
    return tuple(outs)

  @staticmethod
  def backward(ctx, de, del_, *_):
    if de is None and del_ is None:
      return (None,) * 14
    fbuf, ibuf = ctx.saved_tensors
    B, D, loc_ch, H, W, head = ctx.dims
    if _C is not None:
      return (_C.segment_bwd(fbuf, ibuf, de, del_, B, D, loc_ch, H, W, head),) + (None,) * 13
    cap, DL = B * H * W, D + loc_ch
    demb = torch.empty(B, D, H, W, dtype=torch.float32, device=fbuf.device)
    de = _f32c(de, 'd(cluster_embedding)') if de is not None else None
    del_ = _f32c(del_, 'd(cluster_embedding_with_loc)') if del_ is not None else None
    fp = fbuf.data_ptr()
    call('spml_normalize_pack_bwd', ptr(de), ptr(del_), fp, fp + 4 * cap * D,
         fp + 4 * cap * (D + DL), fp + 4 * cap * (D + DL + 1), ibuf.data_ptr() + 4 * head, B, D,
         loc_ch, H * W, EPS, demb.data_ptr(), stream_of(fbuf))
    return (demb,) + (None,) * 13


class GatherPrototypesFn(torch.autograd.Function):
  """spml/models/utils.py:100-116 for ids fresh from segment_by_kmeans: per-segment labels and
  the two prototype sets in one library call."""

  @staticmethod
  def forward(ctx, e, el, cid, bid, sem, inst, m):
    if _C is not None:
      protos, protos_loc, p_sem, p_inst, p_bid, fbuf, cid_c = _C.gather_fwd(
          e, el, cid, bid, sem, inst, int(m), status_word(e.device))
      ctx.save_for_backward(fbuf, cid_c)
      ctx.dims = (int(m), e.shape[1], el.shape[1])
      ctx.mark_non_differentiable(p_sem, p_inst, p_bid)
      ctx.set_materialize_grads(False)
      return protos, protos_loc, p_sem, p_inst, p_bid
    e = _f32c(e, 'gather(embeddings)')
    el = _f32c(el, 'gather(embeddings_with_loc)')
    rows, D, DL = e.shape[0], e.shape[1], el.shape[1]
    dev = e.device
    m = int(m)
    fbuf = torch.empty(m * (D + DL + 2), dtype=torch.float32, device=dev)
    plab = torch.empty(3 * m, dtype=torch.int64, device=dev)
    ws = _workspace(_cached_size('spml_gather_prototypes_workspace_bytes', m, D, DL), dev)
    protos = torch.as_strided(fbuf, (m, D), (D, 1), 0)
    protos_loc = torch.as_strided(fbuf, (m, DL), (DL, 1), m * D)
    cid, bid = _i64c(cid, 'cluster_indices'), _i64c(bid, 'batch_indices')
    sem, inst = _i64c(sem, 'semantic_labels'), _i64c(inst, 'instance_labels')
    fp, lp = fbuf.data_ptr(), plab.data_ptr()
    call('spml_gather_prototypes_fwd', e.data_ptr(), el.data_ptr(), rows, D, DL, cid.data_ptr(),
         bid.data_ptr(), sem.data_ptr(), inst.data_ptr(), m, EPS, fp, fp + 4 * m * D,
         fp + 4 * m * (D + DL), fp + 4 * m * (D + DL + 1), lp, lp + 8 * m, lp + 16 * m,
         status_word(dev).data_ptr(), ws.data_ptr(), ws.numel(), stream_of(e))
    ctx.save_for_backward(fbuf, cid)
    ctx.dims = (m, D, DL)
    p_sem = torch.as_strided(plab, (m,), (1,), 0)
    p_inst = torch.as_strided(plab, (m,), (1,), m)
    p_bid = torch.as_strided(plab, (m,), (1,), 2 * m)
    ctx.mark_non_differentiable(p_sem, p_inst, p_bid)
    ctx.set_materialize_grads(False)       # an unused prototype set costs nothing in the backward
    return protos, protos_loc, p_sem, p_inst, p_bid

  @staticmethod
  def backward(ctx, dp, dpl, *_):
    if dp is None and dpl is None:
      return None, None, None, None, None, None, None
    fbuf, cid = ctx.saved_tensors
    m, D, DL = ctx.dims
    if _C is not None:
      de, del_ = _C.gather_bwd(fbuf, cid, dp, dpl, m, D, DL)
      return de, del_, None, None, None, None, None
    rows = cid.shape[0]
    dev = fbuf.device
    de = del_ = None
    if dp is not None:
      dp = _f32c(dp, 'd(prototypes)')
      de = torch.empty(rows, D, dtype=torch.float32, device=dev)
    if dpl is not None:
      dpl = _f32c(dpl, 'd(prototypes_with_loc)')
      del_ = torch.empty(rows, DL, dtype=torch.float32, device=dev)
    fp = fbuf.data_ptr()
    call('spml_gather_prototypes_bwd', ptr(dp), ptr(dpl), fp, fp + 4 * m * D,
         fp + 4 * m * (D + DL), fp + 4 * m * (D + DL + 1), cid.data_ptr(), rows, D, DL, m, EPS,
         ptr(de), ptr(del_), stream_of(fbuf))
    return de, del_, None, None, None, None, None


ENABLE_ANN, ENABLE_OCC, ENABLE_SIM, ENABLE_ACC = 1, 2, 4, 8


class HeadSpec:
  """The non-differentiable inputs of one spml_head_fwd / spml_head_bwd pair: labels, tags,
  memory bank, switches.  Keeps every tensor whose address is in the argument struct alive."""

  def __init__(self, cid, bid, sem, inst, psem, pinst, pbid, num_classes, enable, kappas, weights,
               max_groups, max_rows_per_group=0, img_tags=None, ptags=None, tag_cols=(0, 0),
               bank=(), nn_tags=False, img_sim_on_plain=False, protos_loc=None,
               nn_threshold=0.95):
    self.call = None
    if _C is not None:
      occ = bool(enable & ENABLE_OCC)
      self.call = _C.HeadCall(
          cid, bid, sem, inst, psem, pinst, pbid, int(num_classes), int(enable),
          [float(k or 0.0) for k in kappas], [float(w or 0.0) for w in weights],
          int(max_groups), int(max_rows_per_group), img_tags, ptags, int(tag_cols[0]),
          int(tag_cols[1]), [b['prototype'] for b in bank], [b['semantic_label'] for b in bank],
          [b['batch_index'] for b in bank],
          [b['semantic_tag'] for b in bank] if (occ and not nn_tags) else [],
          [b['prototype_with_loc'] for b in bank] if (occ and nn_tags) else [],
          bool(nn_tags), bool(img_sim_on_plain), protos_loc, float(nn_threshold))
      self.enable, self.img_sim_on_plain = int(enable), bool(img_sim_on_plain)
      return
    self.keep = []

    def i64(t, name):
      if t is None:
        return None
      t = _i64c(t, name)
      self.keep.append(t)
      return t

    def f32(t, name):
      if t is None:
        return None
      t = _f32c(t.detach(), name)
      self.keep.append(t)
      return t

    def tags2d(t, name):
      t = i64(t, name)
      if t.dim() != 2 or t.stride(1) != 1:
        t = t.reshape(-1, t.shape[-1]).contiguous()
        self.keep.append(t)
      return t

    a = _lib.HeadArgs()
    self.args = a
    self.cid = i64(cid, 'cluster_index').view(-1)
    a.seg = self.cid.data_ptr()
    a.n = self.cid.shape[0]
    for field, t, name in (('bid', bid, 'cluster_batch_index'),
                           ('sem', sem, 'cluster_semantic_label'),
                           ('inst', inst, 'cluster_instance_label'),
                           ('psem', psem, 'prototype_semantic_label'),
                           ('pinst', pinst, 'prototype_instance_label'),
                           ('pbid', pbid, 'prototype_batch_index')):
      t = i64(t, name)
      setattr(a, field, t.data_ptr() if t is not None else None)
    a.m = psem.shape[0]
    a.num_classes = int(num_classes)
    a.enable = int(enable)
    a.kappa_ann, a.kappa_occ, a.kappa_sim = [float(k or 0.0) for k in kappas]
    a.weight_ann, a.weight_occ, a.weight_sim = [float(w or 0.0) for w in weights]
    a.max_groups = max(1, int(max_groups))
    a.max_rows_per_group = int(max_rows_per_group)
    a.nn_tags, a.img_sim_on_plain = int(nn_tags), int(img_sim_on_plain)
    a.nn_threshold, a.eps = float(nn_threshold), EPS
    if enable & ENABLE_OCC:
      if nn_tags:
        t = f32(protos_loc, 'prototype_with_loc')
        a.protos_loc, a.dim_loc = t.data_ptr(), t.shape[1]
        a.wide_tags = int(num_classes > 32)
      else:
        it, pt = tags2d(img_tags, 'semantic_tag'), tags2d(ptags, 'prototype_semantic_tag')
        a.img_tags, a.img_tags_ld, a.tag_rows = it.data_ptr(), it.stride(0), it.shape[0]
        a.ptags, a.ptags_ld = pt.data_ptr(), pt.stride(0)
        a.tag_col0, a.tag_col1 = int(tag_cols[0]), int(tag_cols[1])
        if tag_cols[1] - tag_cols[0] > 64 or tag_cols[1] > it.shape[1]:
          raise ValueError('image tags: at most 64 tag columns inside the tag matrix')
        a.wide_tags = int(tag_cols[1] - tag_cols[0] > 32)
    if len(bank) > _lib.MAX_BANK:
      raise ValueError('memory bank: at most %d entries per call' % _lib.MAX_BANK)
    a.num_bank = len(bank)
    for i, entry in enumerate(bank):
      bp = f32(entry['prototype'], 'memory_prototype')
      bs = i64(entry['semantic_label'], 'memory_prototype_semantic_label')
      a.bank_protos[i], a.bank_psem[i], a.bank_m[i] = bp.data_ptr(), bs.data_ptr(), bp.shape[0]
      if enable & ENABLE_OCC:
        if nn_tags:
          bl = f32(entry['prototype_with_loc'], 'memory_prototype_with_loc')
          bb = i64(entry['batch_index'], 'memory_prototype_batch_index')
          a.bank_protos_loc[i], a.bank_pbid[i] = bl.data_ptr(), bb.data_ptr()
        else:
          bt = tags2d(entry['semantic_tag'], 'memory_prototype_semantic_tag')
          a.bank_tags[i], a.bank_tags_ld[i] = bt.data_ptr(), bt.stride(0)


class HeadLossFn(torch.autograd.Function):
  """Everything of Segsort*.losses() but the conv classifier, forward and backward, as one
  library call each.  Returns five 0-dim tensors: weighted sem_ann, sem_occ, img_sim, the
  top-5 accuracy (entries of disabled losses are 0) and the sum of the enabled losses
  (what train.py:213-219 adds up: `sum(losses)`, same order, same fp32 additions)."""

  @staticmethod
  def forward(ctx, e, el, protos, spec):
    if spec.call is not None:
      out = spec.call.forward(e, el, protos, status_word(e.device))
      ctx.spec = spec
      ctx.set_materialize_grads(False)
      sem_ann, sem_occ, img_sim, acc, total = out.unbind(0)
      ctx.mark_non_differentiable(acc)
      return sem_ann, sem_occ, img_sim, acc, total
    e = _f32c(e, 'cluster_embedding')
    protos = _f32c(protos, 'prototype')
    a = spec.args
    dev = e.device
    a.e, a.dim = e.data_ptr(), e.shape[1]
    if el is not None:
      el = _f32c(el, 'cluster_embedding_with_loc')
      a.el, a.dim_loc = el.data_ptr(), el.shape[1]
    a.protos = protos.data_ptr()
    if e.shape[0] != a.n or protos.shape[0] != a.m:
      raise ValueError('embeddings / prototypes disagree with their label vectors')
    a.status = status_word(dev).data_ptr()
    state = _workspace(_lib.load().spml_head_workspace_bytes(ctypes.byref(a)), dev)
    out = torch.empty(5, dtype=torch.float32, device=dev)
    call('spml_head_fwd', ctypes.byref(a), state.data_ptr(), state.numel(), out.data_ptr(),
         stream_of(e))
    ctx.spec, ctx.state = spec, state
    ctx.save_for_backward(e, el, protos)
    ctx.set_materialize_grads(False)
    sem_ann, sem_occ, img_sim, acc, total = out.unbind(0)
    ctx.mark_non_differentiable(acc)
    return sem_ann, sem_occ, img_sim, acc, total

  @staticmethod
  def backward(ctx, g_ann, g_occ, g_sim, _g_acc, g_total):
    if ctx.spec.call is not None:
      if g_ann is None and g_occ is None and g_sim is None and g_total is None:
        return None, None, None, None
      need_e, need_el, need_p = ctx.needs_input_grad[:3]
      de, del_, dprotos = ctx.spec.call.backward(g_ann, g_occ, g_sim, g_total, bool(need_p))
      return (de if need_e else None), (del_ if need_el else None), dprotos, None
    e, el, protos = ctx.saved_tensors
    spec, state = ctx.spec, ctx.state
    a = spec.args
    gs = [None if g is None else (g if g.dtype == torch.float32 else g.float())
          for g in (g_ann, g_occ, g_sim, g_total)]
    if all(g is None for g in gs):
      return None, None, None, None
    need_e, need_el, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
    sim_on_el = bool(a.enable & ENABLE_SIM) and not a.img_sim_on_plain
    de = torch.empty_like(e)
    del_ = torch.empty_like(el) if (el is not None and sim_on_el) else None
    dprotos = torch.empty_like(protos) if need_p else None
    call('spml_head_bwd', ctypes.byref(a), ptr(state), state.numel(), ptr(gs[0]), ptr(gs[1]),
         ptr(gs[2]), ptr(gs[3]),
         ptr(de), ptr(del_), ptr(dprotos), stream_of(e))
    if el is not None and del_ is None and need_el:
      del_ = torch.zeros_like(el)
    return (de if need_e else None), (del_ if need_el else None), dprotos, None


# ------------------------------------------------------------------------------ entry points
#
# What the operator modules call.  With the ATen binding the autograd nodes are C++
# (binding.cpp: SegmentFn / GatherFn / HeadFn: no torch.autograd.Function.apply, no Python in
# the backward); without it the torch.autograd.Function classes above do the same over ctypes.


def segment_by_kmeans_stage(emb, loc, labels, sem, inst, divisor, semantic_ignore, ignore_index,
                            seeds, k_per_image, num_k, iterations, batch_index_offset):
  """-> (outputs, rows kept, segments, int32 buffer that starts with the image offsets)."""
  if _C is not None:
    ignore_dev = ignore_index if torch.is_tensor(ignore_index) else None
    outs, rows, segments, bits, ibuf = _C.segment(
        emb, loc, labels, sem, inst, int(divisor) if divisor else 0,
        int(semantic_ignore) if semantic_ignore is not None else 0,
        ignore_index is not None, 0 if (ignore_dev is not None or ignore_index is None)
        else int(ignore_index), ignore_dev, seeds, k_per_image, int(num_k), int(iterations),
        int(batch_index_offset), status_word(emb.device))
    if bits:
      raise _status_error(bits)
    return tuple(outs), rows, segments, ibuf
  box = []
  outs = SegmentByKmeansFn.apply(emb, loc, labels, sem, inst, divisor, semantic_ignore,
                                 ignore_index, seeds, k_per_image, num_k, iterations,
                                 batch_index_offset, box)
  rows, segments, ibuf = box[0]
  return outs, rows, segments, ibuf


def gather_prototypes_stage(e, el, cid, bid, sem, inst, m):
  """-> (prototypes, prototypes_with_loc, semantic, instance, batch index of every segment)."""
  if _C is not None:
    return _C.gather(e, el, cid, bid, sem, inst, int(m), status_word(e.device))
  return GatherPrototypesFn.apply(e, el, cid, bid, sem, inst, m)


def head_losses_stage(e, el, protos, spec):
  """-> (sem_ann, sem_occ, img_sim, accuracy, sum of the enabled losses)."""
  if spec.call is not None:
    return _C.head(e, el, protos, spec.call, status_word(e.device))
  return HeadLossFn.apply(e, el, protos, spec)
