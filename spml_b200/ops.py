"""Tensor-level wrappers and autograd Functions over the C ABI (spml_b200/_lib.py).

PyTorch is used here for device memory (the caching allocator owns every buffer
the library reads or writes), the current CUDA stream and autograd bookkeeping.
All arithmetic of the path happens inside libspml_b200.so.
"""

from __future__ import annotations

import ctypes

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
  return t.contiguous()


def _i64c(t, name):
  if not t.is_cuda:
    raise RuntimeError('%s: spml_b200 needs CUDA tensors, got %s (no CPU path)'
                       % (name, t.device))
  if t.dtype != torch.int64:
    t = t.long()
  return t.contiguous()


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
               max_rows_per_group=None, proto_valid=None, path='auto', name=''):
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
    _lib.PROFILE_TAG = ':' + problem.name if problem.name else ''
    call('spml_segsort_fwd', ctypes.byref(d), ptr(stats), None, ptr(loss), ptr(ws), ws.numel(),
         stream_of(emb))
    _lib.PROFILE_TAG = ''
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
    if need_p:
      dprotos = torch.empty_like(protos)
    if need_e or need_p:
      ws = ctx.workspace
      d.reserved |= 4
      _lib.PROFILE_TAG = ':' + problem.name if problem.name else ''
      call('spml_segsort_bwd', ctypes.byref(d), ptr(stats), ptr(grad_loss), 0.0, ptr(demb),
           emb.shape[1], ptr(dprotos), ptr(ws), ws.numel(), stream_of(emb))
      _lib.PROFILE_TAG = ''
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
