"""Rank plumbing for the data-parallel deployment: one process per GPU, the minibatch
sharded over images, no data-path collective.

The reference runs single-process `DataParallel` with a gather of every pixel embedding
to an anchor GPU (spml/models/utils.py:86-127).  Here each rank owns its images end to
end; the only collective of a training step is the backbone's gradient all-reduce
(`all_reduce_gradients`, what DDP does), and timings are reduced with a max over ranks.
Works with the `nccl` backend on GPUs and `gloo` on CPUs (used by the tests).
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
  """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) without it."""
  return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
          int(os.environ.get('LOCAL_RANK', '0')))


def init(backend=None):
  """Initialises the default process group when launched under torchrun."""
  rank, size, local_rank = world()
  if size > 1 and not dist.is_initialized():
    if backend is None:
      backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    kwargs = {}
    if backend == 'nccl':
      torch.cuda.set_device(local_rank)
      kwargs['device_id'] = torch.device('cuda', local_rank)
    dist.init_process_group(backend, **kwargs)
  return rank, size, local_rank


def shard_images(num_images, rank, world_size):
  """Contiguous, balanced [begin, end) range of a global batch owned by `rank`
  (the reference gives every GPU its own DataLoader batch, others.py:62-71)."""
  base, extra = divmod(num_images, world_size)
  begin = rank * base + min(rank, extra)
  return begin, begin + base + (1 if rank < extra else 0)


def batch_index_offset(images_per_rank, rank):
  """Global index of a rank's first image: what `N * gpu_id` is in common.py:376-377."""
  return images_per_rank * rank


def max_over_ranks(value, device=None):
  """Max of a python float over all ranks (device-side timing is the max over ranks)."""
  if not dist.is_initialized() or dist.get_world_size() == 1:
    return float(value)
  t = torch.tensor([float(value)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t[0])


def sum_over_ranks(value, device=None):
  if not dist.is_initialized() or dist.get_world_size() == 1:
    return float(value)
  t = torch.tensor([float(value)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return float(t[0])


def all_reduce_gradients(parameters, bucket_bytes=64 << 20):
  """One bucketed all-reduce (mean) of the gradients of `parameters`: the single
  collective of a step.  NVSwitch makes the cost latency- not link-bound, so the
  buckets are large."""
  if not dist.is_initialized() or dist.get_world_size() == 1:
    return 0
  size = dist.get_world_size()
  grads = [p.grad for p in parameters if p.grad is not None]
  buckets, cur, cur_bytes = [], [], 0
  for g in grads:
    nbytes = g.numel() * g.element_size()
    if cur and (cur_bytes + nbytes > bucket_bytes or g.dtype != cur[0].dtype):
      buckets.append(cur)
      cur, cur_bytes = [], 0
    cur.append(g)
    cur_bytes += nbytes
  if cur:
    buckets.append(cur)
  for bucket in buckets:
    flat = torch.cat([g.reshape(-1) for g in bucket])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(size)
    off = 0
    for g in bucket:
      g.copy_(flat[off:off + g.numel()].view_as(g))
      off += g.numel()
  return len(buckets)


# ------------------------------------------------------------------------------ SURVEY.md 8f-1
#
# Cross-GPU prototype exchange (the reference contrasts every pixel with the prototypes of
# ALL GPUs and lets the gradient flow back, spml/models/utils.py:86-127).  Segments never span
# images, so a rank's prototypes are complete on that rank: the exchange is an all-gather of
# the ranks' [M_r, D] blocks (plus their labels), and its backward a reduce-scatter of
# d(prototypes).  `exchange_prototypes` does it for the data-dependent M_r of the drop-in API;
# ContrastiveHead(exchange_prototypes=True) uses it (default off: north_star's PR1 keeps
# rank-local prototypes).


class _AllGatherRows(torch.autograd.Function):
  """y = concat over ranks of x (equal shapes).  Backward: every rank holds a gradient for
  the whole y; rank r receives the sum over ranks of block r (reduce-scatter)."""

  @staticmethod
  def forward(ctx, x, group):
    size = dist.get_world_size(group)
    x = x.contiguous()
    out = x.new_empty((size * x.shape[0],) + tuple(x.shape[1:]))
    try:
      dist.all_gather_into_tensor(out, x, group=group)
    except (RuntimeError, NotImplementedError):      # backends without the tensor variant
      dist.all_gather(list(out.chunk(size, 0)), x, group=group)
    ctx.group, ctx.rows = group, x.shape[0]
    return out

  @staticmethod
  def backward(ctx, grad):
    grad = grad.contiguous()
    rank = dist.get_rank(ctx.group)
    out = grad.new_empty((ctx.rows,) + tuple(grad.shape[1:]))
    try:
      dist.reduce_scatter_tensor(out, grad, op=dist.ReduceOp.SUM, group=ctx.group)
    except (RuntimeError, NotImplementedError):      # gloo: all-reduce, keep the own block
      total = grad.clone()
      dist.all_reduce(total, op=dist.ReduceOp.SUM, group=ctx.group)
      out = total[rank * ctx.rows:(rank + 1) * ctx.rows].clone()
    return out, None


def all_gather_prototypes(prototypes, *labels, group=None):
  """Gathers the fixed-capacity prototype block of every rank: returns the differentiable
  [world * M_cap, D] prototypes followed by the gathered label tensors (no gradient).
  Rank r's block sits at rows [r * M_cap, (r + 1) * M_cap); a pixel's own segment id becomes
  `r * M_cap + local_id` (`global_segment_ids`).  Single process: the inputs unchanged."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return (prototypes,) + tuple(labels)
  gathered = [_AllGatherRows.apply(prototypes, group)]
  for lab in labels:
    with torch.no_grad():
      gathered.append(_AllGatherRows.apply(lab.detach(), group))
  return tuple(gathered)


def all_gather_rows(tensor, group=None):
  """Non-differentiable all-gather of equally shaped row blocks (gather_and_update_datas,
  spml/models/utils.py:134-154, for one process per GPU)."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return tensor
  with torch.no_grad():
    return _AllGatherRows.apply(tensor.detach(), group)


def global_segment_ids(local_ids, m_cap, group=None):
  """Column of a pixel's own prototype in the gathered bank (negative ids stay negative)."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return local_ids
  return torch.where(local_ids >= 0, local_ids + dist.get_rank(group) * int(m_cap), local_ids)


def exchange_prototypes(prototypes, prototypes_with_loc, semantic_labels, instance_labels,
                        batch_indices, cluster_indices, group=None):
  """What spml/models/utils.py:86-127 gives every GPU, for one process per GPU: the prototypes
  of ALL ranks (rank-major, i.e. ordered by global image index like the reference's sorted
  ids), their labels, and this rank's pixel -> prototype ids moved into the global numbering.

  One all-gather of the [M_r] counts (a host read-back: the result has a data-dependent
  shape), one all-gather of the padded [M_max, D + D'] prototype rows and one of the padded
  int64 label rows.  The prototypes stay differentiable: the backward reduce-scatters the
  bank gradients, so a rank's pixels receive the gradient of every rank's loss, as in the
  reference.  `batch_indices` must already be global (batch_index_offset = B * rank).
  Returns (prototypes, prototypes_with_loc, semantic_labels, instance_labels, batch_indices,
  cluster_indices).  A single process gets its inputs back."""
  if not dist.is_initialized() or dist.get_world_size(group) == 1:
    return (prototypes, prototypes_with_loc, semantic_labels, instance_labels, batch_indices,
            cluster_indices)
  size, rank = dist.get_world_size(group), dist.get_rank(group)
  dev = prototypes.device
  m, d, dl = prototypes.shape[0], prototypes.shape[1], prototypes_with_loc.shape[1]
  counts = torch.zeros(size, dtype=torch.int64, device=dev)
  dist.all_gather_into_tensor(counts, torch.tensor([m], dtype=torch.int64, device=dev),
                              group=group)
  counts = counts.tolist()                                   # the exchange's host read-back
  m_max = max(counts)
  rows = torch.cat([prototypes, prototypes_with_loc], dim=1)
  labels = torch.stack([semantic_labels, instance_labels, batch_indices], dim=1)
  if m < m_max:
    rows = torch.cat([rows, rows.new_zeros(m_max - m, d + dl)], dim=0)
    labels = torch.cat([labels, labels.new_zeros(m_max - m, 3)], dim=0)
  bank = _AllGatherRows.apply(rows, group)                   # [size * m_max, d + dl]
  with torch.no_grad():
    bank_labels = _AllGatherRows.apply(labels, group)
  if any(c != m_max for c in counts):
    keep = torch.cat([torch.arange(c) + r * m_max for r, c in enumerate(counts)]).to(dev)
    bank = bank.index_select(0, keep)
    bank_labels = bank_labels.index_select(0, keep)
  offset = sum(counts[:rank])
  return (bank[:, :d], bank[:, d:], bank_labels[:, 0].contiguous(),
          bank_labels[:, 1].contiguous(), bank_labels[:, 2].contiguous(),
          cluster_indices + offset)
