"""Drop-in replacements for spml/utils/segsort/common.py, backed by libspml_b200.so.

Same function names, positional order, defaults and return values as the
reference module.  Every function needs CUDA tensors; there is no CPU path.
"""

from __future__ import annotations

import weakref

import torch

from . import ops
from . import general_common as common_utils


class SegmentMeta:
  """What segment_by_kmeans knows about a result it returned: the ids are already the dense
  ranks of (image, cluster, label), and how many there are."""

  def __init__(self, num_segments, num_rows, img_off, batch, batch_index_offset):
    self.num_segments = num_segments    # M: distinct (image, cluster, label)
    self.num_rows = num_rows            # N: pixels kept
    self.img_off = img_off              # int32 [batch + 1] first row of each image
    self.batch = batch
    self.batch_index_offset = batch_index_offset


# Explicit handle registry: id(cluster_indices tensor) -> (weak reference, SegmentMeta).  An
# entry only answers for the very tensor object segment_by_kmeans returned (a copy, a slice or
# a `.to()` of it is a different object and simply takes the general path); it disappears
# with the tensor.
_segment_registry = {}


def _register_segments(tensor, meta):
  key = id(tensor)
  _segment_registry[key] = (weakref.ref(tensor, lambda _r, k=key: _segment_registry.pop(k, None)),
                            meta)


def segment_meta(tensor):
  """SegmentMeta of a `cluster_indices` tensor returned by segment_by_kmeans, else None."""
  entry = _segment_registry.get(id(tensor))
  if entry is not None and entry[0]() is tensor:
    return entry[1]
  return None


def calculate_prototypes_from_labels(embeddings, labels, max_label=None):
  """spml/utils/segsort/common.py:11-41."""
  if max_label is None:
    max_label = int(labels.max()) + 1          # host sync, as in the reference
  elif torch.is_tensor(max_label):
    max_label = int(max_label)
  return ops.SegmentPrototypes.apply(embeddings, labels, max_label)


def find_nearest_prototypes(embeddings, prototypes):
  """spml/utils/segsort/common.py:44-64 (first index on ties)."""
  return ops.nearest_prototype(embeddings.detach(), prototypes.detach())


def kmeans_with_initial_labels(embeddings, initial_labels, max_label=None, iterations=10):
  """spml/utils/segsort/common.py:67-97.  The labels carry no gradient (the
  reference records a graph through the M-steps that nothing ever uses)."""
  if max_label is None:
    max_label = int(initial_labels.max()) + 1
  elif torch.is_tensor(max_label):
    max_label = int(max_label)
  x = ops._f32c(embeddings.detach(), 'kmeans_with_initial_labels(embeddings)')
  x = x.view(-1, x.shape[-1])
  n = x.shape[0]
  img_off = torch.tensor([0, n], dtype=torch.int32, device=x.device)
  init = initial_labels.reshape(-1).to(torch.int32)
  _, out64 = ops.kmeans(x, img_off, 1, n, max_label, iterations, init)
  # The segment sums are 2^-32 fixed point and need |x| <= 8 (unit vectors on the training
  # path); the kernels flag anything else (also NaN and labels >= max_label) and return -1
  # ids.  The reference synchronises here as well (`.max()`), so the check costs nothing new.
  if n > 0 and int(out64.min()) < 0:
    raise ValueError('kmeans_with_initial_labels: embeddings must be finite with |x| <= 8 and '
                     'initial labels in [0, max_label)')
  return out64


def kmeans(embeddings, num_clusters, iterations=10):
  """spml/utils/segsort/common.py:100-126 is dead code in the reference: it calls
  initialize_cluster_labels without its required `device` argument (:115-116 vs
  :129-131) and raises TypeError.  Same behaviour here; use segment_by_kmeans."""
  raise TypeError("initialize_cluster_labels() missing 1 required positional argument: "
                  "'device' (the reference's kmeans() has no working callers; use "
                  "segment_by_kmeans or kmeans_with_initial_labels)")


def initialize_cluster_labels(num_clusters, img_dimensions, device):
  """spml/utils/segsort/common.py:129-153: y + ny * x with round-half-even."""
  y_labels = torch.linspace(
      0, num_clusters[0] - 1, img_dimensions[0], device=device).round_().long()
  x_labels = torch.linspace(
      0, num_clusters[1] - 1, img_dimensions[1], device=device).round_().long()
  return y_labels.view(-1, 1) + (y_labels.max() + 1) * x_labels.view(1, -1)


def generate_location_features(img_dimensions, device, feature_type='int'):
  """spml/utils/segsort/common.py:156-189."""
  if feature_type == 'int':
    ys = torch.arange(img_dimensions[0], device=device)
    xs = torch.arange(img_dimensions[1], device=device)
  elif feature_type == 'float':
    ys = torch.linspace(0, 1, img_dimensions[0], device=device)
    xs = torch.linspace(0, 1, img_dimensions[1], device=device)
  else:
    raise ValueError('Type of location features should be either int or float.')
  gy, gx = torch.meshgrid(ys, xs, indexing='ij')
  return torch.stack([gy, gx], dim=2)


def prepare_prototype_labels(semantic_labels, instance_labels, offset=256):
  """spml/utils/segsort/common.py:192-218: unique(sem + inst * offset) -> (key %
  offset for every distinct key, rank of each element's key)."""
  bound = int(offset)                            # a tensor offset is read back here
  inverse, _, uniq_lo, count, _ = ops.unique_inverse(semantic_labels, hi=instance_labels,
                                                     bound=bound)
  m = int(count)                                 # host sync sizes the result
  return uniq_lo[:m], inverse.view(semantic_labels.shape)


def find_majority_label_index(semantic_labels, cluster_labels):
  """spml/utils/segsort/common.py:221-267 (not on the training path)."""
  semantic_labels = semantic_labels.view(-1)
  cluster_labels = cluster_labels.view(-1)
  num_clusters = int(cluster_labels.max()) + 1
  num_classes = int(semantic_labels.max()) + 1
  votes = torch.zeros((num_clusters, num_classes), dtype=torch.long,
                      device=semantic_labels.device)
  votes.index_put_((cluster_labels, semantic_labels),
                   torch.ones_like(semantic_labels), accumulate=True)
  majority = torch.argmax(votes, 1)
  keep = torch.eq(torch.gather(majority, 0, cluster_labels), semantic_labels)
  return keep.nonzero(), majority


_SEED_CACHE = {}
_LOC_CACHE = {}


def _seed_map(num_clusters, height, width, device):
  """Grid seeds of initialize_cluster_labels, built once per shape on the host with
  the reference's own ops (so rounding at .5 matches the CPU run of the reference)
  and kept on the device.  Returns (seeds [H, W], K, every-label-present)."""
  key = (int(num_clusters[0]), int(num_clusters[1]), height, width, str(device))
  if key not in _SEED_CACHE:
    seeds = initialize_cluster_labels(num_clusters, (height, width), 'cpu')
    num_k = int(seeds.max()) + 1
    dense = torch.unique(seeds).numel() == num_k
    _SEED_CACHE[key] = (seeds.to(device), num_k, dense)
  return _SEED_CACHE[key]


def _default_location(height, width, device):
  key = (height, width, str(device))
  if key not in _LOC_CACHE:
    loc = generate_location_features((height, width), 'cpu', 'float') - 0.5
    _LOC_CACHE[key] = loc.view(1, height, width, 2).to(device)
  return _LOC_CACHE[key]


def _compress_seed_maps(cluster_indices, batch, n):
  """common.py:339-344: per-image torch.unique(return_inverse) of the seed map."""
  flat = cluster_indices.reshape(batch, n)
  img = torch.arange(batch, device=flat.device).view(-1, 1).expand(batch, n)
  inverse, uniq_hi, _, count, _ = ops.unique_inverse(flat, hi=img, bound=0)
  total = int(count)                             # host sync (user-supplied maps only)
  per_image = torch.zeros(batch, dtype=torch.int64, device=flat.device)
  per_image.scatter_add_(0, uniq_hi[:total], torch.ones(total, dtype=torch.int64,
                                                        device=flat.device))
  first = torch.cumsum(per_image, 0) - per_image
  dense = inverse.view(batch, n) - first.view(-1, 1)
  return dense, per_image.to(torch.int32), int(per_image.max())


def segment_core(embeddings, labels, num_clusters, cluster_indices, local_features,
                 ignore_index, iterations, batch_index_offset=None, after_pack=None):
  """Everything of segment_by_kmeans up to (not including) the host read-back.  Returns
  fixed-CAPACITY tensors (batch * H * W rows; rows past the live count are padding) and
  device-side counts: (e, el, labels, segment_ids, batch_ids, img_off int32 [B+1],
  num_segments int32 [1]).  `after_pack(batch_ids)` is called between the packing and the
  clustering launch (the fixed-capacity head forks side-stream work there)."""
  if embeddings.dim() != 4:
    raise ValueError('embeddings must be [batch, channels, height, width]')
  if not embeddings.is_cuda:
    raise RuntimeError('segment_by_kmeans: spml_b200 needs CUDA tensors (no CPU path)')
  B, C, H, W = embeddings.shape
  n = H * W
  dev = embeddings.device

  if local_features is None:                                            # :313-317
    local_features = _default_location(H, W, dev)
  k_per_image = None
  if cluster_indices is None:                                           # :320-323
    seeds, num_k, dense = _seed_map(num_clusters, H, W, dev)
    if not dense:
      seeds, k_per_image, num_k = _compress_seed_maps(seeds.view(1, n).expand(B, n), B, n)
      seeds = seeds.view(B, H, W)
  else:
    seeds, k_per_image, num_k = _compress_seed_maps(cluster_indices, B, n)
    seeds = seeds.view(B, H, W)
  if labels is None:                                                    # :326-329
    labels = torch.zeros((B, H, W), dtype=torch.long, device=dev)
  labels_c = ops._i64c(labels, 'segment_by_kmeans(labels)').view(B, n)

  dst, _, img_off = ops.valid_scan(labels_c, ignore_index, B, n)        # :355-365
  if batch_index_offset is None:
    batch_index_offset = B * (dev.index or 0)                           # :376-377
  e, el, lab, bid, seed = ops.NormalizePack.apply(
      embeddings, local_features, labels_c, seeds, dst, batch_index_offset)
  if after_pack is not None:     # work that only needs the batch ids can overlap the clustering
    after_pack(bid)
  _, km = ops.kmeans(el.detach(), img_off, B, n, num_k, iterations, seed, k_per_image)
  # :398-405: rank of (image, cluster, label) among the triples that occur
  inverse, _, _, count, _ = ops.unique_inverse(lab, hi=torch.add(km, bid, alpha=num_k), bound=0,
                                               n_dev=img_off[B:], want_keys=False)
  return e, el, lab, inverse, bid, img_off, count, batch_index_offset


def _prepare_seeds(embeddings, num_clusters, cluster_indices):
  B, _, H, W = embeddings.shape
  n = H * W
  dev = embeddings.device
  k_per_image = None
  if cluster_indices is None:                                           # :320-323
    seeds, num_k, dense = _seed_map(num_clusters, H, W, dev)
    if not dense:
      seeds, k_per_image, num_k = _compress_seed_maps(seeds.view(1, n).expand(B, n), B, n)
      seeds = seeds.view(B, H, W)
  else:
    seeds, k_per_image, num_k = _compress_seed_maps(cluster_indices, B, n)
    seeds = seeds.view(B, H, W)
  return seeds, k_per_image, num_k


def segment_clusters(embeddings, labels=None, num_clusters=[5, 5], cluster_indices=None,
                     local_features=None, ignore_index=None, iterations=10,
                     batch_index_offset=None, semantic_labels=None, instance_labels=None,
                     label_divisor=None, semantic_ignore_index=None):
  """segment_by_kmeans, optionally with the label packing of generate_clusters
  (resnet_deeplab.py:112-117,134-135) done inside the library call: pass `semantic_labels`,
  `instance_labels`, `label_divisor` and `semantic_ignore_index` instead of `labels`.  Returns
  the 5-tuple of segment_by_kmeans, plus (semantic, instance) labels per kept pixel when a
  divisor is given."""
  if embeddings.dim() != 4:
    raise ValueError('embeddings must be [batch, channels, height, width]')
  if not embeddings.is_cuda:
    raise RuntimeError('segment_by_kmeans: spml_b200 needs CUDA tensors (no CPU path)')
  B, _, H, W = embeddings.shape
  dev = embeddings.device
  orig_dtype = embeddings.dtype
  if orig_dtype in (torch.bfloat16, torch.float16):
    # autocast backbones (BASELINE configs[2]): the head computes in fp32 like the reference
    embeddings = embeddings.float()
  if local_features is None:                                            # :313-317
    local_features = _default_location(H, W, dev)
  elif local_features.dtype != torch.float32:
    local_features = local_features.float()
  seeds, k_per_image, num_k = _prepare_seeds(embeddings, num_clusters, cluster_indices)
  packed = semantic_labels is not None
  if not packed and labels is None:                                     # :326-329
    labels = torch.zeros((B, H, W), dtype=torch.long, device=dev)
  if batch_index_offset is None:
    batch_index_offset = B * (dev.index or 0)                           # :376-377
  out, rows, segments, img_off = ops.segment_by_kmeans_stage(
      embeddings, local_features, None if packed else labels, semantic_labels, instance_labels,
      label_divisor, semantic_ignore_index, ignore_index, seeds, k_per_image, num_k, iterations,
      batch_index_offset)
  _register_segments(out[3], SegmentMeta(segments, rows, img_off, B, batch_index_offset))
  return out


def segment_by_kmeans(embeddings, labels=None, num_clusters=[5, 5], cluster_indices=None,
                      local_features=None, ignore_index=None, iterations=10,
                      batch_index_offset=None):
  """spml/utils/segsort/common.py:270-408 as ONE library call (~10 kernels) and ONE host
  synchronisation (to size the returned tensors).

  Returns (embeddings [N, C], embeddings_with_loc [N, C+L], labels [N],
  cluster_indices [N], batch_indices [N]) exactly as the reference does.
  `batch_index_offset` (extra keyword, default = the reference's
  `batch * device.index`) is the index given to the first image.
  """
  return segment_clusters(embeddings, labels, num_clusters, cluster_indices, local_features,
                          ignore_index, iterations, batch_index_offset)[:5]
