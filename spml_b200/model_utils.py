"""Drop-in replacements for the hot-path functions of spml/models/utils.py."""

from __future__ import annotations

import torch

from . import ops
from . import segsort_common


def _cat_on(tensors, device):
  tensors = [t if t.device == device else t.to(device) for t in tensors]
  return tensors[0] if len(tensors) == 1 else torch.cat(tensors, 0)


def gather_clustering_and_update_prototypes(embeddings, embeddings_with_loc, cluster_indices,
                                            batch_indices, semantic_labels, instance_labels,
                                            anchor_device=None):
  """spml/models/utils.py:41-131.

  Takes and returns per-device LISTS like the reference.  With one process per
  GPU (the B200 deployment) every list has one entry and nothing is copied; with
  several entries they are concatenated on `anchor_device`, as the reference's
  gather does, and the results are handed back to each entry's device.

  Returns (prototypes, prototypes_with_loc, prototype_semantic_labels,
  prototype_instance_labels, prototype_batch_indices, cluster_indices), each a
  list over devices.
  """
  sections = [c.shape[0] for c in cluster_indices]
  devices = [c.device for c in cluster_indices]
  anchor = torch.device(anchor_device) if anchor_device is not None else devices[0]
  if anchor.type != 'cuda':
    raise RuntimeError('gather_clustering_and_update_prototypes: spml_b200 needs CUDA '
                       'tensors (no CPU path)')
  if anchor.index is None:
    anchor = torch.device('cuda', torch.cuda.current_device())
  meta = segsort_common.segment_meta(cluster_indices[0]) if len(cluster_indices) == 1 else None

  e = _cat_on(embeddings, anchor)
  el = _cat_on(embeddings_with_loc, anchor)
  cid = _cat_on(cluster_indices, anchor)
  bid = _cat_on(batch_indices, anchor)
  sem = _cat_on(semantic_labels, anchor)
  inst = _cat_on(instance_labels, anchor)

  if meta is not None and meta.num_rows == cid.shape[0] and cid.device == anchor:
    # ids straight from segment_by_kmeans are already the dense ranks of
    # (image, cluster, label): both re-numberings of :95-108 are the identity, and the
    # segment count is known -> no unique, no host sync, one library call.  The kernels
    # check on the device that every segment carries one (batch, sem, inst) triple, i.e. that
    # these really are the labels the ids were made from (ops.check_status reports it).
    new_cid = cid
    protos, protos_loc, p_sem, p_inst, p_bid = ops.gather_prototypes_stage(
        e, el, cid, bid, sem, inst, meta.num_segments)
  else:
    if cid.numel() == 0:
      raise RuntimeError('gather_clustering_and_update_prototypes: no pixels')
    inv1, _, _, _, _ = ops.unique_inverse(cid, hi=bid, bound=0, want_keys=False)   # :95-97
    div = torch.maximum(inst.max(), sem.max()) + 1                                   # :100
    lab = bid * div * div + sem * div + inst
    new_cid, _, plab, count, _ = ops.unique_inverse(lab, hi=inv1, bound=0)          # :106-108
    m = int(count)                                                                   # host sync
    plab = plab[:m]
    p_bid = plab // (div * div)
    p_sem = (plab % (div * div)) // div
    p_inst = plab % div
    protos = ops.SegmentPrototypes.apply(e, new_cid, m)                              # :113-116
    protos_loc = ops.SegmentPrototypes.apply(el, new_cid, m)

  if len(sections) == 1:
    split_cid = [new_cid]
  else:
    split_cid = [c.to(d) for c, d in zip(torch.split(new_cid, sections), devices)]
  return ([protos.to(d) for d in devices], [protos_loc.to(d) for d in devices],
          [p_sem.to(d) for d in devices], [p_inst.to(d) for d in devices],
          [p_bid.to(d) for d in devices], split_cid)


def gather_and_update_datas(datas, anchor_device=None):
  """spml/models/utils.py:134-154: concatenate a per-device list and hand the result to
  every device."""
  devices = [d.device for d in datas]
  anchor = torch.device(anchor_device) if anchor_device is not None else devices[0]
  gathered = _cat_on(list(datas), anchor) if len(datas) > 1 else datas[0]
  return [gathered.to(d) for d in devices]


def gather_multiset_labels_per_batch_by_nearest_neighbor(
    embeddings, prototypes, semantic_prototype_labels, batch_embedding_labels,
    batch_prototype_labels, num_classes=21, top_k=3, threshold=0.95, label_divisor=255):
  """spml/models/utils.py:157-223: multi-hot [rows, num_classes] tags from each row's top_k
  most similar prototypes of the same image that carry a class label (< num_classes) and are
  at least `threshold` similar.  One top-k launch over the [rows, prototypes] similarities
  (never materialised) instead of mm + where + topk + gathers; `label_divisor` is unused in
  the reference too."""
  return ops.nn_multiset_labels(embeddings.detach(), prototypes.detach(),
                                semantic_prototype_labels, batch_embedding_labels,
                                batch_prototype_labels, num_classes, top_k, threshold)


def get_params(model, prefixs, suffixes, exclude=None):
  """spml/models/utils.py:12-38: the trainable parameters of the sub-modules named in
  `prefixs` whose leaf name starts with (or whose full name ends with) one of `suffixes`;
  `exclude` is a list of full names or a substring."""
  modules = dict(model.named_modules())
  for prefix in prefixs:
    module = modules.get(prefix)
    if module is None:
      continue
    for leaf, param in module.named_parameters():
      full = prefix + '.' + leaf
      if isinstance(exclude, list) and full in exclude:
        continue
      if isinstance(exclude, str) and exclude in full:
        continue
      if not param.requires_grad:
        continue
      last = full.rsplit('.', 1)[-1]
      for suffix in suffixes:      # one yield per matching suffix, like the reference's loop
        if last.startswith(suffix) or full.endswith(suffix):
          yield param
