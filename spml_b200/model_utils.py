"""Drop-in replacements for the hot-path functions of spml/models/utils.py."""

from __future__ import annotations

import torch

from . import ops


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
  meta = getattr(cluster_indices[0], '_spml_meta', None) if len(cluster_indices) == 1 else None

  e = _cat_on(embeddings, anchor)
  el = _cat_on(embeddings_with_loc, anchor)
  cid = _cat_on(cluster_indices, anchor)
  bid = _cat_on(batch_indices, anchor)
  sem = _cat_on(semantic_labels, anchor)
  inst = _cat_on(instance_labels, anchor)

  if meta is not None and meta.num_rows == cid.shape[0]:
    # ids straight from segment_by_kmeans are already the dense ranks of
    # (image, cluster, label): both re-numberings of :95-108 are the identity,
    # and the segment count is known -> no unique, no host sync.
    m = meta.num_segments
    new_cid = cid
    p_bid = torch.empty(m, dtype=torch.int64, device=anchor).scatter_(0, cid, bid)
    p_sem = torch.empty(m, dtype=torch.int64, device=anchor).scatter_(0, cid, sem)
    p_inst = torch.empty(m, dtype=torch.int64, device=anchor).scatter_(0, cid, inst)
    new_cid._spml_meta = meta
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

  protos = ops.SegmentPrototypes.apply(e, new_cid, m)                                # :113-116
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


def get_params(model, prefixs, suffixes, exclude=None):
  """spml/models/utils.py:12-38 (host glue, unchanged semantics)."""
  for name, module in model.named_modules():
    for prefix in prefixs:
      if name == prefix:
        for n, p in module.named_parameters():
          n = '.'.join([name, n])
          if type(exclude) == list and n in exclude:
            continue
          if type(exclude) == str and exclude in n:
            continue
          for suffix in suffixes:
            if ((n.split('.')[-1].startswith(suffix) or n.endswith(suffix))
                and p.requires_grad):
              yield p
        break
