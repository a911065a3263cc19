"""CPU ORACLE for the SPML pixel-to-segment contrastive path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference's hot path
(twke18/SPML @ 0800afd) so that the CUDA path can be checked against it.  It is
NOT part of the product: only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The
product (`spml_b200`) never imports it and has no CPU fallback.

Where the arithmetic lives.  The reference is pure Python on top of PyTorch
(`requirements.txt:1` "pytorch >= 1.6"; nothing vendored); its numerics are the
ATen CPU kernels `mm`, `scatter_add_`, `argmax`, `unique`, `argsort`, `exp`,
`log`, `norm`.  The oracle issues the same ATen operations in the same order on
the same dtypes, so on the CPU it is bit-identical to the reference.

Pinning.  The reference has no tests, golden vectors or fixtures for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the reference
itself: `tests/golden/make_golden.py` imports /root/reference in the build
container, runs it on seeded inputs and commits inputs + outputs under
`tests/golden/`; `tests/test_oracle_golden.py` requires the oracle to reproduce
them bit-exactly (integers) / to 0 ulp (floats, same ATen kernels).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  `dtype=torch.float64` runs the same algorithm in double
precision; tests use it to separate "differs from the reference's fp32
rounding" from "differs from the mathematics".
"""

from __future__ import annotations

import torch

# --------------------------------------------------------------------------- A1


def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
  """spml/utils/general/common.py:101-120 `normalize_embedding`."""
  n = torch.norm(x, dim=-1, keepdim=True)
  n = torch.where(torch.ge(n, eps), n, torch.ones_like(n).mul_(eps))
  return x / n


# --------------------------------------------------------------------------- A2


def location_grid(hw, device='cpu', kind='int'):
  """spml/utils/segsort/common.py:156-189 `generate_location_features`."""
  if kind == 'int':
    ys, xs = torch.arange(hw[0], device=device), torch.arange(hw[1], device=device)
  elif kind == 'float':
    ys = torch.linspace(0, 1, hw[0], device=device)
    xs = torch.linspace(0, 1, hw[1], device=device)
  else:
    raise ValueError('Type of location features should be either int or float.')
  gy, gx = torch.meshgrid(ys, xs, indexing='ij')
  return torch.stack([gy, gx], dim=2)


# --------------------------------------------------------------------------- A3


def grid_seed_labels(num_clusters, hw, device='cpu'):
  """spml/utils/segsort/common.py:129-153 `initialize_cluster_labels`
  (column-major numbering y + ny * x, round-half-even)."""
  ry = torch.linspace(0, num_clusters[0] - 1, hw[0], device=device).round_().long()
  rx = torch.linspace(0, num_clusters[1] - 1, hw[1], device=device).round_().long()
  return ry.view(-1, 1) + (ry.max() + 1) * rx.view(1, -1)


# --------------------------------------------------------------------------- A4


def prototypes_from_labels(emb, labels, max_label=None):
  """spml/utils/segsort/common.py:11-41 `calculate_prototypes_from_labels`."""
  emb = emb.view(-1, emb.shape[-1])
  if max_label is None:
    max_label = labels.max() + 1
  acc = torch.zeros((max_label, emb.shape[-1]), dtype=emb.dtype, device=emb.device)
  acc = acc.scatter_add_(0, labels.view(-1, 1).expand(-1, emb.shape[-1]), emb)
  return l2_normalize(acc)


# --------------------------------------------------------------------------- A5


def nearest_prototype(emb, protos):
  """spml/utils/segsort/common.py:44-64 `find_nearest_prototypes`."""
  emb = emb.view(-1, protos.shape[-1])
  return torch.argmax(torch.mm(emb, protos.t()), 1)


# --------------------------------------------------------------------------- A6


def spherical_kmeans(emb, init_labels, max_label=None, iterations=10):
  """spml/utils/segsort/common.py:67-97 `kmeans_with_initial_labels`."""
  if max_label is None:
    max_label = init_labels.max() + 1
  lab = init_labels
  for _ in range(iterations):
    lab = nearest_prototype(emb, prototypes_from_labels(emb, lab, max_label))
  return lab


# --------------------------------------------------------------------------- A7


def prototype_labels(sem, inst, offset=256):
  """spml/utils/segsort/common.py:192-218 `prepare_prototype_labels`."""
  keys, inverse = torch.unique(sem + inst * offset, return_inverse=True)
  return keys % offset, inverse


# --------------------------------------------------------------------------- A8


def segment_by_kmeans(emb_nchw, labels=None, num_clusters=(5, 5), cluster_indices=None,
                      local_features=None, ignore_index=None, iterations=10,
                      device_index=0):
  """spml/utils/segsort/common.py:270-408.  `device_index` stands for
  `tensor.device.index` (None on the CPU, which makes :376-377 raise there)."""
  e = l2_normalize(emb_nchw.permute(0, 2, 3, 1).contiguous())          # :306-310
  B, H, W, C = e.shape
  if local_features is None:                                            # :313-317
    local_features = location_grid((H, W), e.device, 'float') - 0.5
    local_features = local_features.view(1, H, W, 2).expand(B, H, W, 2)
  if cluster_indices is None:                                           # :320-323
    cluster_indices = grid_seed_labels(num_clusters, (H, W), e.device)
    cluster_indices = cluster_indices.view(1, H, W).expand(B, H, W)
  if labels is None:                                                    # :326-329
    labels = torch.zeros((B, H, W), dtype=torch.long, device=e.device)
  out = {k: [] for k in ('lab', 'cid', 'bid', 'e', 'el')}
  for b in range(B):                                                    # :337-388
    lab = labels[b].reshape(-1)
    _, cid = torch.unique(cluster_indices[b].reshape(-1), return_inverse=True)
    k = cid.max() + 1
    eb = e[b].view(-1, C)
    el = l2_normalize(torch.cat(
        [eb, local_features[b].reshape(-1, local_features.shape[-1])], -1))
    if ignore_index is not None:
      keep = torch.ne(lab, ignore_index).nonzero().view(-1)
      lab, cid = lab.index_select(0, keep), cid.index_select(0, keep)
      eb, el = eb.index_select(0, keep), el.index_select(0, keep)
    if eb.shape[0] > 0:
      cid = spherical_kmeans(el, cid, k, iterations)
    bid = torch.zeros_like(cid).fill_(b + B * device_index)            # :376-381
    for key, val in zip(('lab', 'cid', 'bid', 'e', 'el'), (lab, cid, bid, eb, el)):
      out[key].append(val)
  lab, cid, bid = (torch.cat(out[k], 0) for k in ('lab', 'cid', 'bid'))
  eb, el = torch.cat(out['e'], 0), torch.cat(out['el'], 0)
  _, cid = torch.unique(bid * (cid.max() + 1) + cid, return_inverse=True)  # :398-401
  _, cid = prototype_labels(lab, cid, lab.max() + 1)                       # :404-405
  return eb, el, lab, cid, bid


# --------------------------------------------------------------------------- A9


def generate_clusters(emb_nchw, sem, inst, local_features, label_divisor,
                      semantic_ignore_index, num_clusters, iterations):
  """spml/models/embeddings/resnet_deeplab.py:90-148 (labels already resized)."""
  lab = sem * label_divisor + inst
  ignore = lab.max() + 1
  lab = lab.masked_fill(sem == semantic_ignore_index, ignore)
  eb, el, lab, cid, bid = segment_by_kmeans(
      emb_nchw, lab, num_clusters, local_features=local_features,
      ignore_index=ignore, iterations=iterations)
  return {'cluster_embedding': eb, 'cluster_embedding_with_loc': el,
          'cluster_semantic_label': lab // label_divisor,
          'cluster_instance_label': lab % label_divisor,
          'cluster_index': cid, 'cluster_batch_index': bid}


# --------------------------------------------------------------------------- A9 (DensePose)


def generate_clusters_densepose(emb_nchw, sem, inst, local_features, label_divisor,
                                semantic_ignore_index, num_clusters, iterations):
  """spml/models/embeddings/resnet_pspnet_densepose.py:90-166: as generate_clusters,
  then `cluster_embedding_with_loc` is REPLACED by normalize(cat[0.1 * e, loc of the kept
  pixels]) (:141-154)."""
  out = generate_clusters(emb_nchw, sem, inst, local_features, label_divisor,
                          semantic_ignore_index, num_clusters, iterations)
  if local_features is not None:
    keep = (sem != semantic_ignore_index).view(-1).nonzero().view(-1)
    loc = local_features.reshape(-1, local_features.shape[-1]).index_select(0, keep)
    out['cluster_embedding_with_loc'] = l2_normalize(
        torch.cat([out['cluster_embedding'] * 0.1, loc], dim=-1))
  return out


# --------------------------------------------------------------------------- general


def one_hot(labels, max_label=None):
  """spml/utils/general/common.py:76-98."""
  if max_label is None:
    max_label = labels.max() + 1
  flat = labels.reshape(-1, 1)
  out = torch.zeros((flat.shape[0], int(max_label)), dtype=torch.long, device=labels.device)
  out = out.scatter_(1, flat, 1)
  return out.view(list(labels.shape) + [int(max_label)])


def segment_mean(x, index):
  """spml/utils/general/common.py:123-147."""
  x, index = x.view(-1, x.shape[-1]), index.view(-1)
  m = index.max() + 1
  total = torch.zeros((m, x.shape[-1]), dtype=torch.float, device=x.device)
  count = torch.zeros((m,), dtype=torch.float, device=x.device)
  count = count.scatter_add_(0, index, torch.ones_like(index, dtype=torch.float))
  count = torch.where(torch.eq(count, 0), torch.ones_like(count), count)
  total = total.scatter_add_(0, index.view(-1, 1).expand(-1, x.shape[-1]), x)
  return total.div_(count.view(-1, 1))


# --------------------------------------------------------------------------- f3


def nn_multiset_labels(emb, protos, psem, batch_emb, batch_proto, num_classes=21, top_k=3,
                       threshold=0.95):
  """spml/models/utils.py:157-223 `gather_multiset_labels_per_batch_by_nearest_neighbor`:
  for every row the top_k most similar prototypes of the SAME image with a class label
  (< num_classes); those at least `threshold` similar contribute their class to a
  multi-hot [rows, num_classes] tag matrix."""
  emb = emb.view(-1, emb.shape[-1])
  protos = protos.view(-1, emb.shape[-1])
  n = emb.shape[0]
  aff = torch.eq(batch_emb.view(-1, 1), batch_proto.view(1, -1))
  aff = aff & (psem < num_classes).view(1, -1)
  dists = torch.mm(emb, protos.t())
  dists = torch.where(aff, dists, dists.min() - 1)
  nn_d, nn_i = torch.topk(dists, top_k, dim=1)
  got = torch.gather(psem.view(1, -1).expand(n, -1), 1, nn_i)
  got = got.masked_fill(nn_d < threshold, num_classes)
  tags = torch.sum(one_hot(got, num_classes + 1), dim=1)
  return (tags > 0).long()[:, :num_classes]


# --------------------------------------------------------------------------- B1


def gather_and_update_prototypes(embs, embs_loc, cids, bids, sems, insts):
  """spml/models/utils.py:41-131 with the cross-device gather (:86-92) replaced
  by concatenation (what `scatter_gather.gather` does for >= 1-D tensors)."""
  sections = [c.shape[0] for c in cids]
  e, el = torch.cat(list(embs), 0), torch.cat(list(embs_loc), 0)
  cid, bid = torch.cat(list(cids), 0), torch.cat(list(bids), 0)
  sem, inst = torch.cat(list(sems), 0), torch.cat(list(insts), 0)
  _, cid = torch.unique(bid * (cid.max() + 1) + cid, return_inverse=True)   # :95-97
  div = max([inst.max() + 1, sem.max() + 1])                                # :100
  lab = bid * div ** 2 + sem * div + inst
  plab, cid = prototype_labels(lab, cid, lab.max() + 1)                     # :106-108
  p_bid, p_sem, p_inst = plab // div ** 2, (plab % div ** 2) // div, plab % div
  protos = prototypes_from_labels(e, cid)                                   # :113-116
  protos_loc = prototypes_from_labels(el, cid)
  n = len(sections)
  return ([protos] * n, [protos_loc] * n, [p_sem] * n, [p_inst] * n, [p_bid] * n,
          list(torch.split(cid, sections)))


# --------------------------------------------------------------------------- C1 / C2


def _nll_from_similarity(sim, self_sim, same, diff):
  """Shared tail of spml/utils/segsort/loss.py:58-82 and :104-130
  (group_mode 'segsort+')."""
  same_sum = torch.sum(sim * same, 1, keepdim=True)
  same_sum -= self_sim
  num = torch.where(torch.gt(same_sum, 0), same_sum, self_sim)
  den = torch.sum(sim * diff, 1, keepdim=True).add_(num)
  return (num / den).log_().mul_(-1)


def segsort_nll(emb, sem, seg, protos, psem, concentration):
  """spml/utils/segsort/loss.py:15-82 `_calculate_log_likelihood`."""
  emb, protos = emb.view(-1, emb.shape[-1]), protos.view(-1, protos.shape[-1])
  sim = torch.mm(emb, protos.t()).mul_(concentration).exp_()
  self_sim = torch.gather(sim, 1, seg.view(-1, 1))
  same = torch.eq(sem.view(-1, 1), psem.view(1, -1)).to(sim.dtype)
  diff = torch.ne(sem.view(-1, 1), psem.view(1, -1)).to(sim.dtype)
  return _nll_from_similarity(sim, self_sim, same, diff)


def set_segsort_nll(emb, tags, seg, protos, ptags, concentration):
  """spml/utils/segsort/loss.py:85-130 `_one_hot_calculate_log_likelihood`."""
  emb, protos = emb.view(-1, emb.shape[-1]), protos.view(-1, protos.shape[-1])
  sim = torch.mm(emb, protos.t()).mul_(concentration).exp_()
  self_sim = torch.gather(sim, 1, seg.view(-1, 1))
  aff = torch.mm(tags.to(sim.dtype), ptags.t().to(sim.dtype))
  return _nll_from_similarity(sim, self_sim, (aff > 0).to(sim.dtype),
                              (aff == 0).to(sim.dtype))


def segsort_loss(emb, sem, seg, protos, psem, concentration, reduction='mean'):
  """spml/utils/segsort/loss.py:149-190 `SegSortLoss.forward`."""
  nll = segsort_nll(emb, sem, seg, protos, psem, concentration)
  return nll.mean() if reduction == 'mean' else nll.sum() if reduction == 'sum' else nll


def set_segsort_loss(emb, tags, seg, protos, ptags, concentration, reduction='mean'):
  """spml/utils/segsort/loss.py:209-251 `SetSegSortLoss.forward`."""
  nll = set_segsort_nll(emb, tags, seg, protos, ptags, concentration)
  return nll.mean() if reduction == 'mean' else nll.sum() if reduction == 'sum' else nll


# --------------------------------------------------------------------------- C3


def top_k_ranking(emb, labels, protos, plabels, top_k=3):
  """spml/utils/segsort/eval.py:9-52."""
  emb, protos = emb.view(-1, emb.shape[-1]), protos.view(-1, protos.shape[-1])
  order = torch.argsort(torch.mm(emb, protos.t()), 1, descending=True)
  order = order[:, :top_k].contiguous()
  hit = torch.gather(torch.eq(labels.view(-1, 1), plabels.view(1, -1)), 1, order)
  got = torch.gather(plabels.view(-1), 0, order.view(-1)).view(-1, top_k)
  return torch.mean(hit.float()), got


# --------------------------------------------------------------------------- C4


def segsort_losses(cfg, datas, targets, sem_ann_base=None):
  """spml/models/predictions/segsort.py:127-243 `Segsort.losses` (the
  parameter-free predictor).  `sem_ann_base` is the classifier cross-entropy that
  segsort_softmax.py:111-131 puts into sem_ann_loss before the SegSort term is added.
  Returns (sem_ann, sem_occ, img_sim, accuracy)."""
  C = cfg.dataset.num_classes
  cid, emb = datas['cluster_index'], datas['cluster_embedding']
  sem, bid = datas['cluster_semantic_label'], datas['cluster_batch_index']
  protos, psem = targets['prototype'], targets['prototype_semantic_label']
  pbid = targets['prototype_batch_index']
  tags = torch.index_select(targets['semantic_tag'][:, 1:C], 0, bid)        # :146-150
  ptags = targets['prototype_semantic_tag'][:, 1:C]
  mem_p = targets.get('memory_prototype', [])
  mem_s = targets.get('memory_prototype_semantic_label', [])
  mem_b = targets.get('memory_prototype_batch_index', [])
  mem_t = targets.get('memory_prototype_semantic_tag', [])
  if mem_p and mem_s and mem_t and mem_b:                                   # :153-182
    protos = torch.cat([protos] + list(mem_p), 0)
    psem = torch.cat([psem] + list(mem_s), 0)
    ptags = torch.cat([ptags] + [t[:, 1:C] for t in mem_t], 0)
    pbid = torch.cat([pbid] + list(mem_b), 0)
  pix = (sem < C).nonzero().view(-1)                                        # :184-194
  pro = (psem < C).nonzero().view(-1)
  remap = torch.arange(protos.shape[0], dtype=torch.long, device=protos.device)
  remap = remap.masked_fill(psem >= C, remap.max() + 1)
  _, remap = torch.unique(remap, return_inverse=True)
  new_cid = torch.gather(remap, 0, cid)
  sem_ann = segsort_loss(emb.index_select(0, pix), sem.index_select(0, pix),
                         new_cid.index_select(0, pix), protos.index_select(0, pro),
                         psem.index_select(0, pro), cfg.train.sem_ann_concentration)
  if sem_ann_base is not None:        # segsort_softmax.py:196-202 `sem_ann_loss += ...`
    sem_ann = sem_ann_base + sem_ann
  sem_ann = sem_ann * cfg.train.sem_ann_loss_weight
  sem_occ = set_segsort_loss(emb, tags, cid, protos, ptags,
                             cfg.train.sem_occ_concentration)
  sem_occ = sem_occ * cfg.train.sem_occ_loss_weight
  acc, _ = top_k_ranking(protos, psem, protos, psem, 5)                     # :212-217
  emb_loc, inst = datas['cluster_embedding_with_loc'], datas['cluster_instance_label']
  per_image = []
  for b in torch.unique(bid):                                               # :226-240
    idx = (bid == b).nonzero().view(-1)
    e_b, l_b = emb_loc.index_select(0, idx), inst.index_select(0, idx)
    pl_b, c_b = prototype_labels(l_b, cid.index_select(0, idx), l_b.max() + 1)
    p_b = prototypes_from_labels(e_b, c_b)
    per_image.append(segsort_loss(e_b, l_b, c_b, p_b, pl_b,
                                  cfg.train.img_sim_concentration))
  img_sim = sum(per_image) / len(per_image) * cfg.train.img_sim_loss_weight
  return sem_ann, sem_occ, img_sim, acc


# --------------------------------------------------------------------------- C4 (softmax)


def make_classifier(cfg, seed=235, dtype=torch.float32):
  """The `semantic_classifier` stack of segsort_softmax.py:23-38 /
  segsort_softmax_densepose.py:23-38 with seeded weights (state-dict compatible)."""
  import torch.nn as nn
  dim = cfg.network.embedding_dim
  gen = torch.Generator().manual_seed(seed)
  net = nn.Sequential(
      nn.Conv2d(dim, dim * 2, kernel_size=3, padding=1, stride=1, bias=False),
      nn.BatchNorm2d(dim * 2), nn.ReLU(inplace=True), nn.Dropout(p=0.75),
      nn.Conv2d(dim * 2, cfg.dataset.num_classes, kernel_size=1, stride=1, bias=True))
  with torch.no_grad():
    for p in net.parameters():
      p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
    net[1].weight.add_(1.0)
    net[1].running_mean.copy_(torch.randn(dim * 2, generator=gen) * 0.05)
    net[1].running_var.copy_(torch.rand(dim * 2, generator=gen) * 0.5 + 0.75)
  return net.to(dtype)


def classifier_cross_entropy(cfg, classifier, embedding, semantic_label):
  """segsort_softmax.py:111-131: cross-entropy of the conv classifier on DETACHED,
  channel-normalised embeddings, logits resized (bilinear) to the label map, labels
  >= num_classes ignored."""
  import torch.nn.functional as F
  e = embedding.detach()
  e = e / torch.norm(e, dim=1, keepdim=True)
  logits = classifier(e)
  logits = F.interpolate(logits, size=semantic_label.shape[-2:], mode='bilinear')
  lab = semantic_label.masked_fill(semantic_label >= cfg.dataset.num_classes,
                                   cfg.dataset.semantic_ignore_index)
  lab = lab.squeeze_(1).long()
  return F.cross_entropy(logits, lab, ignore_index=cfg.dataset.semantic_ignore_index)


def segsort_softmax_losses(cfg, classifier, datas, targets):
  """spml/models/predictions/segsort_softmax.py:103-242 `SegsortSoftmax.losses`: the
  losses of `Segsort.losses` with the classifier's cross-entropy added to sem_ann_loss
  BEFORE the weighting (`sem_ann_loss += ...; sem_ann_loss *= weight`, :196-202)."""
  ce = classifier_cross_entropy(cfg, classifier, datas['embedding'], targets['semantic_label'])
  return segsort_losses(cfg, datas, targets, sem_ann_base=ce)


def segsort_softmax_densepose_losses(cfg, classifier, datas, targets):
  """spml/models/predictions/segsort_softmax_densepose.py:104-254.  Differences from the
  VOC head: prototype tag sets come from a same-image 1-NN over `prototype_with_loc`
  (threshold 0.95, :174-191) instead of image tags; an all-zero tag row becomes all-ones
  (:186-189); img_sim runs on `cluster_embedding` (not `_with_loc`, :232-250); the memory
  bank needs no tag list (:158-172)."""
  C = cfg.dataset.num_classes
  t = cfg.train
  sem_ann = classifier_cross_entropy(cfg, classifier, datas['embedding'],
                                     targets['semantic_label'])
  sem_occ = img_sim = acc = None
  use_ann, use_occ = t.sem_ann_loss_types != 'none', t.sem_occ_loss_types == 'segsort'
  cid, emb = datas['cluster_index'], datas['cluster_embedding']
  if use_ann or use_occ:
    sem = datas['cluster_semantic_label']
    protos, protos_loc = targets['prototype'], targets['prototype_with_loc']
    psem, pbid = targets['prototype_semantic_label'], targets['prototype_batch_index']
    mem_p = targets.get('memory_prototype', [])
    mem_l = targets.get('memory_prototype_with_loc', [])
    mem_s = targets.get('memory_prototype_semantic_label', [])
    mem_b = targets.get('memory_prototype_batch_index', [])
    if mem_p and mem_s and mem_b:                                           # :158-172
      protos = torch.cat([protos] + list(mem_p), 0)
      protos_loc = torch.cat([protos_loc] + list(mem_l), 0)
      psem = torch.cat([psem] + list(mem_s), 0)
      pbid = torch.cat([pbid] + list(mem_b), 0)
    ptags = nn_multiset_labels(protos_loc, protos_loc, psem, pbid, pbid, num_classes=C,
                               top_k=1, threshold=0.95)                     # :174-185
    empty = torch.max(ptags, dim=1, keepdim=True)[0] == 0
    ptags = ptags.masked_fill(empty.expand(-1, C), 1)                       # :186-189
    tags = torch.index_select(ptags, 0, cid)
    pix = (sem < C).nonzero().view(-1)                                      # :194-204
    pro = (psem < C).nonzero().view(-1)
    remap = torch.arange(protos.shape[0], dtype=torch.long, device=protos.device)
    remap = remap.masked_fill(psem >= C, remap.max() + 1)
    _, remap = torch.unique(remap, return_inverse=True)
    new_cid = torch.gather(remap, 0, cid)
    if use_ann:
      sem_ann = sem_ann + segsort_loss(
          emb.index_select(0, pix), sem.index_select(0, pix), new_cid.index_select(0, pix),
          protos.index_select(0, pro), psem.index_select(0, pro), t.sem_ann_concentration)
      sem_ann = sem_ann * t.sem_ann_loss_weight
    if use_occ:
      sem_occ = set_segsort_loss(emb, tags, cid, protos, ptags, t.sem_occ_concentration)
      sem_occ = sem_occ * t.sem_occ_loss_weight
    acc, _ = top_k_ranking(protos, psem, protos, psem, 5)                   # :223-228
  if t.img_sim_loss_types != 'none':                                        # :231-252
    inst, bid = datas['cluster_instance_label'], datas['cluster_batch_index']
    per_image = []
    for b in torch.unique(bid):
      idx = (bid == b).nonzero().view(-1)
      e_b, l_b = emb.index_select(0, idx), inst.index_select(0, idx)
      pl_b, c_b = prototype_labels(l_b, cid.index_select(0, idx), l_b.max() + 1)
      p_b = prototypes_from_labels(e_b, c_b)
      per_image.append(segsort_loss(e_b, l_b, c_b, p_b, pl_b, t.img_sim_concentration))
    img_sim = sum(per_image) / len(per_image) * t.img_sim_loss_weight
  return sem_ann, sem_occ, img_sim, acc


# --------------------------------------------------------------------------- f2


def majority_label_from_topk(top_k_labels, num_classes=None):
  """spml/utils/segsort/eval.py:55-70."""
  return torch.argmax(torch.sum(one_hot(top_k_labels, num_classes), dim=1), 1)


def segsort_predictions(datas, targets):
  """spml/models/predictions/segsort.py:68-125 `Segsort.predictions`: segment prototypes
  of the image -> top-20 retrieval against a prototype memory bank (in <= 10 chunks of
  queries) -> majority vote per segment -> per pixel."""
  bank, bank_lab = targets['semantic_memory_prototype'], targets['semantic_memory_prototype_label']
  emb, cid = datas['cluster_embedding'], datas['cluster_index']
  _, cid = torch.unique(cid, return_inverse=True)
  m = cid.max() + 1
  protos = prototypes_from_labels(emb, cid, m)
  zeros = torch.zeros(m, dtype=torch.long)
  n = zeros.shape[0]
  pred = torch.zeros((n,), dtype=torch.long)
  topk = torch.zeros((n, 20), dtype=torch.long)
  groups = min(10, n - 1)
  r = n // groups
  split = [i * r for i in range(groups)] + [n]
  for i in range(groups):
    st, ed = split[i], split[i + 1]
    _, lab = top_k_ranking(protos[st:ed], zeros[st:ed], bank, bank_lab, 20)
    pred[st:ed] = majority_label_from_topk(lab)
    topk[st:ed] = lab
  return torch.gather(pred, 0, cid), torch.index_select(topk, 0, cid)


# --------------------------------------------------------------------------- f4


def random_walk_cam(embedding_list, cam, walk_steps=6, power=20):
  """pyscripts/inference/pseudo_softmaxrw_crf.py:135-170: affinities exp(5 cos - 5) of the
  channel-normalised [1, C, h, w] embeddings (one per flip / scale, averaged), ** power,
  column-normalised, squared walk_steps times, applied to the [classes, h, w] maps."""
  affs = []
  for embs in embedding_list:
    embs = embs / torch.norm(embs, dim=1)
    flat = embs.view(embs.shape[1], -1)
    affs.append(torch.matmul(flat.t(), flat).mul_(5).add_(-5).exp_())
  aff = torch.mean(torch.stack(affs, dim=0), dim=0)
  aff_mat = aff ** power
  trans = aff_mat / torch.sum(aff_mat, dim=0, keepdim=True)
  for _ in range(walk_steps):
    trans = torch.matmul(trans, trans)
  return torch.matmul(cam.view(cam.shape[0], -1), trans).view(cam.shape)


# --------------------------------------------------------------------------- step


def contrastive_step(cfg, batch, memory_bank=None, dtype=torch.float32, variant='segsort',
                     classifier=None):
  """One contrastive-loss step as pyscripts/train/train.py:167-219,273 runs it on
  ONE device: generate_clusters (A9) -> gather/prototypes (B1) -> tags (B2,
  train.py:194-202) -> Segsort.forward (C4) -> backward to d(embedding).

  `batch` holds the tensors of spml_b200.synth.make_batch; `memory_bank` is a
  dict of lists keyed 'memory_prototype', ... (train.py:204-208,276-293).
  Returns a dict with every intermediate the parity tests compare.

  `variant`: 'segsort' (segsort.py, parameter-free), 'softmax' (segsort_softmax.py, what
  train.py instantiates) or 'densepose' (resnet_pspnet_densepose.py generate_clusters +
  segsort_softmax_densepose.py, train_densepose.py:159-205: no image tags).  The softmax
  variants take `classifier` (make_classifier) and `batch['semantic_label_full']`
  (the full-resolution label map of targets['semantic_label']; default: the resized one).
  """
  emb = batch['embedding'].to(dtype).clone().requires_grad_(True)
  loc = batch['local_feature'].to(dtype)
  gen = generate_clusters_densepose if variant == 'densepose' else generate_clusters
  cl = gen(emb, batch['semantic_label'], batch['instance_label'], loc,
           cfg.network.label_divisor, cfg.dataset.semantic_ignore_index,
           cfg.network.kmeans_num_clusters, cfg.network.kmeans_iterations)
  p, pl, psem, pinst, pbid, cids = gather_and_update_prototypes(
      [cl['cluster_embedding']], [cl['cluster_embedding_with_loc']],
      [cl['cluster_index']], [cl['cluster_batch_index']],
      [cl['cluster_semantic_label']], [cl['cluster_instance_label']])
  datas = dict(cl)
  datas['cluster_index'] = cids[0]
  tags = batch['semantic_tag']
  targets = {'prototype': p[0], 'prototype_with_loc': pl[0],
             'prototype_semantic_label': psem[0], 'prototype_instance_label': pinst[0],
             'prototype_batch_index': pbid[0]}
  if variant != 'densepose':                         # train_densepose.py:189-199 (commented out)
    targets.update({'semantic_tag': tags,
                    'prototype_semantic_tag': tags.index_select(0, pbid[0])})
  if memory_bank:
    targets.update({k: [t.to(dtype) if t.is_floating_point() else t for t in v]
                    for k, v in memory_bank.items()})
  if variant == 'segsort':
    sem_ann, sem_occ, img_sim, acc = segsort_losses(cfg, datas, targets)
  else:
    datas['embedding'] = emb
    targets['semantic_label'] = batch.get('semantic_label_full', batch['semantic_label'])
    fn = segsort_softmax_densepose_losses if variant == 'densepose' else segsort_softmax_losses
    sem_ann, sem_occ, img_sim, acc = fn(cfg, classifier, datas, targets)
  total = sum(x for x in (sem_ann, sem_occ, img_sim) if x is not None)
  total.backward()
  datas.pop('embedding', None)
  targets.pop('semantic_label', None)
  zero = torch.zeros((), dtype=dtype)
  sem_occ = zero if sem_occ is None else sem_occ
  img_sim = zero if img_sim is None else img_sim
  out = {k: v.detach() for k, v in datas.items()}
  out.update({k: v.detach() for k, v in targets.items() if torch.is_tensor(v)})
  out.update({'sem_ann_loss': sem_ann.detach(), 'sem_occ_loss': sem_occ.detach(),
              'img_sim_loss': img_sim.detach(), 'accuracy': acc.detach(),
              'loss': total.detach(), 'grad_embedding': emb.grad.detach()})
  if classifier is not None:
    out['grad_classifier'] = {k: v.grad.detach().clone() for k, v in
                              classifier.named_parameters() if v.grad is not None}
    classifier.zero_grad()
  return out


def memory_bank_update(bank, targets, size, batch_stride):
  """pyscripts/train/train.py:276-293: FIFO of detached 'prototype*' entries;
  stored batch indices are shifted by batch_size * num_gpus every step."""
  for k, v in targets.items():
    if 'prototype' in k and 'memory' not in k and torch.is_tensor(v):
      bank.setdefault('memory_' + k, []).append(v.clone().detach())
      if len(bank['memory_' + k]) > size:
        bank['memory_' + k] = bank['memory_' + k][1:]
  for t in bank.get('memory_prototype_batch_index', []):
    t += batch_stride
  return bank
