"""Drop-in replacements for spml/models/predictions/segsort.py and segsort_softmax.py:
the operator API boundary of the contrastive head.

`segsort(config)` / `segsort_softmax(config)` return nn.Modules with the reference's
`forward(datas, targets=None, with_loss=True, with_prediction=False)` contract and
the same output keys (sem_ann_loss, sem_occ_loss, img_sim_loss, accuracy).  The
three losses run as one fused launch each (forward) instead of the reference's
index_select copies and per-image Python loop:

  sem_ann  SegSort over labelled pixels x labelled prototypes (+ memory bank);
           the filters of segsort.py:184-201 become a device-side row list and a
           column mask, nothing is copied or re-numbered.
  sem_occ  SetSegSort over all pixels x all prototypes with 20-bit image-tag masks
           instead of the [N, 20] x [20, M] float GEMM of loss.py:107-109.
  img_sim  the per-image loop of segsort.py:220-240 as ONE grouped launch: rows are
           grouped by image and each group only sees its own image's prototypes.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import ops
from . import model_utils
from . import segsort_eval


def _construct_loss(loss_types, concentration):
  """segsort.py:54-66: returns the concentration of an enabled loss, None for 'none'."""
  if loss_types in ('segsort', 'set_segsort'):
    return float(concentration)
  if loss_types == 'none':
    return None
  raise KeyError('Unsupported loss types: {:s}'.format(loss_types))


def _row_groups(batch_indices, proto_batch_indices, max_groups):
  """Row / column offsets of the per-image groups.  Pixels and prototypes are both
  sorted by image (segment ids are ranks of (image, cluster, label)), so every
  image is one contiguous range of rows and of prototype columns."""
  dev = batch_indices.device
  base = torch.minimum(batch_indices.min(), proto_batch_indices.min())
  rel = (batch_indices - base).clamp_(0, max_groups - 1)
  prel = (proto_batch_indices - base).clamp_(0, max_groups - 1)
  rows = torch.zeros(max_groups, dtype=torch.int32, device=dev)
  rows.scatter_add_(0, rel, torch.ones_like(rel, dtype=torch.int32))
  cols = torch.zeros(max_groups, dtype=torch.int32, device=dev)
  cols.scatter_add_(0, prel, torch.ones_like(prel, dtype=torch.int32))
  zero = torch.zeros(1, dtype=torch.int32, device=dev)
  group_off = torch.cat([zero, torch.cumsum(rows, 0, dtype=torch.int32)])
  col_off = torch.cat([zero, torch.cumsum(cols, 0, dtype=torch.int32)])
  return group_off, col_off


class Segsort(nn.Module):
  """spml/models/predictions/segsort.py:15-283 (non-parametric predictor)."""

  def __init__(self, config):
    super(Segsort, self).__init__()
    t = config.train
    self.sem_ann_concentration = _construct_loss(t.sem_ann_loss_types, t.sem_ann_concentration)
    self.sem_ann_loss_weight = t.sem_ann_loss_weight
    occ = 'set_segsort' if t.sem_occ_loss_types == 'segsort' else 'none'
    self.sem_occ_concentration = _construct_loss(occ, t.sem_occ_concentration)
    self.sem_occ_loss_weight = t.sem_occ_loss_weight
    self.img_sim_concentration = _construct_loss(t.img_sim_loss_types, t.img_sim_concentration)
    self.img_sim_loss_weight = t.img_sim_loss_weight
    # configured and constructed by the reference, never computed or returned
    # (segsort.py:42-47): kept for attribute compatibility.
    self.feat_aff_concentration = _construct_loss(t.feat_aff_loss_types,
                                                  t.feat_aff_concentration)
    self.feat_aff_loss_weight = t.feat_aff_loss_weight
    self.semantic_ignore_index = config.dataset.semantic_ignore_index
    self.num_classes = config.dataset.num_classes
    self.label_divisor = config.network.label_divisor

  # ----------------------------------------------------------------------------
  def predictions(self, datas, targets={}):
    """segsort.py:68-125: nearest-neighbour retrieval against a prototype memory bank
    (top-20, majority vote), one launch instead of 10 chunked argsorts."""
    semantic_pred, semantic_topk = None, None
    bank = targets.get('semantic_memory_prototype', None)
    bank_labels = targets.get('semantic_memory_prototype_label', None)
    emb = datas.get('cluster_embedding', None)
    cid = datas.get('cluster_index', None)
    if bank is not None and bank_labels is not None and emb is not None and cid is not None:
      inverse, _, _, count, _ = ops.unique_inverse(cid, want_keys=False)
      m = int(count)
      protos = ops.SegmentPrototypes.apply(emb.detach(), inverse, m)
      zeros = torch.zeros(m, dtype=torch.long, device=protos.device)
      _, topk = segsort_eval.top_k_ranking(protos, zeros, bank, bank_labels, 20)
      majority = segsort_eval.majority_label_from_topk(topk)
      semantic_pred = torch.gather(majority, 0, inverse)
      semantic_topk = torch.index_select(topk, 0, inverse)
    return semantic_pred, semantic_topk

  # ----------------------------------------------------------------------------
  def _contrastive_losses(self, datas, targets):
    C = self.num_classes
    sem_ann_loss = sem_occ_loss = img_sim_loss = sem_ann_acc = None

    if self.sem_ann_concentration is not None or self.sem_occ_concentration is not None:
      cid = datas['cluster_index']
      emb = datas['cluster_embedding']
      sem = datas['cluster_semantic_label']
      bid = datas['cluster_batch_index']
      protos = targets['prototype']
      psem = targets['prototype_semantic_label']
      tags_img = targets['semantic_tag']
      ptags = targets['prototype_semantic_tag']

      mem_p = targets.get('memory_prototype', [])
      mem_s = targets.get('memory_prototype_semantic_label', [])
      mem_b = targets.get('memory_prototype_batch_index', [])
      mem_t = targets.get('memory_prototype_semantic_tag', [])
      # image-tag bit masks: columns 1..C-1 of the 256-wide presence vectors
      # (segsort.py:146-150); one mask per image, gathered per pixel / prototype
      pix_mask = torch.index_select(ops.pack_tags(tags_img[:, 1:C]), 0, bid)
      proto_mask = ops.pack_tags(ptags[:, 1:C])
      if mem_p and mem_s and mem_t and mem_b:                              # :153-182
        protos = torch.cat([protos] + list(mem_p), dim=0)
        psem = torch.cat([psem] + list(mem_s), dim=0)
        proto_mask = torch.cat([proto_mask] + [ops.pack_tags(t[:, 1:C]) for t in mem_t], dim=0)

      n = emb.shape[0]
      if self.sem_ann_concentration is not None:
        # :184-201 labelled pixels x labelled prototypes, without copies
        keep = (sem < C).to(torch.int64).view(1, n)
        _, rows, off = ops.valid_scan(keep, 0, 1, n, want_src=True)
        problem = ops.SegsortProblem(
            sem, cid, psem, self.sem_ann_concentration, _lib.MODE_CLASS,
            row_index=rows, group_off=off, num_groups=1, n_rows=n, max_rows_per_group=n,
            proto_valid=(psem < C))
        sem_ann_loss = ops.SegsortLossFn.apply(emb, protos, problem) * self.sem_ann_loss_weight
      if self.sem_occ_concentration is not None:
        problem = ops.SegsortProblem(pix_mask, cid, proto_mask, self.sem_occ_concentration,
                                     _lib.MODE_TAGS)
        sem_occ_loss = ops.SegsortLossFn.apply(emb, protos, problem) * self.sem_occ_loss_weight
      sem_ann_acc, _ = segsort_eval.top_k_ranking(protos, psem, protos, psem, 5)   # :212-217

    if self.img_sim_concentration is not None:                              # :220-240
      cid = datas['cluster_index']
      emb_loc = datas['cluster_embedding_with_loc']
      inst = datas['cluster_instance_label']
      bid = datas['cluster_batch_index']
      pbid = targets['prototype_batch_index']
      m = pbid.shape[0]
      n = emb_loc.shape[0]
      # per-image prototypes of the location-augmented embeddings == the rows of one
      # batched segment-prototype launch (each segment lives in exactly one image)
      protos_loc = ops.SegmentPrototypes.apply(emb_loc, cid, m)
      pinst = targets.get('prototype_instance_label', None)
      if pinst is None:
        pinst = torch.zeros(m, dtype=torch.int64, device=cid.device).scatter_(0, cid, inst)
      groups = int(targets['semantic_tag'].shape[0])
      group_off, col_off = _row_groups(bid, pbid, groups)
      problem = ops.SegsortProblem(
          inst, cid, pinst, self.img_sim_concentration, _lib.MODE_CLASS,
          reduction=_lib.REDUCE_GROUP_MEAN, group_off=group_off, col_off=col_off,
          num_groups=groups, n_rows=n, max_rows_per_group=n)
      img_sim_loss = ops.SegsortLossFn.apply(emb_loc, protos_loc, problem) * self.img_sim_loss_weight

    return sem_ann_loss, sem_occ_loss, img_sim_loss, sem_ann_acc

  def losses(self, datas, targets={}):
    """segsort.py:127-243."""
    return self._contrastive_losses(datas, targets)

  def forward(self, datas, targets=None, with_loss=True, with_prediction=False):
    targets = targets if targets is not None else {}
    outputs = {}
    if with_prediction:
      semantic_pred, semantic_score = self.predictions(datas, targets)
      outputs.update({'semantic_prediction': semantic_pred, 'semantic_score': semantic_score})
    if with_loss:
      sem_ann_loss, sem_occ_loss, img_sim_loss, sem_ann_acc = self.losses(datas, targets)
      outputs.update({'sem_ann_loss': sem_ann_loss, 'sem_occ_loss': sem_occ_loss,
                      'img_sim_loss': img_sim_loss, 'accuracy': sem_ann_acc})
    return outputs

  def get_params_lr(self):
    return []


class SegsortSoftmax(Segsort):
  """spml/models/predictions/segsort_softmax.py:15-290: the same contrastive head
  plus a 2-conv softmax classifier on DETACHED, L2-normalised embeddings whose
  cross-entropy is added to sem_ann_loss before weighting (:111-131,196-202).
  The classifier is ordinary cuDNN work and stays torch.nn."""

  def __init__(self, config):
    super(SegsortSoftmax, self).__init__(config)
    dim = config.network.embedding_dim
    self.semantic_classifier = nn.Sequential(
        nn.Conv2d(dim, dim * 2, kernel_size=3, padding=1, stride=1, bias=False),
        nn.BatchNorm2d(dim * 2),
        nn.ReLU(inplace=True),
        nn.Dropout(p=0.75),
        nn.Conv2d(dim * 2, config.dataset.num_classes, kernel_size=1, stride=1, bias=True))
    self.softmax_loss = nn.CrossEntropyLoss(ignore_index=config.dataset.semantic_ignore_index)

  def predictions(self, datas, targets={}):
    """segsort_softmax.py:88-101."""
    emb = datas['embedding']
    emb = emb / torch.norm(emb, dim=1, keepdim=True)
    logits = self.semantic_classifier(emb)
    return torch.argmax(logits, dim=1), logits

  def losses(self, datas, targets={}):
    """segsort_softmax.py:103-242."""
    emb = datas['embedding'].detach()
    emb = emb / torch.norm(emb, dim=1, keepdim=True)
    logits = self.semantic_classifier(emb)
    labels = targets.get('semantic_label', None)
    logits = F.interpolate(logits, size=labels.shape[-2:], mode='bilinear')
    labels = labels.masked_fill(labels >= self.num_classes, self.semantic_ignore_index)
    ce = self.softmax_loss(logits, labels.squeeze_(1).long())
    sem_ann, sem_occ, img_sim, acc = self._contrastive_losses(datas, targets)
    if self.sem_ann_concentration is not None:
      # (ce + segsort) * weight, as `sem_ann_loss += ...; sem_ann_loss *= weight`
      sem_ann = ce * self.sem_ann_loss_weight + sem_ann
    else:
      sem_ann = ce
    return sem_ann, sem_occ, img_sim, acc

  def forward(self, datas, targets=None, with_loss=True, with_prediction=False):
    targets = targets if targets is not None else {}
    outputs = {}
    if with_prediction:
      semantic_pred, semantic_logits = self.predictions(datas, targets)
      outputs.update({'semantic_prediction': semantic_pred, 'semantic_logit': semantic_logits})
    if with_loss:
      sem_ann_loss, sem_occ_loss, img_sim_loss, sem_ann_acc = self.losses(datas, targets)
      outputs.update({'sem_ann_loss': sem_ann_loss, 'sem_occ_loss': sem_occ_loss,
                      'img_sim_loss': img_sim_loss, 'accuracy': sem_ann_acc})
    return outputs

  def get_params_lr(self):
    """segsort_softmax.py:270-290."""
    return [
        {'params': [n for n in model_utils.get_params(self, ['semantic_classifier'], ['weight'])],
         'lr': 10},
        {'params': [n for n in model_utils.get_params(self, ['semantic_classifier'], ['bias'])],
         'lr': 20, 'weight_decay': 0},
    ]


def segsort(config):
  """spml/models/predictions/segsort.py:281-283."""
  return Segsort(config)


def segsort_softmax(config):
  """spml/models/predictions/segsort_softmax.py:293-295 (there also named `segsort`)."""
  return SegsortSoftmax(config)
