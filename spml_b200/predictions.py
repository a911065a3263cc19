"""Drop-in replacements for spml/models/predictions/segsort.py, segsort_softmax.py and
segsort_softmax_densepose.py: the operator API boundary of the contrastive head.

`segsort(config)` / `segsort_softmax(config)` / `segsort_softmax_densepose(config)` return
nn.Modules with the reference's `forward(datas, targets=None, with_loss=True,
with_prediction=False)` contract and the same output keys (sem_ann_loss, sem_occ_loss,
img_sim_loss, accuracy).  Everything of `losses()` except the conv classifier of the softmax
variants is ONE library call forward (spml_head_fwd) and one backward (spml_head_bwd):

  sem_ann  SegSort over labelled pixels x labelled prototypes (+ memory bank); the filters of
           segsort.py:184-201 are a device-side row list and a column mask, nothing is copied
           or re-numbered.
  sem_occ  SetSegSort over all pixels x all prototypes with tag bit masks instead of the
           [N, C] x [C, M] float GEMM of loss.py:107-109.  Tags are the image tags (VOC) or,
           in the DensePose head, the class of each prototype's nearest labelled prototype of
           the same image (segsort_softmax_densepose.py:174-191).
  img_sim  the per-image loop of segsort.py:220-240 as one grouped launch: rows are grouped by
           image and each group only sees its own image's prototypes.
  accuracy top-5 retrieval among all prototypes (segsort.py:212-217).

Preconditions shared with the reference's data flow (train.py:167-211), checked on the device
where noted: `cluster_index` are ids in [0, #prototypes) whose pixels all carry one instance
label (checked, reported through ops.check_status); pixels and prototypes are ordered by image
(what segment_by_kmeans / gather_clustering_and_update_prototypes return).
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from . import segsort_eval


def _construct_loss(loss_types, concentration):
  """segsort.py:54-66: returns the concentration of an enabled loss, None for 'none'."""
  if loss_types in ('segsort', 'set_segsort'):
    return float(concentration)
  if loss_types == 'none':
    return None
  raise KeyError('Unsupported loss types: {:s}'.format(loss_types))


def _bank_entries(targets, with_tags, with_loc):
  """The memory-bank lists of train.py:204-208 as per-step entries, or [] when the
  reference would skip the bank (one of the lists it tests is missing / empty)."""
  protos = targets.get('memory_prototype', [])
  sems = targets.get('memory_prototype_semantic_label', [])
  bids = targets.get('memory_prototype_batch_index', [])
  tags = targets.get('memory_prototype_semantic_tag', [])
  locs = targets.get('memory_prototype_with_loc', [])
  if not (protos and sems and bids) or (with_tags and not tags):
    return []
  entries = []
  for i in range(len(protos)):
    entry = {'prototype': protos[i], 'semantic_label': sems[i], 'batch_index': bids[i]}
    if with_tags:
      entry['semantic_tag'] = tags[i]
    if with_loc:
      entry['prototype_with_loc'] = locs[i]
    entries.append(entry)
  return entries


class Segsort(nn.Module):
  """spml/models/predictions/segsort.py:15-283 (non-parametric predictor)."""

  densepose = False

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
    self.last_loss_total = self._contrastive_total = None
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
    """Returns (sem_ann, sem_occ, img_sim, accuracy); sem_ann is already weighted."""
    C = self.num_classes
    use_ann = self.sem_ann_concentration is not None
    use_occ = self.sem_occ_concentration is not None
    use_sim = self.img_sim_concentration is not None
    if not (use_ann or use_occ or use_sim):
      return None, None, None, None
    contrast = use_ann or use_occ
    enable = ((ops.ENABLE_ANN if use_ann else 0) | (ops.ENABLE_OCC if use_occ else 0) |
              (ops.ENABLE_SIM if use_sim else 0) | (ops.ENABLE_ACC if contrast else 0))
    densepose = self.densepose
    cid = datas['cluster_index']
    e = datas['cluster_embedding']
    el = None if densepose else (datas['cluster_embedding_with_loc'] if use_sim else None)
    bank = _bank_entries(targets, with_tags=not densepose, with_loc=densepose) if contrast else []
    img_tags = ptags = protos_loc = None
    if use_occ and not densepose:
      img_tags, ptags = targets['semantic_tag'], targets['prototype_semantic_tag']
    if use_occ and densepose:
      protos_loc = targets['prototype_with_loc']
    # upper bounds the grouped img_sim launch is sized with
    emb_map = datas.get('embedding', None)
    if emb_map is not None:
      groups, rows_per_group = emb_map.shape[0], emb_map.shape[2] * emb_map.shape[3]
    elif targets.get('semantic_tag', None) is not None:
      groups, rows_per_group = targets['semantic_tag'].shape[0], 0
    else:
      bid = datas['cluster_batch_index']
      groups, rows_per_group = int(bid[-1] - bid[0]) + 1, 0            # host sync
    spec = ops.HeadSpec(
        cid, datas['cluster_batch_index'], datas['cluster_semantic_label'],
        datas['cluster_instance_label'], targets['prototype_semantic_label'],
        targets.get('prototype_instance_label', None) if use_sim else None,
        targets['prototype_batch_index'], C, enable,
        (self.sem_ann_concentration, self.sem_occ_concentration, self.img_sim_concentration),
        (self.sem_ann_loss_weight, self.sem_occ_loss_weight, self.img_sim_loss_weight),
        max_groups=groups, max_rows_per_group=rows_per_group, img_tags=img_tags, ptags=ptags,
        tag_cols=(0, C) if densepose else (1, C), bank=bank, nn_tags=densepose,
        img_sim_on_plain=densepose, protos_loc=protos_loc)
    sem_ann, sem_occ, img_sim, acc, total = ops.head_losses_stage(e, el, targets['prototype'],
                                                                  spec)
    # the library's own `sum(losses)` (train.py:213-219) for callers that only need the sum
    # (object.__setattr__: nn.Module.__setattr__ costs ~5 us per assignment)
    object.__setattr__(self, '_contrastive_total', total)
    return (sem_ann if use_ann else None, sem_occ if use_occ else None,
            img_sim if use_sim else None, acc if contrast else None)

  def losses(self, datas, targets={}):
    """segsort.py:127-243."""
    object.__setattr__(self, '_contrastive_total', None)
    out = self._contrastive_losses(datas, targets)
    # `last_loss_total`: sem_ann + sem_occ + img_sim as train.py:213-219 adds them, computed by
    # the same kernel (saves the caller three tiny additions and their autograd nodes)
    object.__setattr__(self, 'last_loss_total', self._contrastive_total)
    object.__setattr__(self, '_contrastive_total', None)
    return out

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
  """spml/models/predictions/segsort_softmax.py:15-290: the same contrastive head plus a
  2-conv softmax classifier on DETACHED, channel-normalised embeddings whose cross-entropy
  is added to sem_ann_loss before the weighting (:111-131,196-202).  The classifier is
  ordinary cuDNN work and stays torch.nn (state-dict compatible with the reference's)."""

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

  def _logits(self, emb):
    emb = emb / torch.norm(emb, dim=1, keepdim=True)
    return self.semantic_classifier(emb)

  def predictions(self, datas, targets={}):
    """segsort_softmax.py:88-101."""
    logits = self._logits(datas['embedding'])
    return torch.argmax(logits, dim=1), logits

  def losses(self, datas, targets={}):
    """segsort_softmax.py:103-242 / segsort_softmax_densepose.py:104-254."""
    logits = self._logits(datas['embedding'].detach())
    labels = targets.get('semantic_label', None)
    logits = F.interpolate(logits, size=labels.shape[-2:], mode='bilinear')
    labels = labels.masked_fill(labels >= self.num_classes, self.semantic_ignore_index)
    ce = self.softmax_loss(logits, labels.squeeze(1).long())
    sem_ann, sem_occ, img_sim, acc = self._contrastive_losses(datas, targets)
    object.__setattr__(self, '_contrastive_total', None)    # the cross-entropy joins sem_ann here
    object.__setattr__(self, 'last_loss_total', None)
    if sem_ann is not None:
      # `sem_ann_loss += segsort; sem_ann_loss *= weight` with the SegSort term already weighted
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
    """segsort_softmax.py:270-290: weights of the classifier at 10x, biases at 20x the base
    learning rate and without weight decay."""
    weights, biases = [], []
    for name, p in self.semantic_classifier.named_parameters():
      if not p.requires_grad:
        continue
      leaf = name.split('.')[-1]
      if leaf.startswith('weight'):
        weights.append(p)
      elif leaf.startswith('bias'):
        biases.append(p)
    return [{'params': weights, 'lr': 10}, {'params': biases, 'lr': 20, 'weight_decay': 0}]


class SegsortSoftmaxDensepose(SegsortSoftmax):
  """spml/models/predictions/segsort_softmax_densepose.py:16-301 (there also named
  `SegsortSoftmax`).  Differences from the VOC head, all inside the one library call:
  prototype tag sets come from a same-image 1-NN over `prototype_with_loc` (threshold 0.95,
  a prototype without a labelled neighbour gets every tag, :174-191), pixels inherit their
  segment's tags, img_sim runs on `cluster_embedding` (:232-250) and the memory bank needs
  no tag list (:158-172)."""

  densepose = True


def segsort(config):
  """spml/models/predictions/segsort.py:281-283."""
  return Segsort(config)


def segsort_softmax(config):
  """spml/models/predictions/segsort_softmax.py:293-295 (there also named `segsort`)."""
  return SegsortSoftmax(config)


def segsort_softmax_densepose(config):
  """spml/models/predictions/segsort_softmax_densepose.py:298-301 (there named `segsort`)."""
  return SegsortSoftmaxDensepose(config)
