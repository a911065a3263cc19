"""Drop-in replacements for spml/utils/segsort/loss.py: SegSortLoss / SetSegSortLoss
with the same constructor and forward signature, computed by the fused
similarity + mask + row-sum kernels of libspml_b200.so (forward and backward)."""

from __future__ import annotations

import torch
from torch.nn.modules.loss import _Loss

from . import _lib
from . import ops

_REDUCTIONS = {'mean': _lib.REDUCE_MEAN, 'sum': _lib.REDUCE_SUM}


def _check_mode(group_mode):
  if group_mode != 'segsort+':
    # every construction site of the reference passes 'segsort+'
    # (segsort.py:56-63, segsort_softmax.py:76-83)
    raise NotImplementedError("spml_b200 implements group_mode='segsort+' only")


def _loss(emb, pix_code, seg, protos, proto_code, kappa, mode, reduction, path='auto'):
  if reduction not in _REDUCTIONS:
    raise NotImplementedError("spml_b200 implements reduction 'mean' and 'sum'")
  problem = ops.SegsortProblem(pix_code, seg, proto_code, kappa, mode,
                               reduction=_REDUCTIONS[reduction], path=path)
  return ops.SegsortLossFn.apply(emb, protos, problem)


class SegSortLoss(_Loss):
  """spml/utils/segsort/loss.py:133-190."""

  def __init__(self, concentration=10, group_mode='segsort+', size_average=None,
               reduce=None, reduction='mean'):
    super(SegSortLoss, self).__init__(size_average, reduce, reduction)
    _check_mode(group_mode)
    self.concentration = concentration
    self.group_mode = group_mode

  def __repr__(self):
    return 'SegSortLoss(concentration={:.2f}, group_mode={})'.format(
        self.concentration, self.group_mode)

  def forward(self, embeddings, semantic_labels, instance_labels, prototypes,
              prototype_semantic_labels, prototype_weights=None):
    return _loss(embeddings, semantic_labels, instance_labels, prototypes,
                 prototype_semantic_labels, self.concentration, _lib.MODE_CLASS,
                 self.reduction)


class SetSegSortLoss(_Loss):
  """spml/utils/segsort/loss.py:193-251: semantic labels are multi-hot [*, C <= 64]
  matrices; two pixels/prototypes are 'same' when their tag sets intersect."""

  def __init__(self, concentration=10, group_mode='segsort+', size_average=None,
               reduce=None, reduction='mean'):
    super(SetSegSortLoss, self).__init__(size_average, reduce, reduction)
    _check_mode(group_mode)
    self.concentration = concentration
    self.group_mode = group_mode

  def __repr__(self):
    return 'SetSegSortLoss(concentration={:.2f}, group_mode={})'.format(
        self.concentration, self.group_mode)

  def forward(self, embeddings, semantic_labels, instance_labels, prototypes,
              prototype_semantic_labels, prototype_weights=None):
    pix = ops.pack_tags(semantic_labels.view(-1, semantic_labels.shape[-1]))
    pro = ops.pack_tags(prototype_semantic_labels.view(-1, prototype_semantic_labels.shape[-1]))
    # the tcgen05 epilogue compares 32-bit codes: wider tag sets take the fp32 kernels
    path = 'fp32' if semantic_labels.shape[-1] > 32 else 'auto'
    return _loss(embeddings, pix, instance_labels, prototypes, pro, self.concentration,
                 _lib.MODE_TAGS, self.reduction, path=path)
