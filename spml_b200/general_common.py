"""Drop-in replacements for the hot-path functions of spml/utils/general/common.py."""

from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def normalize_embedding(embeddings, eps=1e-12):
  """spml/utils/general/common.py:101-120: x / max(||x||_2, eps) over the last dim."""
  return ops.NormalizeRows.apply(embeddings, float(eps))


def resize_labels(labels, size):
  """spml/utils/general/common.py:11-26 (nearest-neighbour label resize; this is the
  caller side of the path and stays a library call)."""
  n, h, w = labels.shape
  labels = F.interpolate(labels.view(n, 1, h, w).float(), size=size, mode='nearest')
  return labels.squeeze_(1).long()


def one_hot(labels, max_label=None):
  """spml/utils/general/common.py:76-98."""
  if max_label is None:
    max_label = int(labels.max()) + 1
  shape = labels.shape
  out = torch.zeros((labels.numel(), int(max_label)), dtype=torch.long, device=labels.device)
  out.scatter_(1, labels.reshape(-1, 1), 1)
  return out.view(list(shape) + [int(max_label)])


def segment_mean(x, index):
  """spml/utils/general/common.py:123-147 (tf.segment_mean).  Not on the training path (its
  only caller is pyscripts/inference/pseudo_denseposerw_crf.py:172): plain library scatter-adds
  on the device, same arithmetic as the reference."""
  x = x.view(-1, x.shape[-1])
  index = index.view(-1)
  m = int(index.max()) + 1
  count = torch.zeros(m, dtype=torch.float32, device=x.device)
  count.scatter_add_(0, index, torch.ones_like(index, dtype=torch.float32))
  count = torch.where(count == 0, torch.ones_like(count), count)
  total = torch.zeros((m, x.shape[-1]), dtype=torch.float32, device=x.device)
  total.scatter_add_(0, index.view(-1, 1).expand(-1, x.shape[-1]), x)
  return total.div_(count.view(-1, 1))
