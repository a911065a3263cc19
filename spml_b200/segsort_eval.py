"""Drop-in replacements for spml/utils/segsort/eval.py."""

from __future__ import annotations

import torch

from . import ops
from . import general_common as common_utils


def top_k_ranking(embeddings, labels, prototypes, prototype_labels, top_k=3):
  """spml/utils/segsort/eval.py:9-52 without the full [Q, M] argsort.  Returns
  (accuracy scalar, [Q, top_k] retrieved labels); similarity descending, lowest
  prototype index first on exact ties."""
  return ops.topk_ranking(embeddings.detach(), labels, prototypes.detach(),
                          prototype_labels, int(top_k))


def majority_label_from_topk(top_k_labels, num_classes=None):
  """spml/utils/segsort/eval.py:55-70."""
  votes = torch.sum(common_utils.one_hot(top_k_labels, num_classes), dim=1)
  return torch.argmax(votes, 1)
