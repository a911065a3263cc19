"""Embedding-affinity random walk for pseudo labels (SURVEY.md 8f-4;
pyscripts/inference/pseudo_softmaxrw_crf.py:135-170).

Adjacent to the hot path, not on it: an [n, n] pixel affinity exp(5 cos - 5), raised to the
20th power, column-normalised into a transition matrix, squared `walk_steps` times and applied
to the class activation maps.  These are plain dense GEMMs (n = H/8 * W/8 ~ 4 096), so they go
to cuBLAS through torch.matmul in fp32 (no TF32), exactly the reference's arithmetic.
"""

from __future__ import annotations

import torch


def embedding_affinity(embeddings):
  """pseudo_softmaxrw_crf.py:136-140: embeddings [1, C, h, w] -> exp(5 (e_i . e_j) - 5) [n, n]
  on channel-normalised embeddings."""
  embs = embeddings / torch.norm(embeddings, dim=1)
  flat = embs.view(embs.shape[1], -1)
  return torch.matmul(flat.t(), flat).mul_(5).add_(-5).exp_()


def random_walk(affinities, cam, walk_steps=6, power=20):
  """pseudo_softmaxrw_crf.py:158-170: `affinities` is a list of [n, n] matrices (one per flip /
  scale, averaged), `cam` the [classes, h, w] activation maps; returns the propagated maps."""
  prev = torch.backends.cuda.matmul.allow_tf32
  torch.backends.cuda.matmul.allow_tf32 = False
  try:
    aff = torch.mean(torch.stack(list(affinities), dim=0), dim=0)
    aff_mat = aff ** power
    trans = aff_mat / torch.sum(aff_mat, dim=0, keepdim=True)
    for _ in range(walk_steps):
      trans = torch.matmul(trans, trans)
    out = torch.matmul(cam.reshape(cam.shape[0], -1), trans)
    return out.view(cam.shape)
  finally:
    torch.backends.cuda.matmul.allow_tf32 = prev
