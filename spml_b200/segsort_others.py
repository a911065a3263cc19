"""Drop-in for spml/utils/segsort/others.py: the on-disk prototype memory bank of the
nearest-neighbour retrieval inference.

Wire format (pyscripts/inference/prototype.py:207-211): one `.npy` file per image holding a
pickled dict {'prototype': float32 [M, D], 'prototype_label': int64 [M]}.
"""

from __future__ import annotations

import glob
import os

import numpy as np
import torch


def load_memory_banks(memory_dir, device=None):
  """spml/utils/segsort/others.py:11-41: concatenates every `*.npy` of `memory_dir` (sorted by
  name) into (prototypes float32 [M, D], prototype_labels int64 [M]).  `device` (extra
  keyword) moves the bank to the GPU it will be searched on."""
  paths = sorted(glob.glob(os.path.join(memory_dir, '*.npy')))
  assert len(paths) > 0, 'No memory stored in the directory'
  protos, labels = [], []
  for path in paths:
    entry = np.load(path, allow_pickle=True).item()
    protos.append(np.asarray(entry['prototype'], dtype=np.float32))
    labels.append(np.asarray(entry['prototype_label'], dtype=np.int64))
  protos = torch.from_numpy(np.concatenate(protos, 0))
  labels = torch.from_numpy(np.concatenate(labels, 0))
  if device is not None:
    protos, labels = protos.to(device), labels.to(device)
  return protos, labels


def save_memory_bank(path, prototypes, prototype_labels):
  """Writes one image's entry the way pyscripts/inference/prototype.py:207-211 does."""
  np.save(path, {'prototype': prototypes.detach().cpu().numpy().astype(np.float32),
                 'prototype_label': prototype_labels.detach().cpu().numpy().astype(np.int64)})
