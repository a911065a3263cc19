"""Rebinds the reference's hot-path symbols to the B200 implementations.

    import spml_b200
    spml_b200.install()          # needs twke18/SPML importable as `spml`

After this, `pyscripts/train/train.py` of the reference runs the contrastive head on
libspml_b200.so without a single changed line: its call sites use module-attribute
lookups (`segsort_common.segment_by_kmeans(...)`, resnet_deeplab.py:128;
`segsort_loss.SegSortLoss(...)`, segsort_softmax.py:76;
`model_utils.gather_clustering_and_update_prototypes(...)`, train.py:180), so
replacing the attributes is sufficient.  `uninstall()` restores the originals.
"""

from __future__ import annotations

import importlib

from . import general_common, model_utils, predictions, segsort_common, segsort_eval, segsort_loss

# reference module -> {attribute: replacement}
BINDINGS = {
    'spml.utils.general.common': {
        'normalize_embedding': general_common.normalize_embedding,
    },
    'spml.utils.segsort.common': {
        'calculate_prototypes_from_labels': segsort_common.calculate_prototypes_from_labels,
        'find_nearest_prototypes': segsort_common.find_nearest_prototypes,
        'kmeans_with_initial_labels': segsort_common.kmeans_with_initial_labels,
        'prepare_prototype_labels': segsort_common.prepare_prototype_labels,
        'segment_by_kmeans': segsort_common.segment_by_kmeans,
    },
    'spml.utils.segsort.loss': {
        'SegSortLoss': segsort_loss.SegSortLoss,
        'SetSegSortLoss': segsort_loss.SetSegSortLoss,
    },
    'spml.utils.segsort.eval': {
        'top_k_ranking': segsort_eval.top_k_ranking,
    },
    'spml.models.utils': {
        'gather_clustering_and_update_prototypes':
            model_utils.gather_clustering_and_update_prototypes,
        'gather_and_update_datas': model_utils.gather_and_update_datas,
    },
    'spml.models.predictions.segsort': {
        'Segsort': predictions.Segsort,
        'segsort': predictions.segsort,
    },
    'spml.models.predictions.segsort_softmax': {
        'SegsortSoftmax': predictions.SegsortSoftmax,
        'segsort': predictions.segsort_softmax,
    },
}

_saved = {}


def install(strict=True):
  """Returns the list of 'module.attr' names that were rebound."""
  done = []
  for mod_name, attrs in BINDINGS.items():
    try:
      mod = importlib.import_module(mod_name)
    except ImportError:
      if strict:
        raise
      continue
    for attr, repl in attrs.items():
      key = mod_name + '.' + attr
      if key not in _saved:
        _saved[key] = getattr(mod, attr)
      setattr(mod, attr, repl)
      done.append(key)
  return done


def uninstall():
  for key, orig in list(_saved.items()):
    mod_name, attr = key.rsplit('.', 1)
    setattr(importlib.import_module(mod_name), attr, orig)
    del _saved[key]
