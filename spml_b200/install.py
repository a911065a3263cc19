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

from . import general_common, head, model_utils, predictions, segsort_common, segsort_eval
from . import segsort_loss

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
        'gather_multiset_labels_per_batch_by_nearest_neighbor':
            model_utils.gather_multiset_labels_per_batch_by_nearest_neighbor,
    },
    # A9: the generate_clusters METHOD of the embedding models ('Class.attr' rebinds a class
    # attribute); the backbone / head forward of those classes stays the reference's (cuDNN)
    'spml.models.embeddings.resnet_deeplab': {
        'ResnetDeeplab.generate_clusters': head.generate_clusters_method(),
    },
    'spml.models.embeddings.resnet_pspnet': {
        'ResnetPspnet.generate_clusters': head.generate_clusters_method(),
    },
    'spml.models.embeddings.resnet_pspnet_densepose': {
        'ResnetPspnet.generate_clusters': head.generate_clusters_method(densepose=True),
    },
    'spml.models.predictions.segsort': {
        'Segsort': predictions.Segsort,
        'segsort': predictions.segsort,
    },
    'spml.models.predictions.segsort_softmax': {
        'SegsortSoftmax': predictions.SegsortSoftmax,
        'segsort': predictions.segsort_softmax,
    },
    'spml.models.predictions.segsort_softmax_densepose': {
        'SegsortSoftmax': predictions.SegsortSoftmaxDensepose,
        'segsort': predictions.segsort_softmax_densepose,
    },
}


def _holder(mod, attr):
  """('Class.attr' -> the class object, 'attr'); plain names live on the module."""
  obj = mod
  parts = attr.split('.')
  for part in parts[:-1]:
    obj = getattr(obj, part)
  return obj, parts[-1]

_saved = {}


# modules whose bindings replace whole classes / methods of the reference (the fused stage
# calls); level='operators' leaves them alone, so the reference's OWN generate_clusters and
# losses() code runs on top of the rebound functions and loss classes
_CLASS_LEVEL = ('spml.models.embeddings.', 'spml.models.predictions.')


def install(strict=True, level='all'):
  """Returns the list of 'module.attr' names that were rebound.  level: 'all' (default) or
  'operators' (functions and loss classes only)."""
  if level not in ('all', 'operators'):
    raise ValueError("level must be 'all' or 'operators'")
  done = []
  for mod_name, attrs in BINDINGS.items():
    if level == 'operators' and mod_name.startswith(_CLASS_LEVEL):
      continue
    try:
      mod = importlib.import_module(mod_name)
    except ImportError:
      if strict:
        raise
      continue
    for attr, repl in attrs.items():
      key = mod_name + ':' + attr
      holder, name = _holder(mod, attr)
      if key not in _saved:
        _saved[key] = holder.__dict__[name] if isinstance(holder, type) else getattr(holder, name)
      setattr(holder, name, repl)
      done.append(mod_name + '.' + attr)
  return done


def uninstall():
  for key, orig in list(_saved.items()):
    mod_name, attr = key.split(':', 1)
    holder, name = _holder(importlib.import_module(mod_name), attr)
    setattr(holder, name, orig)
    del _saved[key]
