"""The contrastive head as one object: what pyscripts/train/train.py:167-219,276-293
does between the backbone's `embedding` tensor and the scalar loss, on one device.

    head = ContrastiveHead(config)
    out = head(embedding, semantic_label, instance_label, semantic_tag, local_feature)
    out['loss'].backward()          # d loss / d embedding re-enters the backbone
    head.update_memory_bank()       # train.py:276-293

It only composes the reference-named operators of this package (generate_clusters
-> gather_clustering_and_update_prototypes -> Segsort.forward), so it is also the
shortest description of how the pieces fit.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from . import model_utils, predictions, segsort_common


def generate_clusters(embeddings, semantic_labels, instance_labels, local_features,
                      label_divisor, semantic_ignore_index, num_clusters, iterations,
                      batch_index_offset=None):
  """spml/models/embeddings/resnet_deeplab.py:90-148 (labels already at the embedding
  resolution).  The ignore id `labels.max() + 1` stays on the device."""
  if semantic_labels is not None and instance_labels is not None:
    labels = semantic_labels * label_divisor + instance_labels
    ignore_index = labels.max() + 1
    # torch.where, not masked_fill: a tensor fill value would cost a device->host read-back
    labels = torch.where(semantic_labels == semantic_ignore_index, ignore_index, labels)
  else:
    labels, ignore_index = None, None
  emb, emb_loc, lab, cid, bid = segsort_common.segment_by_kmeans(
      embeddings, labels, num_clusters, local_features=local_features,
      ignore_index=ignore_index, iterations=iterations, batch_index_offset=batch_index_offset)
  return {'cluster_embedding': emb, 'cluster_embedding_with_loc': emb_loc,
          'cluster_semantic_label': lab // label_divisor,
          'cluster_instance_label': lab % label_divisor,
          'cluster_index': cid, 'cluster_batch_index': bid}


class ContrastiveHead(nn.Module):

  def __init__(self, config, softmax_classifier=False):
    super(ContrastiveHead, self).__init__()
    self.config = config
    self.predictor = (predictions.SegsortSoftmax(config) if softmax_classifier
                      else predictions.Segsort(config))
    self.memory_bank_size = int(getattr(config.train, 'memory_bank_size', 0))
    self.memory_banks = {}
    self._last_targets = None
    self._last_batch = 0

  def forward(self, embedding, semantic_label, instance_label, semantic_tag,
              local_feature=None):
    cfg = self.config
    datas = generate_clusters(
        embedding, semantic_label, instance_label, local_feature, cfg.network.label_divisor,
        cfg.dataset.semantic_ignore_index, cfg.network.kmeans_num_clusters,
        cfg.network.kmeans_iterations,
        batch_index_offset=0)   # rank-local image indices: `semantic_tag` is this rank's
    (protos, protos_loc, psem, pinst, pbid, cids) = (
        model_utils.gather_clustering_and_update_prototypes(
            [datas['cluster_embedding']], [datas['cluster_embedding_with_loc']],
            [datas['cluster_index']], [datas['cluster_batch_index']],
            [datas['cluster_semantic_label']], [datas['cluster_instance_label']]))
    datas['cluster_index'] = cids[0]
    datas['embedding'] = embedding
    targets = {'prototype': protos[0], 'prototype_with_loc': protos_loc[0],
               'prototype_semantic_label': psem[0], 'prototype_instance_label': pinst[0],
               'prototype_batch_index': pbid[0], 'semantic_tag': semantic_tag,
               'semantic_label': semantic_label,
               'prototype_semantic_tag': torch.index_select(semantic_tag, 0, pbid[0])}
    targets.update(self.memory_banks)                                  # train.py:204-208
    out = self.predictor(datas, targets)
    losses = [out[k] for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss')
              if out.get(k, None) is not None]
    out['loss'] = sum(losses)                                          # train.py:213-219
    out['datas'], out['targets'] = datas, targets
    self._last_targets, self._last_batch = targets, embedding.shape[0]
    return out

  @torch.no_grad()
  def update_memory_bank(self, num_replicas=1):
    """train.py:276-293: FIFO of detached 'prototype*' entries; stored batch indices
    move up by batch_size * num_gpus every step."""
    if self._last_targets is None:
      return
    for k, v in self._last_targets.items():
      if 'prototype' in k and 'memory' not in k and torch.is_tensor(v):
        bank = self.memory_banks.setdefault('memory_' + k, [])
        bank.append(v.clone().detach())
        if len(bank) > self.memory_bank_size:
          self.memory_banks['memory_' + k] = bank[1:]
    for t in self.memory_banks.get('memory_prototype_batch_index', []):
      t += self._last_batch * num_replicas
    self._last_targets = None
