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

from . import general_common, model_utils, predictions, segsort_common


def generate_clusters(embeddings, semantic_labels, instance_labels, local_features,
                      label_divisor, semantic_ignore_index, num_clusters, iterations,
                      batch_index_offset=None, densepose=False):
  """spml/models/embeddings/resnet_deeplab.py:90-148 / resnet_pspnet.py:90-148 (labels already
  at the embedding resolution).  The label packing (`sem * divisor + inst`, pixels with the
  ignore label dropped) and the decoding of the kept pixels' labels happen inside the one
  library call of segment_by_kmeans.  `densepose=True` adds the post-step of
  resnet_pspnet_densepose.py:141-154: cluster_embedding_with_loc becomes
  normalize(cat[0.1 * cluster_embedding, local features of the kept pixels])."""
  if semantic_labels is not None and instance_labels is not None:
    emb, emb_loc, _, cid, bid, sem, inst = segsort_common.segment_clusters(
        embeddings, None, num_clusters, local_features=local_features, iterations=iterations,
        batch_index_offset=batch_index_offset, semantic_labels=semantic_labels,
        instance_labels=instance_labels, label_divisor=label_divisor,
        semantic_ignore_index=semantic_ignore_index)
  else:
    emb, emb_loc, lab, cid, bid = segsort_common.segment_by_kmeans(
        embeddings, None, num_clusters, local_features=local_features, iterations=iterations,
        batch_index_offset=batch_index_offset)
    sem, inst = lab // label_divisor, lab % label_divisor
  if densepose and local_features is not None:
    loc = local_features.reshape(-1, local_features.shape[-1])
    if semantic_labels is not None:
      keep = (semantic_labels != semantic_ignore_index).view(-1).nonzero().view(-1)
      loc = torch.index_select(loc, 0, keep)
    emb_loc = general_common.normalize_embedding(torch.cat([emb * 0.1, loc], dim=-1))
  return {'cluster_embedding': emb, 'cluster_embedding_with_loc': emb_loc,
          'cluster_semantic_label': sem, 'cluster_instance_label': inst,
          'cluster_index': cid, 'cluster_batch_index': bid}


def generate_clusters_method(densepose=False):
  """A drop-in for the `generate_clusters` METHOD of the reference's embedding models
  (ResnetDeeplab / ResnetPspnet, and the DensePose ResnetPspnet with densepose=True); it reads
  the same four attributes of `self`."""
  def method(self, embeddings, semantic_labels, instance_labels, local_features=None):
    return generate_clusters(embeddings, semantic_labels, instance_labels, local_features,
                             self.label_divisor, self.semantic_ignore_index,
                             self.kmeans_num_clusters, self.kmeans_iterations,
                             densepose=densepose)
  method.__name__ = 'generate_clusters'
  return method


class ContrastiveHead(nn.Module):

  def __init__(self, config, softmax_classifier=False, variant=None, exchange_prototypes=False):
    """`variant`: 'segsort' (predictions/segsort.py), 'softmax' (segsort_softmax.py, the
    class train.py instantiates) or 'densepose' (resnet_pspnet_densepose.py clusters +
    segsort_softmax_densepose.py, train_densepose.py:159-205).
    `exchange_prototypes`: under torch.distributed (one process per GPU) contrast this rank's
    pixels with the prototypes of ALL ranks, gradients flowing back across ranks, which is
    what the reference's anchor-GPU gather computes (spml/models/utils.py:86-127); image tags
    are gathered like train.py:194-202.  Off: rank-local prototypes (north_star, PR1)."""
    super(ContrastiveHead, self).__init__()
    self.config = config
    self.exchange_prototypes = bool(exchange_prototypes)
    self.variant = variant or ('softmax' if softmax_classifier else 'segsort')
    self.predictor = {'segsort': predictions.Segsort, 'softmax': predictions.SegsortSoftmax,
                      'densepose': predictions.SegsortSoftmaxDensepose}[self.variant](config)
    self.memory_bank_size = int(getattr(config.train, 'memory_bank_size', 0))
    self.memory_banks = {}
    object.__setattr__(self, '_last_targets', None)
    self._last_batch = 0

  def forward(self, embedding, semantic_label, instance_label, semantic_tag,
              local_feature=None, semantic_label_full=None):
    """`semantic_label_full`: the full-resolution label map the classifier of the softmax
    variants is trained on (targets['semantic_label']); default: `semantic_label`."""
    cfg = self.config
    exchange = (self.exchange_prototypes and torch.distributed.is_available()
                and torch.distributed.is_initialized()
                and torch.distributed.get_world_size() > 1)
    rank = torch.distributed.get_rank() if exchange else 0
    datas = generate_clusters(
        embedding, semantic_label, instance_label, local_feature, cfg.network.label_divisor,
        cfg.dataset.semantic_ignore_index, cfg.network.kmeans_num_clusters,
        cfg.network.kmeans_iterations,
        # rank-local image indices (`semantic_tag` is this rank's) unless the prototypes of
        # all ranks are exchanged: then global ones, like common.py:376-377's N * gpu_id
        batch_index_offset=embedding.shape[0] * rank,
        densepose=self.variant == 'densepose')
    (protos, protos_loc, psem, pinst, pbid, cids) = (
        model_utils.gather_clustering_and_update_prototypes(
            [datas['cluster_embedding']], [datas['cluster_embedding_with_loc']],
            [datas['cluster_index']], [datas['cluster_batch_index']],
            [datas['cluster_semantic_label']], [datas['cluster_instance_label']]))
    if exchange:
      from . import distributed
      protos, protos_loc, psem, pinst, pbid, cids = [[t] for t in distributed.exchange_prototypes(
          protos[0], protos_loc[0], psem[0], pinst[0], pbid[0], cids[0])]
      semantic_tag = distributed.all_gather_rows(semantic_tag)          # train.py:194-198
    datas['cluster_index'] = cids[0]
    datas['embedding'] = embedding
    targets = {'prototype': protos[0], 'prototype_with_loc': protos_loc[0],
               'prototype_semantic_label': psem[0], 'prototype_instance_label': pinst[0],
               'prototype_batch_index': pbid[0],
               'semantic_label': semantic_label_full if semantic_label_full is not None
                                 else semantic_label}
    if self.variant != 'densepose':                 # train_densepose.py:189-199 (commented out)
      targets['semantic_tag'] = semantic_tag
      targets['prototype_semantic_tag'] = torch.index_select(semantic_tag, 0, pbid[0])
    targets.update(self.memory_banks)                                  # train.py:204-208
    out = self.predictor(datas, targets)
    total = getattr(self.predictor, 'last_loss_total', None)
    if total is not None:
      # the head kernel already added the losses it produced (same order, same fp32 adds)
      object.__setattr__(self.predictor, 'last_loss_total', None)
      out['loss'] = total
    else:
      losses = [out[k] for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss')
                if out.get(k, None) is not None]
      out['loss'] = sum(losses)                                        # train.py:213-219
    out['datas'], out['targets'] = datas, targets
    # (object.__setattr__: nn.Module.__setattr__ costs ~5 us per assignment)
    object.__setattr__(self, '_last_targets', targets)
    object.__setattr__(self, '_last_batch', embedding.shape[0])
    return out

  @torch.no_grad()
  def update_memory_bank(self, num_replicas=1):
    """train.py:276-293: FIFO of detached 'prototype*' entries; stored batch indices move up
    by batch_size * num_gpus every step.  The reference clones every entry; the tensors
    here are this step's own outputs, which nothing writes to afterwards, so the bank keeps
    them as they are (detached) and the batch indices are re-created, not updated in place."""
    if self._last_targets is None:
      return
    for k, v in self._last_targets.items():
      if 'prototype' in k and 'memory' not in k and torch.is_tensor(v):
        bank = self.memory_banks.setdefault('memory_' + k, [])
        bank.append(v.detach() if v.requires_grad else v)
        if len(bank) > self.memory_bank_size:
          self.memory_banks['memory_' + k] = bank[1:]
    key = 'memory_prototype_batch_index'
    if self.memory_banks.get(key):
      stride = self._last_batch * num_replicas
      # one multi-tensor launch for the whole list
      self.memory_banks[key] = list(torch._foreach_add(self.memory_banks[key], stride))
    object.__setattr__(self, '_last_targets', None)
