"""Drives the REFERENCE's own code (the copy of twke18/SPML under baseline/_ref, installed by
scripts/install_reference.py) through one training-step section of
pyscripts/train/train.py:167-219,273-293 on the CPU.  Baseline infrastructure for
`bench.py --impl reference` / `cpu_baseline` only; the product never imports this.

Shims (none changes arithmetic): the copy carries the one-line
`device.index or 0` patch; `scatter_gather.gather` (asserts on CPU tensors) is replaced by
torch.cat; the config is a SimpleNamespace tree instead of easydict.
"""

import sys
import types

import torch


def make_runner(ref_root):
  if ref_root not in sys.path:
    sys.path.insert(0, ref_root)
  import spml.models.utils as MU
  MU.scatter_gather.gather = lambda xs, dev, dim=0: torch.cat(list(xs), dim)
  import spml.models.embeddings.resnet_deeplab as ED
  import spml.models.embeddings.resnet_pspnet_densepose as EP
  import spml.models.predictions.segsort as P
  import spml.models.predictions.segsort_softmax as PS
  import spml.models.predictions.segsort_softmax_densepose as PD
  models = {}

  def step(cfg, w, batch, bank):
    """`bank`: dict of lists ('memory_prototype', ...), updated in place (train.py:276-293)."""
    densepose = w.variant == 'densepose'
    if w.name not in models:
      mod = {'segsort': P, 'softmax': PS, 'densepose': PD}[w.variant]
      model = mod.segsort(cfg)
      if w.variant != 'segsort':
        model.eval()              # dropout off: the timed work is the same, results are stable
      models[w.name] = model
    model = models[w.name]
    me = types.SimpleNamespace(label_divisor=cfg.network.label_divisor,
                               semantic_ignore_index=cfg.dataset.semantic_ignore_index,
                               kmeans_num_clusters=cfg.network.kmeans_num_clusters,
                               kmeans_iterations=cfg.network.kmeans_iterations)
    emb = batch['embedding'].clone().requires_grad_(True)
    net = EP.ResnetPspnet if densepose else ED.ResnetDeeplab
    datas = net.generate_clusters(me, emb, batch['semantic_label'], batch['instance_label'],
                                  batch['local_feature'])
    p, pl, psl, pil, pbi, ci = MU.gather_clustering_and_update_prototypes(
        [datas['cluster_embedding']], [datas['cluster_embedding_with_loc']],
        [datas['cluster_index']], [datas['cluster_batch_index']],
        [datas['cluster_semantic_label']], [datas['cluster_instance_label']], 'cpu')
    datas['cluster_index'] = ci[0]
    targets = dict(prototype=p[0], prototype_with_loc=pl[0], prototype_semantic_label=psl[0],
                   prototype_instance_label=pil[0], prototype_batch_index=pbi[0])
    if not densepose:
      tags = MU.gather_and_update_datas([batch['semantic_tag']], 'cpu')[0]
      targets.update(semantic_tag=tags, prototype_semantic_tag=tags.index_select(0, pbi[0]))
    targets.update({k: list(v) for k, v in bank.items()})
    if w.variant != 'segsort':
      datas['embedding'] = emb
      targets['semantic_label'] = batch.get('semantic_label_full', batch['semantic_label']).clone()
    out = model(datas, targets)
    loss = sum(out[k] for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss')
               if out[k] is not None)
    model.zero_grad()
    loss.backward()
    with torch.no_grad():
      for k in list(targets.keys()):
        if 'prototype' in k and 'memory' not in k:
          key = 'memory_' + k
          bank.setdefault(key, []).append(targets[k].clone().detach())
          if len(bank[key]) > w.memory_bank_size:
            bank[key] = bank[key][1:]
      for t in bank.get('memory_prototype_batch_index', []):
        t += w.batch
    return {'loss': loss.detach(), 'grad_embedding': emb.grad}

  return step
