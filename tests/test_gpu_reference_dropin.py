"""The drop-in claim, demonstrated on a GPU with the reference's OWN code: the training-step
section of pyscripts/train/train.py:167-219 (generate_clusters of the embedding model ->
gather_clustering_and_update_prototypes -> gather_and_update_datas -> prediction model
forward -> backward, memory bank as train.py:276-293) runs twice on cuda:0 from the copy of
twke18/SPML under baseline/_ref (scripts/install_reference.py): once untouched, once after
spml_b200.install() has rebound the hot-path symbols inside the imported `spml.*` modules.
Same inputs; integer outputs must be identical, losses and gradients within 1e-3."""

import importlib
import os
import sys
import types

import pytest
import torch

from conftest import ROOT
from helpers import norm_err
from oracle import spml_oracle as O
from spml_b200 import synth

REF = os.path.join(ROOT, 'baseline', '_ref')
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'spml')),
                                 reason='baseline/_ref not installed (scripts/install_reference.py)')]


@pytest.fixture()
def reference():
  sys.path.insert(0, REF)
  try:
    yield
  finally:
    import spml_b200
    spml_b200.uninstall()
    sys.path.remove(REF)
    for k in [k for k in sys.modules if k == 'spml' or k.startswith('spml.')]:
      del sys.modules[k]


def train_steps(w, steps, rebound):
  """What train.py / train_densepose.py do per iteration, on one GPU, with the reference's
  modules as imported (rebound or not).  Returns per-step dicts of CPU tensors."""
  import spml_b200
  spml_b200.uninstall()
  if rebound:
    spml_b200.install(level=rebound)
  densepose = w.variant == 'densepose'
  emb_mod = importlib.import_module('spml.models.embeddings.' + (
      'resnet_pspnet_densepose' if densepose else 'resnet_deeplab'))
  emb_cls = emb_mod.ResnetPspnet if densepose else emb_mod.ResnetDeeplab
  model_utils = importlib.import_module('spml.models.utils')
  pred_mod = importlib.import_module('spml.models.predictions.' + {
      'segsort': 'segsort', 'softmax': 'segsort_softmax',
      'densepose': 'segsort_softmax_densepose'}[w.variant])
  cfg = synth.make_config(w)
  me = types.SimpleNamespace(label_divisor=cfg.network.label_divisor,
                             semantic_ignore_index=cfg.dataset.semantic_ignore_index,
                             kmeans_num_clusters=cfg.network.kmeans_num_clusters,
                             kmeans_iterations=cfg.network.kmeans_iterations)
  prediction_model = pred_mod.segsort(cfg).cuda()
  if w.variant != 'segsort':
    prediction_model.semantic_classifier.load_state_dict(O.make_classifier(cfg).state_dict())
    prediction_model.eval()
  memory_banks = {}
  results = []
  for step in range(steps):
    batch = {k: v.cuda() for k, v in synth.make_batch(w, step=step).items()}
    emb = batch['embedding'].clone().requires_grad_(True)
    # train.py:167 (the embedding model's forward ends in generate_clusters)
    datas = emb_cls.generate_clusters(me, emb, batch['semantic_label'], batch['instance_label'],
                                      batch['local_feature'])
    datas['embedding'] = emb
    # train.py:170-192
    (protos, protos_loc, psem, pinst, pbid, cids) = (
        model_utils.gather_clustering_and_update_prototypes(
            [datas['cluster_embedding']], [datas['cluster_embedding_with_loc']],
            [datas['cluster_index']], [datas['cluster_batch_index']],
            [datas['cluster_semantic_label']], [datas['cluster_instance_label']], 'cuda:0'))
    labels = {'semantic_label': batch.get('semantic_label_full', batch['semantic_label']),
              'prototype': protos[0], 'prototype_with_loc': protos_loc[0],
              'prototype_semantic_label': psem[0], 'prototype_instance_label': pinst[0],
              'prototype_batch_index': pbid[0]}
    datas['cluster_index'] = cids[0]
    if not densepose:                                          # train.py:194-202
      tags = model_utils.gather_and_update_datas([batch['semantic_tag']], 'cuda:0')
      labels['semantic_tag'] = tags[0]
      labels['prototype_semantic_tag'] = torch.index_select(tags[0], 0, pbid[0])
    for k in memory_banks:                                     # train.py:204-208
      labels[k] = [m.to('cuda:0') for m in memory_banks[k]]
    outputs = prediction_model(datas, labels)                  # train.py:211
    losses = [outputs[k] for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss')
              if outputs.get(k) is not None]
    loss = sum(losses)                                         # train.py:213-219
    prediction_model.zero_grad()
    loss.backward()                                            # train.py:273
    res = {k: datas[k].detach().cpu() for k in datas if k != 'embedding'}
    res.update({k: v.detach().cpu() for k, v in labels.items()
                if torch.is_tensor(v) and k.startswith('prototype')})
    res.update({k: (outputs[k].detach().cpu() if outputs.get(k) is not None else None)
                for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss', 'accuracy')})
    res['loss'] = loss.detach().cpu()
    res['grad_embedding'] = emb.grad.detach().cpu()
    results.append(res)
    with torch.no_grad():                                      # train.py:276-293
      for k in labels:
        if 'prototype' in k and 'memory' not in k:
          memory = labels[k].clone().detach()
          key = 'memory_' + k
          memory_banks.setdefault(key, []).append(memory)
          if len(memory_banks[key]) > w.memory_bank_size:
            memory_banks[key] = memory_banks[key][1:]
      for t in memory_banks.get('memory_prototype_batch_index', []):
        t += w.batch
  return results


@pytest.mark.parametrize('name,steps', [('small', 3), ('tiny_softmax', 2), ('tiny_densepose', 2),
                                        ('voc_scribble_b1', 2)])
def test_reference_code_on_rebound_symbols(reference, name, steps):
  w = synth.WORKLOADS[name]
  plain = train_steps(w, steps, rebound=None)
  # 'operators': the reference's own generate_clusters / losses() code on the rebound functions
  # and loss classes; 'all': the fused stage calls behind the reference's class names
  for level in ('operators', 'all'):
    ours = train_steps(w, steps, rebound=level)
    import spml.utils.segsort.common as common
    assert common.segment_by_kmeans.__module__ == 'spml_b200.segsort_common'
    for step, (a, b) in enumerate(zip(plain, ours)):
      for k, v in a.items():
        what = '%s step %d %s' % (level, step, k)
        if v is None:
          assert b[k] is None, what
        elif not v.is_floating_point():
          assert torch.equal(v, b[k]), what
        elif v.dim() == 0:
          assert abs(float(v) - float(b[k])) <= 1e-3 * abs(float(v)) + 1e-6, (what, v, b[k])
        else:
          assert norm_err(b[k], v) < 1e-3, what
  from spml_b200 import ops
  ops.check_status()
