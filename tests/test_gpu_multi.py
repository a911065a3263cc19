"""Two-GPU NCCL test of the cross-GPU prototype exchange: with
ContrastiveHead(exchange_prototypes=True) every rank contrasts its pixels with the prototypes
of BOTH ranks and gradients cross ranks, which must equal what the reference's anchor-GPU
gather computes (spml/models/utils.py:41-131 over two-entry lists, train.py:167-219), here
restated by the CPU oracle.  Skipped on a single-GPU box."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import norm_err
from oracle import spml_oracle as O
from spml_b200 import synth

pytestmark = pytest.mark.gpu


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def oracle_two_devices(w, cfg, batches):
  """The reference's semantics for two GPUs in one process: per-device clustering with global
  image indices, ONE prototype set, per-device losses, summed."""
  embs, datas = [], []
  for d, batch in enumerate(batches):
    emb = batch['embedding'].clone().requires_grad_(True)
    sem, inst = batch['semantic_label'], batch['instance_label']
    lab = sem * cfg.network.label_divisor + inst
    ign = lab.max() + 1
    lab = lab.masked_fill(sem == cfg.dataset.semantic_ignore_index, ign)
    e, el, lab, cid, bid = O.segment_by_kmeans(
        emb, lab, cfg.network.kmeans_num_clusters, local_features=batch['local_feature'],
        ignore_index=ign, iterations=cfg.network.kmeans_iterations, device_index=d)
    embs.append(emb)
    datas.append({'cluster_embedding': e, 'cluster_embedding_with_loc': el,
                  'cluster_semantic_label': lab // cfg.network.label_divisor,
                  'cluster_instance_label': lab % cfg.network.label_divisor,
                  'cluster_index': cid, 'cluster_batch_index': bid})
  p, pl, psem, pinst, pbid, cids = O.gather_and_update_prototypes(
      [d['cluster_embedding'] for d in datas], [d['cluster_embedding_with_loc'] for d in datas],
      [d['cluster_index'] for d in datas], [d['cluster_batch_index'] for d in datas],
      [d['cluster_semantic_label'] for d in datas], [d['cluster_instance_label'] for d in datas])
  tags = torch.cat([b['semantic_tag'] for b in batches], 0)            # train.py:194-198
  losses = []
  for d in range(2):
    datas[d]['cluster_index'] = cids[d]
    targets = {'prototype': p[d], 'prototype_with_loc': pl[d], 'prototype_semantic_label': psem[d],
               'prototype_instance_label': pinst[d], 'prototype_batch_index': pbid[d],
               'semantic_tag': tags, 'prototype_semantic_tag': tags.index_select(0, pbid[d])}
    losses.append(O.segsort_losses(cfg, datas[d], targets)[:3])
  sum(sum(l) for l in losses).backward()
  return {'prototype': p[0].detach(), 'psem': psem[0], 'pbid': pbid[0],
          'cids': [c.detach() for c in cids], 'losses': [[float(x) for x in l] for l in losses],
          'grads': [e.grad for e in embs]}


def _worker(rank, port, out):
  os.environ.update(RANK=str(rank), WORLD_SIZE='2', LOCAL_RANK=str(rank),
                    MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
  from spml_b200.head import ContrastiveHead
  w = synth.WORKLOADS['small']
  cfg = synth.make_config(w)
  batches = [synth.make_batch(w, seed=300 + r) for r in range(2)]
  b = {k: v.cuda() for k, v in batches[rank].items()}
  head = ContrastiveHead(cfg, exchange_prototypes=True).cuda()
  emb = b['embedding'].clone().requires_grad_(True)
  o = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'], b['local_feature'])
  o['loss'].backward()
  torch.cuda.synchronize()
  out[rank] = {'prototype': o['targets']['prototype'].detach().cpu(),
               'psem': o['targets']['prototype_semantic_label'].cpu(),
               'pbid': o['targets']['prototype_batch_index'].cpu(),
               'cid': o['datas']['cluster_index'].cpu(),
               'losses': [float(o[k]) for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss')],
               'grad': emb.grad.cpu()}
  dist.destroy_process_group()


def test_prototype_exchange_matches_the_reference_gather():
  if torch.cuda.device_count() < 2:
    pytest.skip('needs two GPUs')
  w = synth.WORKLOADS['small']
  cfg = synth.make_config(w)
  want = oracle_two_devices(w, cfg, [synth.make_batch(w, seed=300 + r) for r in range(2)])
  port = _free_port()
  with mp.Manager() as m:
    out = m.dict()
    mp.spawn(_worker, args=(port, out), nprocs=2, join=True)
    got = [out[0], out[1]]
  for r in range(2):
    assert torch.equal(got[r]['psem'], want['psem']) and torch.equal(got[r]['pbid'], want['pbid'])
    assert torch.equal(got[r]['cid'], want['cids'][r])
    assert float((got[r]['prototype'] - want['prototype']).abs().max()) < 2e-6
    for a, b in zip(got[r]['losses'], want['losses'][r]):
      assert abs(a - b) <= 1e-3 * abs(b) + 1e-7, (r, got[r]['losses'], want['losses'][r])
    # d(L_0 + L_1) / d(embedding of rank r): includes the other rank's loss through the bank
    assert norm_err(got[r]['grad'], want['grads'][r]) < 1e-3
