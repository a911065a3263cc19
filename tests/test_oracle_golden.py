"""Pins the CPU oracle on outputs of the reference itself (tests/golden/*.pt,
made by tests/golden/make_golden.py from /root/reference).  Same ATen kernels in
the same order => the oracle must match bit for bit on the CPU."""

import pytest
import torch

from conftest import load_golden
from oracle import spml_oracle as O
from spml_b200 import synth


def same(a, b):
  assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
  assert torch.equal(a, b), (a - b).abs().max() if a.is_floating_point() else (a != b).sum()


def test_normalize(units):
  same(O.l2_normalize(units['normalize']['x']), units['normalize']['y'])


def test_location_and_seeds(units):
  same(O.location_grid((5, 7), 'cpu', 'float'), units['location_float'])
  same(O.location_grid((5, 7), 'cpu', 'int'), units['location_int'])
  with pytest.raises(ValueError):
    O.location_grid((5, 7), 'cpu', 'bogus')
  for key, val in units.items():
    if key.startswith('seeds_'):
      a, b = key[len('seeds_'):].split('_')
      nc = [int(v) for v in a.split('x')]
      hw = [int(v) for v in b.split('x')]
      same(O.grid_seed_labels(nc, hw), val)


def test_prototypes_and_nearest(units):
  u = units['prototypes']
  same(O.prototypes_from_labels(u['e'], u['lab'], 6), u['p'])
  same(O.prototypes_from_labels(u['e'], u['lab']), u['p_auto'])
  assert u['p'][3].abs().max() == 0 and u['p'][5].abs().max() == 0   # empty labels
  t = units['nearest_ties']
  same(O.nearest_prototype(t['e'], t['p']), t['idx'])
  assert t['idx'].tolist()[:2] == [1, 0]


def test_kmeans_empty_cluster(units):
  u = units['kmeans_empty']
  same(O.spherical_kmeans(u['e'], u['lab0'], 4, 10), u['lab'])


def test_prototype_labels(units):
  u = units['prototype_labels']
  plab, inv = O.prototype_labels(u['sem'], u['inst'], 256)
  same(plab, u['plab'])
  same(inv, u['inv'])


def test_segsort_loss_and_grads(units):
  u = units['segsort']
  e = u['e'].clone().requires_grad_(True)
  p = u['p'].clone().requires_grad_(True)
  loss = O.segsort_loss(e, u['sem'], u['seg'], p, u['psem'], u['kappa'])
  loss.backward()
  same(loss.detach(), u['loss'])
  same(e.grad, u['de'])
  same(p.grad, u['dp'])
  same(O.segsort_nll(u['e'], u['sem'], u['seg'], u['p'], u['psem'], u['kappa']), u['nll'])


def test_set_segsort_loss_and_grads(units):
  u = units['set_segsort']
  e = u['e'].clone().requires_grad_(True)
  p = u['p'].clone().requires_grad_(True)
  loss = O.set_segsort_loss(e, u['tags'], u['seg'], p, u['ptags'], u['kappa'])
  loss.backward()
  same(loss.detach(), u['loss'])
  same(e.grad, u['de'])
  same(p.grad, u['dp'])


def test_topk(units):
  u = units['topk']
  acc, lab = O.top_k_ranking(u['q'], u['ql'], u['p'], u['pl'], 5)
  same(acc, u['acc'])
  same(lab, u['labels'])


def test_segment_by_kmeans_edges(units):
  u = units['segment_ignore_image']
  got = O.segment_by_kmeans(u['emb'], u['labels'], [2, 2], ignore_index=7, iterations=3)
  for a, k in zip(got, ('ce', 'cel', 'cl', 'ci', 'cb')):
    same(a, u[k])
  assert (u['cb'] == 0).all()            # image 1 contributed no pixels
  u = units['segment_user_clusters']
  got = O.segment_by_kmeans(u['emb'], u['labels'], [2, 2], cluster_indices=u['cmap'],
                            iterations=2)
  for a, k in zip(got, ('ce', 'cel', 'cl', 'ci', 'cb')):
    same(a, u[k])


@pytest.fixture(scope='module')
def units2():
  return load_golden('units2.pt')


def test_nn_multiset_labels(units2):
  """models/utils.py:157-223 (DensePose tag propagation)."""
  for key in ('nn_tags_k1', 'nn_tags_k3'):
    u = units2[key]
    same(O.nn_multiset_labels(u['p'], u['p'], u['psem'], u['pbid'], u['pbid'], u['num_classes'],
                              u['top_k'], u['threshold']), u['tags'])
  u = units2['nn_tags_k1']
  assert int(u['tags'][u['pbid'] == 2].sum()) == 0       # image without a labelled prototype
  assert int(u['tags'].sum()) > 0
  u = units2['nn_tags_queries']
  same(O.nn_multiset_labels(u['q'], u['p'], u['psem'], u['qbid'], u['pbid'], u['num_classes'],
                            u['top_k'], u['threshold']), u['tags'])


def test_predictions(units2):
  """segsort.py:68-125 (retrieval inference)."""
  u = units2['predictions']
  pred, topk = O.segsort_predictions(
      {'cluster_embedding': u['emb'], 'cluster_index': u['cid']},
      {'semantic_memory_prototype': u['bank'], 'semantic_memory_prototype_label': u['bank_label']})
  same(pred, u['pred'])
  same(topk, u['topk'])


def test_segment_mean_one_hot_gather(units2):
  u = units2['segment_mean']
  same(O.segment_mean(u['x'], u['index']), u['mean'])
  assert u['mean'][4].abs().max() == 0                   # empty segment: 0 / 1
  u = units2['one_hot']
  same(O.one_hot(u['labels']), u['auto'])
  same(O.one_hot(u['labels'], 8), u['wide'])
  u = units2['gather_datas']
  for t in u['out']:
    same(t, torch.cat(u['in'], 0))
  u = units2['gather_two_devices']
  L = u['lists']
  out = O.gather_and_update_prototypes(
      L['cluster_embedding'], L['cluster_embedding_with_loc'], L['cluster_index'],
      L['cluster_batch_index'], L['cluster_semantic_label'], L['cluster_instance_label'])
  for got, key in zip(out, ('prototype', 'prototype_with_loc', 'prototype_semantic_label',
                            'prototype_instance_label', 'prototype_batch_index',
                            'cluster_index')):
    for a, b in zip(got, u['out'][key]):
      same(a, b)


@pytest.mark.parametrize('name,steps', [('tiny', 3), ('small', 3), ('tiny_softmax', 3),
                                        ('tiny_densepose', 2), ('tiny_densepose_shipped', 1)])
def test_full_step_matches_reference(name, steps):
  """Consecutive steps (memory bank filling up) of the whole path, for the three heads:
  segsort.py, segsort_softmax.py (eval mode) and the DensePose variant."""
  w = synth.WORKLOADS[name]
  cfg = synth.make_config(w)
  torch.set_num_threads(1)
  bank = {}
  classifier = O.make_classifier(cfg).eval() if w.variant != 'segsort' else None
  for step in range(steps):
    g = load_golden('%s_step%d.pt' % (name, step))
    batch = synth.make_batch(w, seed=g['meta']['seed'], step=step)
    for k, v in g['inputs'].items():      # the generator is reproducible
      same(batch[k], v)
    for k, v in g['bank'].items():
      assert len(v) == len(bank[k])
      for a, b in zip(bank[k], v):
        same(a, b)
    out = O.contrastive_step(cfg, batch, bank, variant=w.variant, classifier=classifier)
    for k, v in g['outputs'].items():
      if k == 'cluster_index_before_gather':
        continue
      if k == 'grad_classifier':
        for name_, grad in v.items():
          same(out[k][name_], grad)
        continue
      same(out[k], v)
    targets = {k: out[k] for k in out if k.startswith('prototype')}
    O.memory_bank_update(bank, targets, w.memory_bank_size, w.batch)


def test_fp64_oracle_close_to_fp32():
  """The double-precision run of the same algorithm bounds the fp32 rounding of
  the reference: losses agree to ~1e-6 relative when the segment ids agree."""
  w = synth.WORKLOADS['tiny']
  cfg = synth.make_config(w)
  batch = synth.make_batch(w)
  a = O.contrastive_step(cfg, batch)
  b = O.contrastive_step(cfg, batch, dtype=torch.float64)
  if torch.equal(a['cluster_index'], b['cluster_index']):
    for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss'):
      assert abs(a[k].item() - b[k].item()) <= 1e-5 * max(1.0, abs(b[k].item()))
