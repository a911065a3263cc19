"""GPU parity tests of the individual operators (reference-named API) against the
golden unit vectors made by the reference and against the CPU oracle."""

import ctypes
import os

import pytest
import torch

from helpers import norm_err
from oracle import spml_oracle as O
from spml_b200 import _lib, general_common, model_utils, ops, segsort_common, segsort_eval
from spml_b200 import segsort_loss, synth

pytestmark = pytest.mark.gpu


def cu(t):
  return t.cuda()


def close(a, b, tol=2e-6):
  a = a.cpu()
  assert a.shape == b.shape, (a.shape, b.shape)
  err = float((a - b).abs().max()) if a.numel() else 0.0
  assert err <= tol, err


def test_normalize_embedding(units):
  u = units['normalize']
  x = cu(u['x']).requires_grad_(True)
  y = general_common.normalize_embedding(x)
  close(y.detach(), u['y'], 1e-7)
  g = torch.randn(u['x'].shape, generator=torch.Generator().manual_seed(1))
  y.backward(cu(g))
  xr = u['x'].clone().requires_grad_(True)
  O.l2_normalize(xr).backward(g)
  rows = [0, 1, 3, 5]          # rows 2 / 4 sit on the clamped branch (scale 1e12)
  close(x.grad.cpu()[rows], xr.grad[rows], 1e-5)
  assert torch.allclose(x.grad.cpu()[[2, 4]], xr.grad[[2, 4]], rtol=1e-5)


def test_prototypes_from_labels(units):
  u = units['prototypes']
  close(segsort_common.calculate_prototypes_from_labels(cu(u['e']), cu(u['lab']), 6), u['p'])
  close(segsort_common.calculate_prototypes_from_labels(cu(u['e']), cu(u['lab'])), u['p_auto'])


def test_prototype_backward_matches_autograd():
  g = torch.Generator().manual_seed(3)
  e = O.l2_normalize(torch.randn(500, 66, generator=g))
  seg = torch.randint(0, 17, (500,), generator=g)
  w = torch.randn(19, 66, generator=g)
  er = e.clone().requires_grad_(True)
  (O.prototypes_from_labels(er, seg, 19) * w).sum().backward()
  ec = cu(e).requires_grad_(True)
  (segsort_common.calculate_prototypes_from_labels(ec, cu(seg), 19) * cu(w)).sum().backward()
  assert norm_err(ec.grad.cpu(), er.grad) < 1e-5


def test_nearest_prototype_ties(units):
  t = units['nearest_ties']
  got = segsort_common.find_nearest_prototypes(cu(t['e']), cu(t['p'])).cpu()
  assert torch.equal(got, t['idx'])


@pytest.mark.parametrize('path', ['', 'cluster', 'small', 'tc', 'fp32'])
def test_kmeans_with_empty_cluster(units, path, monkeypatch):
  if path:
    monkeypatch.setenv('SPML_B200_KMEANS', path)
  u = units['kmeans_empty']
  got = segsort_common.kmeans_with_initial_labels(cu(u['e']), cu(u['lab0']), 4, 10).cpu()
  assert torch.equal(got, u['lab'])


@pytest.mark.parametrize('path', ['cluster', 'small', 'fp32'])
def test_kmeans_empty_cluster_next_to_near_zero_scores(path, monkeypatch):
  """An empty cluster is the zero vector (score exactly 0); pixels that are almost orthogonal to
  every live prototype are ambiguous against it and go through the exact re-check, which must
  see zeros for the empty cluster (the cluster kernel keeps those rows in a scratch that outlives
  the call).  Every kernel against the oracle."""
  monkeypatch.setenv('SPML_B200_KMEANS', path)
  g = torch.Generator().manual_seed(3)
  dim, k, n = 16, 6, 3000
  e = torch.zeros(n, dim)
  lab0 = torch.zeros(n, dtype=torch.long)
  for c in range(3):                                   # three live clusters along axes 0..2
    rows = slice(c * 900, (c + 1) * 900)
    e[rows, c] = 1.0
    e[rows] += 0.05 * torch.randn(900, dim, generator=g)
    lab0[rows] = c
  rest = slice(2700, n)                                # pixels in the orthogonal complement
  e[rest, 8:] = torch.randn(n - 2700, dim - 8, generator=g)
  e[rest, :3] = 1e-4 * torch.randn(n - 2700, 3, generator=g)
  lab0[rest] = torch.randint(0, 3, (n - 2700,), generator=g)
  e = O.l2_normalize(e)
  # fill the scratch of an earlier call with other prototypes first
  segsort_common.kmeans_with_initial_labels(cu(O.l2_normalize(torch.randn(n, dim, generator=g))),
                                            cu(torch.randint(0, k, (n,), generator=g)), k, 3)
  want = O.spherical_kmeans(e, lab0, k, 10)
  got = segsort_common.kmeans_with_initial_labels(cu(e), cu(lab0), k, 10).cpu()
  assert int((got != want).sum()) == 0


@pytest.mark.parametrize('n,dim,k', [(5000, 66, 36), (3000, 37, 144), (4097, 130, 100)])
def test_kmeans_matches_oracle(n, dim, k):
  g = torch.Generator().manual_seed(n + dim)
  centres = torch.randn(k, dim, generator=g)
  cell = torch.arange(n) * k // n
  e = O.l2_normalize(centres[cell] + 0.7 * torch.randn(n, dim, generator=g))
  lab0 = cell[torch.randperm(n, generator=g)]
  want = O.spherical_kmeans(e, lab0, k, 10)
  got = segsort_common.kmeans_with_initial_labels(cu(e), cu(lab0), k, 10).cpu()
  assert int((got != want).sum()) == 0


@pytest.mark.parametrize('n,dim,k,batch', [(5000, 66, 36, 1), (20000, 66, 36, 3), (3000, 37, 300, 1),
                                           (60000, 66, 36, 4), (37636, 37, 128, 1), (70000, 130, 64, 2),
                                           (16000, 66, 64, 2), (700, 5, 7, 3),
                                           (9000, 128, 1000, 1), (4097, 64, 129, 2),
                                           (130, 16, 7, 1),
                                           # several tiles per CTA AND several prototype tiles per
                                           # pixel tile (two-stage ring, both prefetch paths)
                                           (30000, 37, 300, 2), (25000, 66, 200, 1),
                                           # one thread-block cluster per image (kmeans_cluster.cu):
                                           # two accumulator rounds (K > 64), 5 channel slots,
                                           # small clusters, an image with no rows
                                           (16000, 37, 128, 1), (3000, 130, 40, 1), (2000, 66, 36, 2),
                                           (16384, 66, 36, 1), (9, 66, 36, 4), (28000, 66, 64, 2)])
def test_kmeans_tensor_core_equals_fp32(n, dim, k, batch, monkeypatch):
  """The tcgen05 E-step (with its exact re-check of near-ties) and the fp32 CUDA-core
  E-step return identical labels, also on data with no cluster structure (many near-ties)."""
  g = torch.Generator().manual_seed(7 * n + dim)
  for noise in (0.7, 30.0):
    centres = torch.randn(k, dim, generator=g)
    cell = torch.arange(n) * k // n
    e = O.l2_normalize(centres[cell] + noise * torch.randn(n, dim, generator=g)).cuda()
    lab0 = cell[torch.randperm(n, generator=g)].to(torch.int32).cuda()
    per = (n + batch - 1) // batch
    img_off = torch.tensor([min(i * per, n) for i in range(batch + 1)], dtype=torch.int32).cuda()
    got = {}
    paths = ('fp32', 'tc') + (('small', 'cluster') if k <= 128 else ())   # kmeans_small / _cluster.cu
    for path in paths:
      monkeypatch.setenv('SPML_B200_KMEANS', path)
      got[path], _ = ops.kmeans(e, img_off, batch, per, k, 10, lab0, want_i64=False)
      torch.cuda.synchronize()
    for path in paths[1:]:
      assert torch.equal(got['fp32'], got[path]), (path, int((got['fp32'] != got[path]).sum()))


def test_prepare_prototype_labels(units):
  u = units['prototype_labels']
  plab, inv = segsort_common.prepare_prototype_labels(cu(u['sem']), cu(u['inst']), 256)
  assert torch.equal(plab.cpu(), u['plab']) and torch.equal(inv.cpu(), u['inv'])


def test_unique_inverse_random():
  g = torch.Generator().manual_seed(9)
  for n, hi_max, lo_max in ((1, 1, 1), (1000, 7, 50), (70000, 300, 2000), (257, 1, 100000)):
    hi = torch.randint(0, hi_max, (n,), generator=g)
    lo = torch.randint(0, lo_max, (n,), generator=g)
    keys, inv = torch.unique(hi * (lo.max() + 1) + lo, return_inverse=True)
    got_inv, uhi, ulo, count, bound = ops.unique_inverse(cu(lo), hi=cu(hi), bound=0)
    m = int(count)
    assert m == keys.numel() and int(bound) == int(lo.max()) + 1
    assert torch.equal(got_inv.cpu(), inv)
    assert torch.equal((uhi[:m] * bound + ulo[:m]).cpu(), keys)


def test_segsort_loss(units):
  u = units['segsort']
  e = cu(u['e']).requires_grad_(True)
  p = cu(u['p']).requires_grad_(True)
  loss = segsort_loss.SegSortLoss(u['kappa'])(e, cu(u['sem']), cu(u['seg']), p, cu(u['psem']))
  loss.backward()
  assert abs(float(loss) - float(u['loss'])) <= 1e-5 * abs(float(u['loss']))
  assert norm_err(e.grad.cpu(), u['de']) < 1e-4
  assert norm_err(p.grad.cpu(), u['dp']) < 1e-4


def test_set_segsort_loss(units):
  u = units['set_segsort']
  e = cu(u['e']).requires_grad_(True)
  p = cu(u['p']).requires_grad_(True)
  loss = segsort_loss.SetSegSortLoss(u['kappa'])(e, cu(u['tags']), cu(u['seg']), p,
                                                cu(u['ptags']))
  loss.backward()
  assert abs(float(loss) - float(u['loss'])) <= 1e-5 * abs(float(u['loss']))
  assert norm_err(e.grad.cpu(), u['de']) < 1e-4
  assert norm_err(p.grad.cpu(), u['dp']) < 1e-4


@pytest.mark.parametrize('n,m,dim,kappa', [(1000, 200, 64, 6.0), (777, 65, 66, 16.0),
                                           (300, 1500, 130, 12.0), (129, 3, 37, 10.0)])
def test_segsort_loss_random(n, m, dim, kappa):
  g = torch.Generator().manual_seed(n * 7 + m)
  protos = O.l2_normalize(torch.randn(m, dim, generator=g))
  seg = torch.randint(0, m, (n,), generator=g)
  psem = torch.randint(0, 5, (m,), generator=g)
  e = O.l2_normalize(protos[seg] + 0.5 * torch.randn(n, dim, generator=g))
  er, pr = e.clone().requires_grad_(True), protos.clone().requires_grad_(True)
  want = O.segsort_loss(er.double(), psem[seg], seg, pr.double(), psem, kappa)
  want.backward()
  ec, pc = cu(e).requires_grad_(True), cu(protos).requires_grad_(True)
  got = segsort_loss.SegSortLoss(kappa)(ec, cu(psem[seg]), cu(seg), pc, cu(psem))
  got.backward()
  assert abs(float(got) - float(want)) <= 2e-5 * abs(float(want))
  assert norm_err(ec.grad.cpu(), er.grad) < 1e-4
  assert norm_err(pc.grad.cpu(), pr.grad) < 1e-4


def test_topk(units):
  u = units['topk']
  acc, lab = segsort_eval.top_k_ranking(cu(u['q']), cu(u['ql']), cu(u['p']), cu(u['pl']), 5)
  assert torch.equal(lab.cpu(), u['labels'])
  assert abs(float(acc) - float(u['acc'])) < 1e-6
  maj = segsort_eval.majority_label_from_topk(lab, 3)
  assert torch.equal(maj.cpu(), u['majority'])


@pytest.mark.parametrize('m,dim,nq', [(700, 64, 97), (40000, 32, 50), (1500, 130, 40)])
def test_topk_masks_and_cluster_split(m, dim, nq):
  """Fixed-capacity banks: dead queries / prototypes are skipped, the bank is split over the
  8-CTA cluster (m = 40 000 also exceeds the per-CTA live-tile table), and the merged top-5
  equals a full sort over the live prototypes."""
  g = torch.Generator().manual_seed(m + dim)
  p = O.l2_normalize(torch.randn(m, dim, generator=g))
  q = O.l2_normalize(torch.randn(nq, dim, generator=g))
  pl, ql = torch.randint(0, 21, (m,), generator=g), torch.randint(0, 21, (nq,), generator=g)
  pvalid = (torch.rand(m, generator=g) < 0.3)
  pvalid[m // 3: m // 2] = False                 # whole dead column tiles
  qvalid = (torch.rand(nq, generator=g) < 0.8)
  live = torch.nonzero(pvalid).flatten()
  sim = q[qvalid] @ p[live].t()
  order = torch.argsort(sim, dim=1, descending=True, stable=True)[:, :5]
  want = pl[live][order]
  acc, lab = ops.topk_ranking(cu(q), cu(ql), cu(p), cu(pl), 5, qvalid=cu(qvalid.to(torch.uint8)),
                              pvalid=cu(pvalid.to(torch.uint8)))
  got = lab.cpu()[qvalid]
  assert float((got != want).float().mean()) < 2e-3      # near-ties may swap
  want_acc = (want == ql[qvalid][:, None]).float().mean()
  assert abs(float(acc) - float(want_acc)) < 2e-3


def test_topk_large_k20():
  g = torch.Generator().manual_seed(11)
  q = O.l2_normalize(torch.randn(333, 64, generator=g))
  p = O.l2_normalize(torch.randn(5000, 64, generator=g))
  ql, pl = torch.randint(0, 21, (333,), generator=g), torch.randint(0, 21, (5000,), generator=g)
  acc_w, lab_w = O.top_k_ranking(q, ql, p, pl, 20)
  acc, lab = segsort_eval.top_k_ranking(cu(q), cu(ql), cu(p), cu(pl), 20)
  assert float((lab.cpu() != lab_w).float().mean()) < 1e-3    # near-ties may swap
  assert abs(float(acc) - float(acc_w)) < 1e-3


def test_segment_by_kmeans_edges(units):
  u = units['segment_ignore_image']
  got = segsort_common.segment_by_kmeans(cu(u['emb']), cu(u['labels']), [2, 2], ignore_index=7,
                                         iterations=3)
  for a, k in zip(got, ('ce', 'cel', 'cl', 'ci', 'cb')):
    if a.is_floating_point():
      close(a.detach(), u[k])
    else:
      assert torch.equal(a.cpu(), u[k]), k
  u = units['segment_user_clusters']
  got = segsort_common.segment_by_kmeans(cu(u['emb']), cu(u['labels']), [2, 2],
                                         cluster_indices=cu(u['cmap']), iterations=2)
  for a, k in zip(got, ('ce', 'cel', 'cl', 'ci', 'cb')):
    if a.is_floating_point():
      close(a.detach(), u[k])
    else:
      assert torch.equal(a.cpu(), u[k]), k


def test_segment_by_kmeans_more_tiles_than_resident_ctas():
  """12 images of 128 x 128: 3 072 pre-pass tiles (more than can be resident at once: the flat
  look-back relies on the ticket order) and 12 k-means clusters of 16 CTAs (two waves of the
  cluster kernel).  Ids bit-exact against the oracle."""
  g = torch.Generator().manual_seed(41)
  B, D, H, W = 12, 32, 128, 128
  centres = torch.randn(50, D, generator=g)
  region = torch.randint(0, 50, (B, H // 8, W // 8), generator=g)
  region = region.repeat_interleave(8, 1).repeat_interleave(8, 2)
  emb = (centres[region] + 0.8 * torch.randn(B, H, W, D, generator=g)).permute(0, 3, 1, 2).contiguous()
  labels = region.clone()
  labels[torch.rand(B, H, W, generator=g) < 0.07] = 255
  labels[5] = 255                                    # an image without pixels
  want = O.segment_by_kmeans(emb, labels, (6, 6), ignore_index=255, iterations=10)
  got = segsort_common.segment_by_kmeans(cu(emb), cu(labels), [6, 6], ignore_index=255, iterations=10)
  for a, b_ in zip(got, want):
    if a.is_floating_point():
      close(a.detach(), b_)
    else:
      assert torch.equal(a.cpu(), b_)


@pytest.mark.parametrize('path', ['cluster', 'small', 'fp32'])
def test_segment_by_kmeans_clusters_per_image_differ(path, monkeypatch):
  """User-supplied seed maps with a different number of clusters in every image (k_per_image):
  4 images, so the default would be the cluster kernel; every kernel against the oracle."""
  monkeypatch.setenv('SPML_B200_KMEANS', path)
  g = torch.Generator().manual_seed(13)
  B, D, H, W = 4, 16, 48, 40
  emb = torch.randn(B, D, H, W, generator=g)
  labels = torch.randint(0, 5, (B, H, W), generator=g)
  ks = (5, 9, 1, 7)
  cmap = torch.stack([(torch.arange(H * W) * k // (H * W)).view(H, W) * 3 + 2 for k in ks])   # ids with gaps
  want = O.segment_by_kmeans(emb, labels, (3, 3), cluster_indices=cmap, iterations=6)
  got = segsort_common.segment_by_kmeans(cu(emb), cu(labels), [3, 3], cluster_indices=cu(cmap),
                                         iterations=6)
  for a, b_ in zip(got, want):
    if a.is_floating_point():
      close(a.detach(), b_)
    else:
      assert torch.equal(a.cpu(), b_)


def test_gather_generic_path_equals_fast_path():
  """gather_clustering_and_update_prototypes: the re-numbering path of
  models/utils.py:95-108 and the shortcut for ids fresh from segment_by_kmeans."""
  w = synth.WORKLOADS['small']
  b = {k: v.cuda() for k, v in synth.make_batch(w).items()}
  from spml_b200.head import generate_clusters
  d = generate_clusters(b['embedding'], b['semantic_label'], b['instance_label'],
                        b['local_feature'], w.label_divisor, w.ignore_index,
                        list(w.num_clusters), w.iterations)
  args = ([d['cluster_embedding']], [d['cluster_embedding_with_loc']], [d['cluster_index']],
          [d['cluster_batch_index']], [d['cluster_semantic_label']],
          [d['cluster_instance_label']])
  fast = model_utils.gather_clustering_and_update_prototypes(*args)
  plain = d['cluster_index'].clone()          # a copy is not in the segment registry
  args = (args[0], args[1], [plain]) + args[3:]
  slow = model_utils.gather_clustering_and_update_prototypes(*args)
  for a, c in zip(fast, slow):
    assert torch.equal(a[0], c[0])
  # two "devices" (lists of two): the reference's multi-GPU gather semantics
  half = d['cluster_index'].shape[0] // 2
  two = [[t[0][:half], t[0][half:]] for t in args]
  multi = model_utils.gather_clustering_and_update_prototypes(*two)
  assert torch.equal(torch.cat(multi[5]), slow[5][0]) and torch.equal(multi[0][0], slow[0][0])


def test_abi_rejects_bad_arguments():
  lib = _lib.load()
  rc = lib.spml_segment_prototypes_fwd(None, 10, None, 8, None, 4, 1e-12, None, None, None, 0,
                                       None)
  assert rc == -1 and b'null' in lib.spml_last_error()
  x = torch.zeros(4, 200, device='cuda')
  with pytest.raises(RuntimeError, match='exceeds'):
    ops.nearest_prototype(x, x)
  with pytest.raises(RuntimeError, match='CUDA'):
    general_common.normalize_embedding(torch.zeros(3, 4))


# ------------------------------------------------------------------------------ round 2


@pytest.fixture(scope='module')
def units2():
  from conftest import load_golden
  return load_golden('units2.pt')


def test_nn_multiset_labels_golden(units2):
  """models/utils.py:157-223 against the reference's outputs (k = 1, 3; separate queries)."""
  for key in ('nn_tags_k1', 'nn_tags_k3'):
    u = units2[key]
    got = model_utils.gather_multiset_labels_per_batch_by_nearest_neighbor(
        cu(u['p']), cu(u['p']), cu(u['psem']), cu(u['pbid']), cu(u['pbid']),
        num_classes=u['num_classes'], top_k=u['top_k'], threshold=u['threshold'])
    assert torch.equal(got.cpu(), u['tags']), key
  u = units2['nn_tags_queries']
  got = model_utils.gather_multiset_labels_per_batch_by_nearest_neighbor(
      cu(u['q']), cu(u['p']), cu(u['psem']), cu(u['qbid']), cu(u['pbid']),
      num_classes=u['num_classes'], top_k=u['top_k'], threshold=u['threshold'])
  assert torch.equal(got.cpu(), u['tags'])


def test_nn_multiset_labels_random():
  g = torch.Generator().manual_seed(5)
  m, d, nc = 700, 37, 15
  base = O.l2_normalize(torch.randn(60, d, generator=g))
  p = O.l2_normalize(base[torch.randint(0, 60, (m,), generator=g)] + 0.05 * torch.randn(m, d, generator=g))
  pbid = torch.sort(torch.randint(0, 4, (m,), generator=g))[0]
  psem = torch.randint(0, nc + 3, (m,), generator=g)
  want = O.nn_multiset_labels(p, p, psem, pbid, pbid, nc, 1, 0.95)
  got = model_utils.gather_multiset_labels_per_batch_by_nearest_neighbor(
      cu(p), cu(p), cu(psem), cu(pbid), cu(pbid), num_classes=nc, top_k=1, threshold=0.95)
  assert float((got.cpu() != want).float().mean()) < 1e-3     # similarities within 1 ulp of 0.95


def test_predictions_golden(units2):
  """Segsort.predictions (segsort.py:68-125): top-20 retrieval + majority vote."""
  from spml_b200 import predictions
  u = units2['predictions']
  model = predictions.segsort(synth.make_config(synth.WORKLOADS['tiny']))
  pred, topk = model.predictions(
      {'cluster_embedding': cu(u['emb']), 'cluster_index': cu(u['cid'])},
      {'semantic_memory_prototype': cu(u['bank']),
       'semantic_memory_prototype_label': cu(u['bank_label'])})
  assert torch.equal(topk.cpu(), u['topk'])
  assert torch.equal(pred.cpu(), u['pred'])
  out = model({'cluster_embedding': cu(u['emb']), 'cluster_index': cu(u['cid'])},
              {'semantic_memory_prototype': cu(u['bank']),
               'semantic_memory_prototype_label': cu(u['bank_label'])},
              with_loss=False, with_prediction=True)
  assert torch.equal(out['semantic_prediction'].cpu(), u['pred'])


def test_segment_mean_one_hot_gather_datas(units2):
  u = units2['segment_mean']
  close(general_common.segment_mean(cu(u['x']), cu(u['index'])), u['mean'], 1e-6)
  u = units2['one_hot']
  assert torch.equal(general_common.one_hot(cu(u['labels'])).cpu(), u['auto'])
  assert torch.equal(general_common.one_hot(cu(u['labels']), 8).cpu(), u['wide'])
  u = units2['gather_datas']
  got = model_utils.gather_and_update_datas([cu(t) for t in u['in']], 'cuda:0')
  assert len(got) == 2
  for a, b in zip(got, u['out']):
    assert torch.equal(a.cpu(), b)


def test_gather_two_device_lists_golden(units2):
  """models/utils.py:41-131 over two-entry lists (the reference's multi-GPU semantics:
  global batch indices, one prototype set for both) against the reference's outputs."""
  u = units2['gather_two_devices']
  L = {k: [cu(t) for t in v] for k, v in u['lists'].items()}
  out = model_utils.gather_clustering_and_update_prototypes(
      L['cluster_embedding'], L['cluster_embedding_with_loc'], L['cluster_index'],
      L['cluster_batch_index'], L['cluster_semantic_label'], L['cluster_instance_label'], 'cuda:0')
  for got, key in zip(out, ('prototype', 'prototype_with_loc', 'prototype_semantic_label',
                            'prototype_instance_label', 'prototype_batch_index',
                            'cluster_index')):
    for a, b in zip(got, u['out'][key]):
      if b.is_floating_point():
        close(a.detach(), b)
      else:
        assert torch.equal(a.cpu(), b), key


@pytest.mark.parametrize('n,dim,k', [(65536, 128, 256), (65536, 128, 1024), (262144, 128, 64)])
def test_kmeans_sweep_sizes_match_oracle(n, dim, k):
  """BASELINE configs[4] sizes (the regime the tcgen05 E-step is dispatched for): ids
  bit-exact against the CPU oracle."""
  prob = synth.sweep_problem(n, dim, k)
  want = O.spherical_kmeans(prob['embedding'], prob['seed_label'], k, 10)
  got = segsort_common.kmeans_with_initial_labels(cu(prob['embedding']), cu(prob['seed_label']),
                                                  k, 10).cpu()
  assert int((got != want).sum()) == 0


def test_segsort_sweep_corner_matches_oracle():
  """SegSort forward + backward at a sweep corner (N = 65 536, M = 1 024, D = 128) against the
  fp64 oracle, row-chunked on the CPU side to bound memory."""
  n, m, dim, kappa = 65536, 1024, 128, 10.0
  g = torch.Generator().manual_seed(77)
  protos = O.l2_normalize(torch.randn(m, dim, generator=g))
  seg = torch.randint(0, m, (n,), generator=g)
  psem = torch.randint(0, 21, (m,), generator=g)
  e = O.l2_normalize(protos[seg] + 0.5 * torch.randn(n, dim, generator=g))
  pr = protos.double().requires_grad_(True)
  total = torch.zeros((), dtype=torch.float64)
  de = torch.empty(n, dim, dtype=torch.float64)
  for c0 in range(0, n, 8192):
    ec = e[c0:c0 + 8192].double().requires_grad_(True)
    part = O.segsort_loss(ec, psem[seg[c0:c0 + 8192]], seg[c0:c0 + 8192], pr, psem, kappa,
                          reduction='sum') / n
    part.backward()
    total += part.detach()
    de[c0:c0 + 8192] = ec.grad
  ec, pc = cu(e).requires_grad_(True), cu(protos).requires_grad_(True)
  got = segsort_loss.SegSortLoss(kappa)(ec, cu(psem[seg]), cu(seg), pc, cu(psem))
  got.backward()
  assert abs(float(got) - float(total)) <= 1e-4 * abs(float(total))
  assert norm_err(ec.grad.cpu(), de) < 1e-3
  assert norm_err(pc.grad.cpu(), pr.grad) < 1e-3


@pytest.mark.parametrize('path', ['fp32', 'tc'])
@pytest.mark.parametrize('masked', [False, True])
@pytest.mark.parametrize('n,m,rows,dim', [(3000, 900, 300, 64), (5000, 260, 130, 66), (700, 40, 3, 32)])
def test_segsort_prototype_gradient_rows(path, masked, n, m, rows, dim):
  """spml_segsort_bwd_rows: the gradient of the first `rows` prototypes only (the rest of the
  bank is detached) equals the same rows of the full gradient; the rows behind stay zero."""
  g = torch.Generator().manual_seed(n + m + rows)
  protos = O.l2_normalize(torch.randn(m, dim, generator=g))
  seg = torch.randint(0, rows, (n,), generator=g)        # pixels belong to current segments
  psem = torch.randint(0, 5, (m,), generator=g)
  e = O.l2_normalize(protos[seg] + 0.5 * torch.randn(n, dim, generator=g))
  valid = None
  if masked:
    valid = torch.rand(m, generator=g) < 0.6
    valid[seg.unique()] = True
    valid = cu(valid.to(torch.uint8))

  def run(limit):
    ec, pc = cu(e).requires_grad_(True), cu(protos).requires_grad_(True)
    problem = ops.SegsortProblem(cu(psem[seg]), cu(seg), cu(psem), 10.0, _lib.MODE_CLASS,
                                 proto_valid=valid, path=path, proto_grad_rows=limit)
    ops.SegsortLossFn.apply(ec, pc, problem).backward()
    return ec.grad, pc.grad

  de_full, dp_full = run(None)
  de_lim, dp_lim = run(rows)
  assert torch.equal(de_full, de_lim)
  assert float(dp_full[:rows].abs().max()) > 0
  assert norm_err(dp_lim[:rows].cpu(), dp_full[:rows].cpu()) < 1e-5
  assert float(dp_lim[rows:].abs().max()) == 0


def test_set_segsort_wide_tags_match_oracle():
  """40 tag columns: bits 32..39 must take part (the tcgen05 epilogue compares 32-bit codes,
  so wide tag sets run on the fp32 kernels)."""
  g = torch.Generator().manual_seed(13)
  n, m, dim, cols = 900, 150, 32, 40
  protos = O.l2_normalize(torch.randn(m, dim, generator=g))
  seg = torch.randint(0, m, (n,), generator=g)
  ptags = torch.zeros(m, cols, dtype=torch.long)
  ptags[torch.arange(m), torch.randint(32, cols, (m,), generator=g)] = 1   # only high bits
  tags = ptags[seg]
  e = O.l2_normalize(protos[seg] + 0.5 * torch.randn(n, dim, generator=g))
  er, pr = e.double().requires_grad_(True), protos.double().requires_grad_(True)
  want = O.set_segsort_loss(er, tags, seg, pr, ptags, 8.0)
  want.backward()
  ec, pc = cu(e).requires_grad_(True), cu(protos).requires_grad_(True)
  got = segsort_loss.SetSegSortLoss(8.0)(ec, cu(tags), cu(seg), pc, cu(ptags))
  got.backward()
  assert abs(float(got) - float(want)) <= 2e-5 * abs(float(want))
  assert norm_err(ec.grad.cpu(), er.grad) < 1e-4
  assert norm_err(pc.grad.cpu(), pr.grad) < 1e-4


def test_bf16_embeddings_are_accepted():
  """BASELINE configs[2] feeds a bf16 (autocast) backbone output: the head computes in fp32
  on the up-cast values and hands a bf16 gradient back."""
  w = synth.WORKLOADS['small']
  b = {k: v.cuda() for k, v in synth.make_batch(w).items()}
  cfg = synth.make_config(w)
  from spml_b200.head import ContrastiveHead
  head = ContrastiveHead(cfg).cuda()
  e16 = b['embedding'].bfloat16().requires_grad_(True)
  out16 = head(e16, b['semantic_label'], b['instance_label'], b['semantic_tag'], b['local_feature'])
  out16['loss'].backward()
  e32 = e16.detach().float().requires_grad_(True)
  out32 = head(e32, b['semantic_label'], b['instance_label'], b['semantic_tag'], b['local_feature'])
  out32['loss'].backward()
  assert e16.grad.dtype == torch.bfloat16
  assert torch.equal(out16['datas']['cluster_index'], out32['datas']['cluster_index'])
  assert abs(float(out16['loss']) - float(out32['loss'])) < 1e-6
  assert norm_err(e16.grad.float().cpu(), e32.grad.cpu()) < 1e-2


def test_inconsistent_labels_are_reported():
  """The shortcut of gather_clustering_and_update_prototypes for ids fresh from
  segment_by_kmeans checks on the device that the labels it is given are the ones the ids
  were made from; a violation surfaces through the status word."""
  w = synth.WORKLOADS['small']
  b = {k: v.cuda() for k, v in synth.make_batch(w).items()}
  from spml_b200.head import generate_clusters
  d = generate_clusters(b['embedding'], b['semantic_label'], b['instance_label'],
                        b['local_feature'], w.label_divisor, w.ignore_index,
                        list(w.num_clusters), w.iterations)
  ops.check_status()
  wrong = torch.arange(d['cluster_index'].shape[0], device='cuda') % 7
  model_utils.gather_clustering_and_update_prototypes(
      [d['cluster_embedding']], [d['cluster_embedding_with_loc']], [d['cluster_index']],
      [d['cluster_batch_index']], [wrong], [d['cluster_instance_label']])
  with pytest.raises(RuntimeError, match='different semantic'):
    ops.check_status()
  ops.check_status()      # the word was cleared


def test_calls_follow_the_tensor_device():
  """every call runs on the device of its tensors, not on the thread's current
  device (the reference's train.py gathers on cuda:{n-1} from a thread whose device is 0)."""
  if torch.cuda.device_count() < 2:
    pytest.skip('needs two GPUs')
  x = torch.randn(64, 16, device='cuda:1')
  with torch.cuda.device(0):
    y = general_common.normalize_embedding(x)
  assert y.device == x.device
  close(y, O.l2_normalize(x.cpu()), 1e-6)


def test_memory_bank_files_and_retrieval(tmp_path, units2):
  """the per-image .npy bank format (segsort/others.py:11-41) feeds
  Segsort.predictions."""
  from spml_b200 import predictions, segsort_others
  u = units2['predictions']
  half = u['bank'].shape[0] // 2
  segsort_others.save_memory_bank(str(tmp_path / 'img_b.npy'), u['bank'][half:], u['bank_label'][half:])
  segsort_others.save_memory_bank(str(tmp_path / 'img_a.npy'), u['bank'][:half], u['bank_label'][:half])
  bank, labels = segsort_others.load_memory_banks(str(tmp_path), device='cuda')
  assert torch.equal(bank.cpu(), u['bank']) and torch.equal(labels.cpu(), u['bank_label'])
  model = predictions.segsort(synth.make_config(synth.WORKLOADS['tiny']))
  pred, _ = model.predictions({'cluster_embedding': cu(u['emb']), 'cluster_index': cu(u['cid'])},
                              {'semantic_memory_prototype': bank,
                               'semantic_memory_prototype_label': labels})
  assert torch.equal(pred.cpu(), u['pred'])


def test_retrieval_at_inference_scale():
  """f2 shape: 262 144 pixels (one 512 x 512 image at full resolution) in 576 segments against a
  50 000-prototype bank, k = 20, vs the oracle's full argsort."""
  g = torch.Generator().manual_seed(21)
  n, segs, m, dim = 262144, 576, 50000, 64
  centres = O.l2_normalize(torch.randn(200, dim, generator=g))
  cid = torch.randint(0, segs, (n,), generator=g)
  seg_centre = torch.randint(0, 200, (segs,), generator=g)
  emb = O.l2_normalize(centres[seg_centre[cid]] + 0.4 * torch.randn(n, dim, generator=g))
  bank_c = torch.randint(0, 200, (m,), generator=g)
  bank = O.l2_normalize(centres[bank_c] + 0.4 * torch.randn(m, dim, generator=g))
  bank_lab = bank_c % 21
  want_pred, want_topk = O.segsort_predictions(
      {'cluster_embedding': emb, 'cluster_index': cid},
      {'semantic_memory_prototype': bank, 'semantic_memory_prototype_label': bank_lab})
  from spml_b200 import predictions
  model = predictions.segsort(synth.make_config(synth.WORKLOADS['tiny']))
  pred, topk = model.predictions({'cluster_embedding': cu(emb), 'cluster_index': cu(cid)},
                                 {'semantic_memory_prototype': cu(bank),
                                  'semantic_memory_prototype_label': cu(bank_lab)})
  assert float((topk.cpu() != want_topk).float().mean()) < 1e-3      # near-ties may swap
  assert float((pred.cpu() != want_pred).float().mean()) < 1e-3


@pytest.mark.parametrize('nq,m,dim,k', [(576, 50000, 64, 20), (130, 9000, 37, 5), (1000, 4097, 128, 20),
                                        (7, 5000, 64, 3), (300, 6000, 66, 24)])
def test_topk_tensor_core_equals_fma(nq, m, dim, k, monkeypatch):
  """topk_tc.cu (tcgen05 candidates + exact re-scoring) returns the indices of the FMA kernel:
  ragged tiles, K padding, two K blocks, duplicated bank rows (ties resolve to the lower index)."""
  from spml_b200 import _lib
  g = torch.Generator().manual_seed(nq + m)
  centres = O.l2_normalize(torch.randn(64, dim, generator=g))
  q = O.l2_normalize(centres[torch.randint(0, 64, (nq,), generator=g)] + 0.3 * torch.randn(nq, dim, generator=g))
  p = O.l2_normalize(centres[torch.randint(0, 64, (m,), generator=g)] + 0.3 * torch.randn(m, dim, generator=g))
  p[m // 2:m // 2 + 40] = p[3]                       # 41 identical rows
  p[m - 17:] = p[5]
  qlab = torch.randint(0, 21, (nq,), generator=g)
  plab = torch.randint(0, 21, (m,), generator=g)
  got = {}
  for path in ('fma', 'tc'):
    monkeypatch.setenv('SPML_B200_TOPK', path)
    labels = torch.zeros(nq, k, dtype=torch.int64, device='cuda')
    index = torch.zeros(nq, k, dtype=torch.int64, device='cuda')
    hits = torch.empty(2, dtype=torch.int32, device='cuda')
    qc, pc, qlc, plc = cu(q), cu(p), cu(qlab), cu(plab)   # (kept alive: the call takes raw pointers)
    ops.call('spml_topk_ranking', ops.ptr(qc), nq, ops.ptr(pc), m, dim, ops.ptr(qlc),
             ops.ptr(plc), None, None, k, ops.ptr(labels), ops.ptr(index), ops.ptr(hits),
             ops.stream_of(qc))
    torch.cuda.synchronize()
    got[path] = (labels.cpu(), index.cpu(), hits.cpu())
  assert torch.equal(got['fma'][1], got['tc'][1]), int((got['fma'][1] != got['tc'][1]).sum())
  assert torch.equal(got['fma'][0], got['tc'][0])
  want_hits = int((got['tc'][0] == qlab.view(-1, 1)).sum())
  assert [int(v) for v in got['tc'][2]] == [want_hits, nq], (got['tc'][2], want_hits)
  assert [int(v) for v in got['fma'][2]] == [want_hits, nq], (got['fma'][2], want_hits)
  # ... and the order of the reference's argsort where the scores are well separated
  sim = q.double() @ p.double().t()
  want = torch.argsort(sim, dim=1, descending=True, stable=True)[:, :k]
  assert float((want != got['tc'][1]).float().mean()) < 2e-3      # near-ties in fp32 may swap


def test_random_walk_matches_oracle():
  """SURVEY.md 8f-4 (pseudo_softmaxrw_crf.py:135-170)."""
  from spml_b200 import random_walk
  g = torch.Generator().manual_seed(4)
  embs = [torch.randn(1, 32, 24, 20, generator=g) for _ in range(2)]
  cam = torch.rand(21, 24, 20, generator=g)
  want = O.random_walk_cam(embs, cam, walk_steps=6)
  got = random_walk.random_walk([random_walk.embedding_affinity(cu(e)) for e in embs], cu(cam), 6)
  assert norm_err(got.cpu(), want) < 1e-4


def test_integration_snippet_runs():
  """The ctypes stub of INTEGRATION.md section 2, executed as written."""
  import re
  from conftest import ROOT
  text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
  code = next(b for b in re.findall(r'```python\n(.*?)```', text, flags=re.S)
              if 'spml_segment_prototypes_fwd' in b)
  ns = {}
  exec(compile(code, 'INTEGRATION.md', 'exec'), ns)
  g = torch.Generator().manual_seed(8)
  e = O.l2_normalize(torch.randn(300, 20, generator=g))
  lab = torch.randint(0, 11, (300,), generator=g)
  close(ns['calculate_prototypes_from_labels'](cu(e), cu(lab), 12),
        O.prototypes_from_labels(e, lab, 12))


def test_kmeans_cluster_path_is_taken(monkeypatch):
  """With SPML_B200_KMEANS=cluster the shipped 512 x 512 shapes run the cluster kernel (so the
  comparison above exercises it); larger maps and K > 128 fall back.  By default it runs from
  batch 4 on (where it is measured faster than the small-K kernel)."""
  from spml_b200 import _lib
  lib = _lib.load()
  monkeypatch.delenv('SPML_B200_KMEANS', raising=False)
  assert lib.spml_debug_kmeans_path(1, 16384, 66, 36) == 2
  assert lib.spml_debug_kmeans_path(4, 16384, 66, 36) == 3
  assert lib.spml_debug_kmeans_path(4, 16384, 66, 64) == 2
  monkeypatch.setenv('SPML_B200_KMEANS', 'cluster')
  assert lib.spml_debug_kmeans_path(1, 16384, 66, 36) == 3
  assert lib.spml_debug_kmeans_path(4, 16384, 66, 36) == 3
  assert lib.spml_debug_kmeans_path(2, 12000, 66, 64) == 3
  assert lib.spml_debug_kmeans_path(2, 16384, 66, 64) != 3     # (does not fit in shared memory)
  assert lib.spml_debug_kmeans_path(1, 16000, 37, 128) == 3
  assert lib.spml_debug_kmeans_path(1, 3000, 130, 40) == 3
  assert lib.spml_debug_kmeans_path(1, 37636, 66, 128) != 3
  assert lib.spml_debug_kmeans_path(1, 16384, 66, 129) != 3


@pytest.mark.parametrize('path', ['cluster', 'small', 'tc', 'fp32'])
def test_kmeans_rejects_inputs_outside_the_fixed_point_range(path, monkeypatch):
  """the segment sums are 2^-32 fixed point (|x| <= 8); anything else used to give
  silently wrong labels.  Every kernel now flags it and the wrapper raises."""
  monkeypatch.setenv('SPML_B200_KMEANS', path)
  g = torch.Generator().manual_seed(2)
  e = O.l2_normalize(torch.randn(600, 20, generator=g))
  lab0 = torch.randint(0, 6, (600,), generator=g)
  ok = segsort_common.kmeans_with_initial_labels(cu(e), cu(lab0), 6, 3)
  assert int(ok.min()) >= 0
  bad = e.clone()
  bad[17, 3] = float('nan')
  with pytest.raises(ValueError, match='finite'):
    segsort_common.kmeans_with_initial_labels(cu(bad), cu(lab0), 6, 3)
  with pytest.raises(ValueError, match='finite'):
    segsort_common.kmeans_with_initial_labels(cu(e * 100.0), cu(lab0), 6, 3)
