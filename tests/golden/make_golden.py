"""Generates tests/golden/*.pt by running the REFERENCE itself (twke18/SPML) on CPU.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

The reference is imported from where it lies, never copied.  Three shims make
it run on CPU; none changes the arithmetic:
  1. spml/utils/segsort/common.py:376 reads `tensor.device.index`, which is None
     on the CPU -> the module source is patched IN MEMORY to `... .index or 0`.
  2. spml/models/utils.py:86-92 gathers with torch.nn.parallel.scatter_gather,
     which asserts on CPU tensors -> replaced by torch.cat (same semantics).
  3. spml.config needs easydict (absent) -> a SimpleNamespace tree is passed.

Each golden file is a dict {'inputs': ..., 'outputs': ..., 'meta': ...} of small
CPU tensors, saved with torch.save.
"""

import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

from spml_b200 import synth  # noqa: E402


def load_reference():
  sys.path.insert(0, REF)
  import spml.utils.general.common  # noqa: F401
  import spml.utils.segsort as pkg
  path = os.path.join(REF, 'spml/utils/segsort/common.py')
  src = open(path).read()
  needle = 'gpu_id = cur_cluster_indices.device.index\n'
  assert src.count(needle) == 1
  src = src.replace(needle, 'gpu_id = cur_cluster_indices.device.index or 0\n')
  mod = types.ModuleType('spml.utils.segsort.common')
  mod.__file__ = path
  exec(compile(src, path, 'exec'), mod.__dict__)
  sys.modules['spml.utils.segsort.common'] = mod
  pkg.common = mod
  import spml.models.utils as MU
  MU.scatter_gather.gather = lambda xs, dev, dim=0: torch.cat(list(xs), dim)
  import spml.utils.segsort.loss as L
  import spml.utils.segsort.eval as E
  import spml.utils.general.common as G
  import spml.models.predictions.segsort as P
  import spml.models.predictions.segsort_softmax as PS
  import spml.models.predictions.segsort_softmax_densepose as PD
  import spml.models.embeddings.resnet_deeplab as ED
  import spml.models.embeddings.resnet_pspnet_densepose as EP
  return types.SimpleNamespace(common=mod, loss=L, eval=E, general=G, model_utils=MU,
                               segsort=P, segsort_softmax=PS, segsort_softmax_densepose=PD,
                               deeplab=ED.ResnetDeeplab, pspnet_densepose=EP.ResnetPspnet)


def reference_model(ref, cfg, variant, classifier_state=None):
  """The reference's prediction model for a workload variant; the softmax variants get
  the seeded classifier weights and run in eval mode (dropout p = 0.75 draws from the
  global RNG in train mode, which no other device reproduces)."""
  if variant == 'segsort':
    return ref.segsort.segsort(cfg)
  mod = ref.segsort_softmax_densepose if variant == 'densepose' else ref.segsort_softmax
  model = mod.segsort(cfg)
  model.semantic_classifier.load_state_dict(classifier_state)
  return model.eval()


def reference_step(ref, cfg, batch, bank, variant='segsort', classifier_state=None):
  """Appendix A of SURVEY.md: A9 (the reference's own generate_clusters method, called on
  a stand-in `self` that carries the four attributes it reads) -> B1 -> B2 ->
  prediction model forward -> backward."""
  emb = batch['embedding'].clone().requires_grad_(True)
  sem, inst = batch['semantic_label'], batch['instance_label']
  me = types.SimpleNamespace(label_divisor=cfg.network.label_divisor,
                             semantic_ignore_index=cfg.dataset.semantic_ignore_index,
                             kmeans_num_clusters=cfg.network.kmeans_num_clusters,
                             kmeans_iterations=cfg.network.kmeans_iterations)
  net = ref.pspnet_densepose if variant == 'densepose' else ref.deeplab
  cl = net.generate_clusters(me, emb, sem, inst, batch['local_feature'])
  ce, cel = cl['cluster_embedding'], cl['cluster_embedding_with_loc']
  csl, cil = cl['cluster_semantic_label'], cl['cluster_instance_label']
  ci, cb = cl['cluster_index'], cl['cluster_batch_index']
  p, pl, psl, pil, pbi, ci2 = ref.model_utils.gather_clustering_and_update_prototypes(
      [ce], [cel], [ci], [cb], [csl], [cil], 'cpu')
  datas = dict(cluster_index=ci2[0], cluster_embedding=ce, cluster_embedding_with_loc=cel,
               cluster_semantic_label=csl, cluster_instance_label=cil,
               cluster_batch_index=cb)
  tags = batch['semantic_tag']
  targets = dict(prototype=p[0], prototype_with_loc=pl[0], prototype_semantic_label=psl[0],
                 prototype_instance_label=pil[0], prototype_batch_index=pbi[0])
  if variant != 'densepose':      # train_densepose.py:189-199 has the tag gather commented out
    targets.update(semantic_tag=tags,
                   prototype_semantic_tag=tags.index_select(0, pbi[0]))   # train.py:199-202
  targets.update({k: list(v) for k, v in bank.items()})
  model = reference_model(ref, cfg, variant, classifier_state)
  if variant != 'segsort':
    datas['embedding'] = emb
    targets['semantic_label'] = batch.get('semantic_label_full', sem).clone()
  out = model(datas, targets)
  losses = [out[k] for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss')
            if out[k] is not None]
  total = sum(losses)                                                    # train.py:213-219
  total.backward()
  datas.pop('embedding', None)
  targets.pop('semantic_label', None)
  res = {k: v.detach().clone() for k, v in datas.items()}
  res['cluster_index_before_gather'] = ci.detach().clone()
  res.update({k: v.detach().clone() for k, v in targets.items() if torch.is_tensor(v)})
  for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss', 'accuracy'):
    res[k] = out[k].detach().clone() if out[k] is not None else torch.zeros(())
  res['loss'] = total.detach().clone()
  res['grad_embedding'] = emb.grad.detach().clone()
  if variant != 'segsort':
    res['grad_classifier'] = {k: v.grad.detach().clone()
                              for k, v in model.semantic_classifier.named_parameters()}
  return res, targets


def bank_update(bank, targets, size, stride):
  """pyscripts/train/train.py:276-293, executed literally."""
  with torch.no_grad():
    for k in targets.keys():
      if 'prototype' in k and 'memory' not in k:
        memory = targets[k].clone().detach()
        key = 'memory_' + k
        bank.setdefault(key, []).append(memory)
        if len(bank[key]) > size:
          bank[key] = bank[key][1:]
    for t in bank.get('memory_prototype_batch_index', []):
      t += stride


def golden_steps(ref, name, steps, seed=235):
  w = synth.WORKLOADS[name]
  cfg = synth.make_config(w)
  bank = {}
  files = []
  state = None
  if w.variant != 'segsort':
    from oracle import spml_oracle as O      # only for the seeded classifier WEIGHTS (inputs)
    state = O.make_classifier(cfg).state_dict()
  for s in range(steps):
    batch = synth.make_batch(w, seed=seed, step=s)
    bank_in = {k: [t.clone() for t in v] for k, v in bank.items()}
    res, targets = reference_step(ref, cfg, batch, bank, w.variant, state)
    fn = os.path.join(HERE, '%s_step%d.pt' % (name, s))
    torch.save({'inputs': batch, 'bank': bank_in, 'outputs': res,
                'meta': {'workload': name, 'seed': seed, 'step': s,
                         'torch': torch.__version__}}, fn)
    files.append(fn)
    bank_update(bank, targets, w.memory_bank_size, w.batch)
  return files


def golden_units(ref):
  """Edge cases of the individual functions (sizes of a few dozen elements)."""
  g = torch.Generator().manual_seed(235)
  C, L, E, G = ref.common, ref.loss, ref.eval, ref.general
  u = {}
  # A1: zero vector stays zero (0 / 1e-12), tiny vector below eps is scaled by 1/eps.
  x = torch.randn(6, 5, generator=g)
  x[2] = 0
  x[4] = 1e-14
  u['normalize'] = {'x': x, 'y': G.normalize_embedding(x)}
  # A2 / A3
  u['location_float'] = C.generate_location_features((5, 7), 'cpu', 'float')
  u['location_int'] = C.generate_location_features((5, 7), 'cpu', 'int')
  for nc, hw in (((2, 3), (4, 5)), ((6, 6), (128, 128)), ((8, 16), (194, 194)),
                 ((3, 3), (24, 20)), ((12, 12), (97, 97))):
    u['seeds_%dx%d_%dx%d' % (nc + hw)] = C.initialize_cluster_labels(list(nc), hw, 'cpu')
  # A4 with an empty label (zero prototype) / A5 ties pick the first index.
  e = G.normalize_embedding(torch.randn(40, 8, generator=g))
  lab = torch.randint(0, 5, (40,), generator=g)
  lab[lab == 3] = 2
  u['prototypes'] = {'e': e, 'lab': lab, 'p': C.calculate_prototypes_from_labels(e, lab, 6),
                     'p_auto': C.calculate_prototypes_from_labels(e, lab)}
  p = torch.zeros(4, 8)
  p[1, 0] = 1
  p[2, 0] = 1
  e2 = torch.zeros(3, 8)
  e2[0, 0] = 1       # ties between prototypes 1 and 2 -> 1
  e2[1, 0] = -1      # all-negative except zero prototypes 0 and 3 -> 0
  u['nearest_ties'] = {'e': e2, 'p': p, 'idx': C.find_nearest_prototypes(e2, p)}
  # A6: k-means where one seed cluster is empty from the start.
  lab0 = torch.randint(0, 4, (40,), generator=g)
  lab0[lab0 == 1] = 0
  u['kmeans_empty'] = {'e': e, 'lab0': lab0,
                       'lab': C.kmeans_with_initial_labels(e, lab0, 4, 10)}
  # A7
  sem = torch.randint(0, 4, (30,), generator=g)
  ins = torch.randint(0, 6, (30,), generator=g)
  pl, inv = C.prepare_prototype_labels(sem, ins, 256)
  u['prototype_labels'] = {'sem': sem, 'inst': ins, 'plab': pl, 'inv': inv}
  # C1: a class with a single segment takes the fall-back-to-self branch.
  N, M, D = 24, 6, 8
  pe = G.normalize_embedding(torch.randn(N, D, generator=g)).requires_grad_(True)
  pp = G.normalize_embedding(torch.randn(M, D, generator=g)).requires_grad_(True)
  seg = torch.randint(0, M, (N,), generator=g)
  psem = torch.tensor([0, 0, 1, 2, 2, 3])
  psem_pix = psem[seg]
  loss = L.SegSortLoss(10.0)(pe, psem_pix, seg, pp, psem)
  loss.backward()
  nll = L._calculate_log_likelihood(pe.detach(), psem_pix, seg, pp.detach(), psem, 10.0,
                                    'segsort+')
  u['segsort'] = {'e': pe.detach(), 'sem': psem_pix, 'seg': seg, 'p': pp.detach(),
                  'psem': psem, 'kappa': 10.0, 'loss': loss.detach(), 'nll': nll,
                  'de': pe.grad.clone(), 'dp': pp.grad.clone()}
  # C2: image 2 is background-only (all-zero tag rows): own segment in num and den.
  pe2 = pe.detach().clone().requires_grad_(True)
  pp2 = pp.detach().clone().requires_grad_(True)
  img_tags = torch.tensor([[1, 0, 0, 1, 0], [0, 1, 0, 0, 0], [0, 0, 0, 0, 0]])
  pimg = torch.tensor([0, 0, 1, 1, 2, 2])
  ptags = img_tags[pimg]
  tags = ptags[seg]
  loss2 = L.SetSegSortLoss(8.0)(pe2, tags, seg, pp2, ptags)
  loss2.backward()
  u['set_segsort'] = {'e': pe2.detach(), 'tags': tags, 'seg': seg, 'p': pp2.detach(),
                      'ptags': ptags, 'kappa': 8.0, 'loss': loss2.detach(),
                      'de': pe2.grad.clone(), 'dp': pp2.grad.clone()}
  # C3
  q = G.normalize_embedding(torch.randn(12, D, generator=g))
  ql = torch.randint(0, 3, (12,), generator=g)
  bank = G.normalize_embedding(torch.randn(30, D, generator=g))
  bl = torch.randint(0, 3, (30,), generator=g)
  acc, topk = E.top_k_ranking(q, ql, bank, bl, 5)
  u['topk'] = {'q': q, 'ql': ql, 'p': bank, 'pl': bl, 'acc': acc, 'labels': topk,
               'majority': E.majority_label_from_topk(topk, 3)}
  # A8 with every pixel of image 1 ignored and a user-supplied cluster map.
  emb = torch.randn(2, 6, 8, 8, generator=g)
  labels = torch.randint(0, 3, (2, 8, 8), generator=g)
  labels[1] = 7
  ce, cel, cl, ci, cb = C.segment_by_kmeans(emb, labels, [2, 2], ignore_index=7, iterations=3)
  u['segment_ignore_image'] = {'emb': emb, 'labels': labels, 'ce': ce, 'cel': cel,
                               'cl': cl, 'ci': ci, 'cb': cb}
  cmap = torch.randint(0, 50, (2, 8, 8), generator=g) * 3
  labels2 = torch.randint(0, 3, (2, 8, 8), generator=g)
  ce, cel, cl, ci, cb = C.segment_by_kmeans(emb, labels2, [2, 2], cluster_indices=cmap,
                                            iterations=2)
  u['segment_user_clusters'] = {'emb': emb, 'labels': labels2, 'cmap': cmap, 'ce': ce,
                                'cel': cel, 'cl': cl, 'ci': ci, 'cb': cb}
  fn = os.path.join(HERE, 'units.pt')
  torch.save(u, fn)
  return [fn]


def golden_units2(ref):
  """Round-2 unit cases (their own file and RNG stream, so units.pt stays byte-stable):
  NN tag propagation, retrieval predictions, segment_mean, one_hot, list gathers."""
  g = torch.Generator().manual_seed(236)
  C, E, G, MU = ref.common, ref.eval, ref.general, ref.model_utils
  u = {}
  # f3: same-image top-k neighbours with a class label; image 2 has no labelled prototype
  # (every neighbour is masked -> no tag), threshold cuts the far ones.
  M, D, NC = 40, 7, 6
  base = G.normalize_embedding(torch.randn(8, D, generator=g))
  protos = G.normalize_embedding(base[torch.randint(0, 8, (M,), generator=g)]
                                 + 0.15 * torch.randn(M, D, generator=g))
  pbid = torch.sort(torch.randint(0, 3, (M,), generator=g))[0]
  psem = torch.randint(0, NC + 2, (M,), generator=g)
  psem[pbid == 2] = NC
  for k, thr in ((1, 0.95), (3, 0.9)):
    u['nn_tags_k%d' % k] = {
        'p': protos, 'psem': psem, 'pbid': pbid, 'num_classes': NC, 'top_k': k,
        'threshold': thr,
        'tags': MU.gather_multiset_labels_per_batch_by_nearest_neighbor(
            protos, protos, psem, pbid, pbid, num_classes=NC, top_k=k, threshold=thr)}
  q = G.normalize_embedding(protos[:25] + 0.05 * torch.randn(25, D, generator=g))
  qbid = pbid[:25].clone()
  u['nn_tags_queries'] = {
      'q': q, 'qbid': qbid, 'p': protos, 'psem': psem, 'pbid': pbid, 'num_classes': NC,
      'top_k': 2, 'threshold': 0.8,
      'tags': MU.gather_multiset_labels_per_batch_by_nearest_neighbor(
          q, protos, psem, qbid, pbid, num_classes=NC, top_k=2, threshold=0.8)}
  # f2: Segsort.predictions (top-20 retrieval + majority vote) with non-dense cluster ids
  N, D2, NB = 300, 12, 90
  cid = torch.randint(0, 23, (N,), generator=g) * 3 + 5
  centres = G.normalize_embedding(torch.randn(75, D2, generator=g))
  emb = G.normalize_embedding(centres[cid] + 0.3 * torch.randn(N, D2, generator=g))
  bank = G.normalize_embedding(centres[torch.randint(0, 75, (NB,), generator=g)]
                               + 0.3 * torch.randn(NB, D2, generator=g))
  bank_lab = torch.randint(0, 5, (NB,), generator=g)
  cfg = synth.make_config(synth.WORKLOADS['tiny'])
  pred, topk = ref.segsort.segsort(cfg).predictions(
      {'cluster_embedding': emb, 'cluster_index': cid},
      {'semantic_memory_prototype': bank, 'semantic_memory_prototype_label': bank_lab})
  u['predictions'] = {'emb': emb, 'cid': cid, 'bank': bank, 'bank_label': bank_lab,
                      'pred': pred, 'topk': topk}
  # general/common.py: segment_mean with an empty segment, one_hot
  x = torch.randn(50, 6, generator=g)
  idx = torch.randint(0, 9, (50,), generator=g)
  idx[idx == 4] = 3
  u['segment_mean'] = {'x': x, 'index': idx, 'mean': G.segment_mean(x, idx)}
  lab = torch.randint(0, 5, (4, 7), generator=g)
  u['one_hot'] = {'labels': lab, 'auto': G.one_hot(lab), 'wide': G.one_hot(lab, 8)}
  # B2: gather_and_update_datas over a two-entry list
  a, b = torch.randint(0, 2, (2, 256), generator=g), torch.randint(0, 2, (3, 256), generator=g)
  got = MU.gather_and_update_datas([a, b], 'cpu')
  u['gather_datas'] = {'in': [a, b], 'out': [t.clone() for t in got]}
  # B1 over a two-entry list (the reference's multi-GPU semantics: global batch indices)
  w = synth.WORKLOADS['tiny']
  cfgw = synth.make_config(w)
  me = types.SimpleNamespace(label_divisor=cfgw.network.label_divisor,
                             semantic_ignore_index=cfgw.dataset.semantic_ignore_index,
                             kmeans_num_clusters=cfgw.network.kmeans_num_clusters,
                             kmeans_iterations=cfgw.network.kmeans_iterations)
  lists = {k: [] for k in ('cluster_embedding', 'cluster_embedding_with_loc', 'cluster_index',
                           'cluster_batch_index', 'cluster_semantic_label',
                           'cluster_instance_label')}
  batches = []
  for dev in range(2):
    batch = synth.make_batch(w, seed=240 + dev)
    batches.append(batch)
    cl = ref.deeplab.generate_clusters(me, batch['embedding'], batch['semantic_label'],
                                       batch['instance_label'], batch['local_feature'])
    cl['cluster_batch_index'] = cl['cluster_batch_index'] + dev * w.batch   # common.py:376-377
    for k in lists:
      lists[k].append(cl[k].detach())
  out = MU.gather_clustering_and_update_prototypes(
      lists['cluster_embedding'], lists['cluster_embedding_with_loc'], lists['cluster_index'],
      lists['cluster_batch_index'], lists['cluster_semantic_label'],
      lists['cluster_instance_label'], 'cpu')
  u['gather_two_devices'] = {
      'inputs': batches, 'lists': lists,
      'out': {k: [t.detach().clone() for t in v] for k, v in zip(
          ('prototype', 'prototype_with_loc', 'prototype_semantic_label',
           'prototype_instance_label', 'prototype_batch_index', 'cluster_index'), out)}}
  fn = os.path.join(HERE, 'units2.pt')
  torch.save(u, fn)
  return [fn]


def main():
  torch.manual_seed(235)
  torch.set_num_threads(1)      # fixed summation order for the fixtures
  ref = load_reference()
  files = golden_units(ref)
  files += golden_steps(ref, 'tiny', 3)
  files += golden_steps(ref, 'small', 3)
  files += golden_units2(ref)
  files += golden_steps(ref, 'tiny_softmax', 3)
  files += golden_steps(ref, 'tiny_densepose', 2)
  files += golden_steps(ref, 'tiny_densepose_shipped', 1)
  for f in files:
    print('%8d  %s' % (os.path.getsize(f), os.path.relpath(f, ROOT)))


if __name__ == '__main__':
  main()
