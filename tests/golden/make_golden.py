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
  return types.SimpleNamespace(common=mod, loss=L, eval=E, general=G, model_utils=MU,
                               segsort=P)


def reference_step(ref, cfg, batch, bank):
  """Appendix A of SURVEY.md: A9 -> B1 -> B2 -> Segsort.forward -> backward."""
  C = ref.common
  emb = batch['embedding'].clone().requires_grad_(True)
  sem, inst = batch['semantic_label'], batch['instance_label']
  div = cfg.network.label_divisor
  labels = sem * div + inst                                   # resnet_deeplab.py:112-117
  ign = labels.max() + 1
  labels = labels.masked_fill(sem == cfg.dataset.semantic_ignore_index, ign)
  ce, cel, cl, ci, cb = C.segment_by_kmeans(
      emb, labels, cfg.network.kmeans_num_clusters,
      local_features=batch['local_feature'], ignore_index=ign,
      iterations=cfg.network.kmeans_iterations)
  csl, cil = cl // div, cl % div
  p, pl, psl, pil, pbi, ci2 = ref.model_utils.gather_clustering_and_update_prototypes(
      [ce], [cel], [ci], [cb], [csl], [cil], 'cpu')
  datas = dict(cluster_index=ci2[0], cluster_embedding=ce, cluster_embedding_with_loc=cel,
               cluster_semantic_label=csl, cluster_instance_label=cil,
               cluster_batch_index=cb)
  tags = batch['semantic_tag']
  targets = dict(prototype=p[0], prototype_with_loc=pl[0], prototype_semantic_label=psl[0],
                 prototype_instance_label=pil[0], prototype_batch_index=pbi[0],
                 semantic_tag=tags,
                 prototype_semantic_tag=tags.index_select(0, pbi[0]))   # train.py:199-202
  targets.update({k: list(v) for k, v in bank.items()})
  model = ref.segsort.segsort(cfg)
  out = model(datas, targets)
  total = out['sem_ann_loss'] + out['sem_occ_loss'] + out['img_sim_loss']
  total.backward()
  res = {k: v.detach().clone() for k, v in datas.items()}
  res['cluster_index_before_gather'] = ci.detach().clone()
  res.update({k: v.detach().clone() for k, v in targets.items() if torch.is_tensor(v)})
  res.update({k: out[k].detach().clone() for k in
              ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss', 'accuracy')})
  res['loss'] = total.detach().clone()
  res['grad_embedding'] = emb.grad.detach().clone()
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
  for s in range(steps):
    batch = synth.make_batch(w, seed=seed, step=s)
    bank_in = {k: [t.clone() for t in v] for k, v in bank.items()}
    res, targets = reference_step(ref, cfg, batch, bank)
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


def main():
  torch.manual_seed(235)
  torch.set_num_threads(1)      # fixed summation order for the fixtures
  ref = load_reference()
  files = golden_units(ref)
  files += golden_steps(ref, 'tiny', 3)
  files += golden_steps(ref, 'small', 3)
  for f in files:
    print('%8d  %s' % (os.path.getsize(f), os.path.relpath(f, ROOT)))


if __name__ == '__main__':
  main()
