"""Shared helpers of the parity tests: run the CUDA path on a synthetic batch and
collect the same keys as oracle.spml_oracle.contrastive_step."""

import torch

from spml_b200 import synth
from spml_b200.head import ContrastiveHead

TENSOR_KEYS = ('cluster_embedding', 'cluster_embedding_with_loc', 'cluster_semantic_label',
               'cluster_instance_label', 'cluster_index', 'cluster_batch_index', 'prototype',
               'prototype_with_loc', 'prototype_semantic_label', 'prototype_instance_label',
               'prototype_batch_index', 'prototype_semantic_tag')
SCALAR_KEYS = ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss', 'accuracy', 'loss')


def to_cuda(batch):
  return {k: v.cuda() for k, v in batch.items()}


def make_head(cfg, w):
  """ContrastiveHead of the workload's variant on the GPU; the softmax variants get the
  oracle's seeded classifier weights and run in eval mode (like the golden runs)."""
  head = ContrastiveHead(cfg, variant=w.variant).cuda()
  if w.variant != 'segsort':
    from oracle import spml_oracle as O
    head.predictor.semantic_classifier.load_state_dict(O.make_classifier(cfg).state_dict())
    head.eval()
  return head


def cuda_step(head, batch):
  """One step of the CUDA path.  Returns CPU tensors."""
  b = to_cuda(batch)
  emb = b['embedding'].clone().requires_grad_(True)
  head.zero_grad()
  out = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'],
             b['local_feature'], b.get('semantic_label_full'))
  out['loss'].backward()
  res = {}
  for k in TENSOR_KEYS:
    src = out['datas'] if k.startswith('cluster') else out['targets']
    if k in src:
      res[k] = src[k].detach().cpu()
  zero = torch.zeros(())
  for k in SCALAR_KEYS:
    res[k] = out[k].detach().cpu() if out.get(k) is not None else zero
  res['grad_embedding'] = emb.grad.detach().cpu()
  if head.variant != 'segsort':
    res['grad_classifier'] = {k: v.grad.detach().cpu() for k, v in
                              head.predictor.semantic_classifier.named_parameters()}
  return res


def rel_err(a, b):
  a, b = float(a), float(b)
  return abs(a - b) / max(abs(b), 1e-12)


def norm_err(a, b):
  """||a - b|| / ||b|| in double."""
  a, b = a.double(), b.double()
  return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_step(ours, ref, ref64=None, tol=1e-3, what='', escapes=None):
  """Integer outputs bit-exact; floats within `tol` of the reference (north_star:
  1e-3 relative fp32).  Where the reference's own fp32 rounding is the larger error
  (the `same - self` cancellation), being at least as close to the
  fp64 run of the same algorithm as the reference is also accepted; every key that needed
  that second criterion is appended to `escapes` (callers assert on the list)."""
  msgs = []
  escapes = escapes if escapes is not None else []
  for k in ('cluster_semantic_label', 'cluster_instance_label', 'cluster_index',
            'cluster_batch_index', 'prototype_semantic_label', 'prototype_instance_label',
            'prototype_batch_index', 'prototype_semantic_tag'):
    if k not in ref:
      continue
    if ours[k].shape != ref[k].shape:
      msgs.append('%s %s: shape %s vs %s' % (what, k, tuple(ours[k].shape), tuple(ref[k].shape)))
    elif not torch.equal(ours[k], ref[k]):
      msgs.append('%s %s: %d of %d entries differ' % (what, k, int((ours[k] != ref[k]).sum()),
                                                      ref[k].numel()))
  if msgs:
    return msgs
  for k in ('cluster_embedding', 'cluster_embedding_with_loc', 'prototype', 'prototype_with_loc'):
    err = float((ours[k] - ref[k]).abs().max())
    if err > 2e-6:
      msgs.append('%s %s: max abs err %.3g' % (what, k, err))
  for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss', 'loss'):
    e32 = abs(float(ours[k]) - float(ref[k]))
    ok = e32 <= tol * abs(float(ref[k])) + 1e-7
    if not ok and ref64 is not None:
      e64 = abs(float(ours[k]) - float(ref64[k]))
      ok = e64 <= abs(float(ref[k]) - float(ref64[k])) + 1e-5 * abs(float(ref64[k]))
      if ok:
        escapes.append('%s %s' % (what, k))
    if not ok:
      msgs.append('%s %s: ours %.8g ref %.8g%s' % (
          what, k, float(ours[k]), float(ref[k]),
          '' if ref64 is None else ' ref64 %.8g' % float(ref64[k])))
  if abs(float(ours['accuracy']) - float(ref['accuracy'])) > 1e-6:
    msgs.append('%s accuracy: ours %.6f ref %.6f' % (what, float(ours['accuracy']),
                                                     float(ref['accuracy'])))
  gerr = norm_err(ours['grad_embedding'], ref['grad_embedding'])
  if gerr > tol and ref64 is not None:
    g64 = ref64['grad_embedding'].float()
    if norm_err(ours['grad_embedding'], g64) <= norm_err(ref['grad_embedding'], g64) + 1e-5:
      escapes.append('%s grad_embedding (%.3g)' % (what, gerr))
      gerr = 0.0
  if gerr > tol:
    msgs.append('%s grad_embedding: ||d|| / ||ref|| = %.3g' % (what, gerr))
  if 'grad_classifier' in ref:
    for name, g in ref['grad_classifier'].items():
      err = norm_err(ours['grad_classifier'][name], g)
      if err > tol:
        msgs.append('%s d(classifier.%s): ||d|| / ||ref|| = %.3g' % (what, name, err))
  return msgs
