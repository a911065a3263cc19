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


def cuda_step(head, batch):
  """One step of the CUDA path.  Returns CPU tensors."""
  b = to_cuda(batch)
  emb = b['embedding'].clone().requires_grad_(True)
  out = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'],
             b['local_feature'])
  out['loss'].backward()
  res = {}
  for k in TENSOR_KEYS:
    src = out['datas'] if k.startswith('cluster') else out['targets']
    res[k] = src[k].detach().cpu()
  for k in SCALAR_KEYS:
    res[k] = out[k].detach().cpu()
  res['grad_embedding'] = emb.grad.detach().cpu()
  return res


def rel_err(a, b):
  a, b = float(a), float(b)
  return abs(a - b) / max(abs(b), 1e-12)


def norm_err(a, b):
  """||a - b|| / ||b|| in double."""
  a, b = a.double(), b.double()
  return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_step(ours, ref, ref64=None, tol=1e-3, what=''):
  """Integer outputs bit-exact; floats within `tol` of the reference (north_star:
  1e-3 relative fp32).  Where the reference's own fp32 rounding is the larger error
  (the `same - self` cancellation), being at least as close to the
  fp64 run of the same algorithm as the reference is also accepted."""
  msgs = []
  for k in ('cluster_semantic_label', 'cluster_instance_label', 'cluster_index',
            'cluster_batch_index', 'prototype_semantic_label', 'prototype_instance_label',
            'prototype_batch_index', 'prototype_semantic_tag'):
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
      gerr = 0.0
  if gerr > tol:
    msgs.append('%s grad_embedding: ||d|| / ||ref|| = %.3g' % (what, gerr))
  return msgs
