"""Which operand precision does the similarity GEMM of the SegSort losses need?  (CPU only.)

The tcgen05 kernels split the fp32 unit vectors into bf16 hi + lo and issue three products per
similarity (hi.hi + lo.hi + hi.lo).  VERDICT r1 asked for a measurement of cheaper schemes
against the parity bar (loss and d loss / d embedding within 1e-3 relative of the reference).
This script takes the clustered pixels and prototypes of a workload from the fp64 oracle, replaces
the similarity matrix S = E P^T by what each scheme would deliver (operands rounded to the
scheme's input type, products and sums exact: the tensor core accumulates in fp32), and reports
the relative error of the loss and of the gradient (everything downstream of S in fp64, the
gradient through S taken as the exact GEMM, which isolates the effect of the similarity itself).

    python scripts/precision_schemes.py [workload ...]   > profiles/r2b_precision_schemes.md
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spml_oracle as O  # noqa: E402  (a study script, not the product)
from spml_b200 import synth  # noqa: E402


def rnd(x, kind):
  if kind == 'bf16':
    return x.to(torch.float32).to(torch.bfloat16).to(torch.float64)
  if kind == 'fp16':
    return x.to(torch.float32).to(torch.float16).to(torch.float64)
  if kind == 'tf32':   # 10 explicit mantissa bits, round to nearest even
    i = x.to(torch.float32).view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32).to(torch.float64)
  raise KeyError(kind)


def similarity(e, p, scheme):
  if scheme == 'exact':
    return e @ p.t()
  kind, n = scheme.split('x')
  eh, ph = rnd(e, kind), rnd(p, kind)
  if n == '1':
    return eh @ ph.t()
  el = rnd(e - eh, kind)
  if n == '2':                      # pixel rows split, prototype rows rounded
    return (eh + el) @ ph.t()
  pl = rnd(p - ph, kind)
  return eh @ ph.t() + el @ ph.t() + eh @ pl.t()     # n == '3': what the kernels issue


def loss_and_grad(e, sem, seg, p, psem, kappa, scheme):
  e = e.clone().requires_grad_(True)
  exact = e @ p.t()
  with torch.no_grad():
    delta = similarity(e.detach(), p, scheme) - exact
  sim = ((exact + delta) * kappa).exp()
  self_sim = torch.gather(sim, 1, seg.view(-1, 1))
  same = torch.eq(sem.view(-1, 1), psem.view(1, -1)).to(sim.dtype)
  nll = O._nll_from_similarity(sim, self_sim, same, 1.0 - same)
  loss = nll.mean()
  loss.backward()
  return loss.detach(), e.grad.detach()


SCHEMES = ('bf16x1', 'bf16x2', 'bf16x3', 'fp16x1', 'fp16x2', 'fp16x3', 'tf32x1')
print('# Operand precision of the similarity GEMM against the 1e-3 parity bar\n')
print('`python scripts/precision_schemes.py` (CPU, fp64 everywhere except the rounding of the GEMM '
      'operands). xN = products issued per similarity: x1 both operands rounded, x2 pixel rows '
      'split hi + lo against rounded prototypes, x3 hi.hi + lo.hi + hi.lo (shipped, with bf16). '
      'Entries: relative error of the loss / relative L2 error of d loss / d embedding; '
      '**bold** = misses 1e-3.\n')
for name in (sys.argv[1:] or ['voc_scribble_b1', 'voc_tag_b2', 'small']):
  w = synth.WORKLOADS[name]
  cfg = synth.make_config(w)
  out = O.contrastive_step(cfg, synth.make_batch(w), dtype=torch.float64)
  e = out['cluster_embedding'].double()
  sem, seg = out['cluster_semantic_label'], out['cluster_index']
  p, psem = out['prototype'].double(), out['prototype_semantic_label']
  print('## %s: %d pixels x %d prototypes, D = %d\n' % (name, e.shape[0], p.shape[0], e.shape[1]))
  print('| kappa | ' + ' | '.join(SCHEMES) + ' |')
  print('|---|' + '---|' * len(SCHEMES))
  for kappa in (6.0, 12.0, 16.0):
    l0, g0 = loss_and_grad(e, sem, seg, p, psem, kappa, 'exact')
    cells = []
    for s in SCHEMES:
      l, g = loss_and_grad(e, sem, seg, p, psem, kappa, s)
      le = float((l - l0).abs() / l0.abs())
      ge = float((g - g0).norm() / g0.norm())
      cell = '%.1e / %.1e' % (le, ge)
      cells.append('**%s**' % cell if max(le, ge) > 1e-3 else cell)
    print('| %g | ' % kappa + ' | '.join(cells) + ' |')
  print()
