"""The tcgen05 SegSort kernels against the fp32 CUDA-core kernels and the fp64 oracle."""

import pytest
import torch

from helpers import norm_err
from oracle import spml_oracle as O
from spml_b200 import _lib, ops

pytestmark = pytest.mark.gpu


def make_problem(n, m, dim, classes, seed, kappa):
  g = torch.Generator().manual_seed(seed)
  protos = O.l2_normalize(torch.randn(m, dim, generator=g))
  seg = torch.randint(0, m, (n,), generator=g)
  psem = torch.randint(0, classes, (m,), generator=g)
  e = O.l2_normalize(protos[seg] + 0.5 * torch.randn(n, dim, generator=g))
  return e, psem[seg], seg, protos, psem


@pytest.mark.parametrize('n,m,dim,kappa', [(128, 128, 64, 6.0), (1000, 200, 64, 12.0),
                                           (777, 65, 66, 16.0), (5000, 1500, 128, 12.0),
                                           (300, 40, 37, 10.0), (2049, 513, 130, 8.0),
                                           (4096, 64, 32, 6.0)])
def test_forward_tc_matches_fp32_and_oracle(n, m, dim, kappa):
  e, sem, seg, protos, psem = make_problem(n, m, dim, 5, n + m, kappa)
  want = float(O.segsort_loss(e.double(), sem, seg, protos.double(), psem, kappa))
  ec, pc = e.cuda(), protos.cuda()
  out = {}
  for path in ('fp32', 'tc'):
    prob = ops.SegsortProblem(sem.cuda(), seg.cuda(), psem.cuda(), kappa, _lib.MODE_CLASS,
                              path=path)
    out[path] = float(ops.SegsortLossFn.apply(ec, pc, prob))
  assert abs(out['fp32'] - want) <= 2e-5 * abs(want), (out, want)
  assert abs(out['tc'] - want) <= 1e-4 * abs(want), (out, want)


def test_forward_tc_tags_and_column_filter():
  n, m, dim = 3000, 700, 64
  g = torch.Generator().manual_seed(5)
  e, _, seg, protos, _ = make_problem(n, m, dim, 5, 17, 8.0)
  img = torch.randint(0, 6, (m,), generator=g)
  img_tags = (torch.rand(6, 20, generator=g) < 0.2).long()
  ptags = img_tags[img]
  tags = ptags[seg]
  want = float(O.set_segsort_loss(e.double(), tags, seg, protos.double(), ptags, 8.0))
  pm, cm = ops.pack_tags(tags.cuda()), ops.pack_tags(ptags.cuda())
  for path in ('fp32', 'tc'):
    prob = ops.SegsortProblem(pm, seg.cuda(), cm, 8.0, _lib.MODE_TAGS, path=path)
    got = float(ops.SegsortLossFn.apply(e.cuda(), protos.cuda(), prob))
    assert abs(got - want) <= 1e-4 * abs(want), (path, got, want)
  # sem_ann-style filters: a row subset and a column mask, against the oracle on copies
  psem = torch.randint(0, 8, (m,), generator=g)
  sem = psem[seg]
  keep_r = (sem < 5).nonzero().view(-1)
  keep_c = (psem < 5).nonzero().view(-1)
  remap = torch.full((m,), -1, dtype=torch.long)
  remap[keep_c] = torch.arange(keep_c.numel())
  want = float(O.segsort_loss(e[keep_r].double(), sem[keep_r], remap[seg[keep_r]],
                              protos[keep_c].double(), psem[keep_c], 6.0))
  keep = (sem < 5).long().view(1, n).cuda()
  _, rows, off = ops.valid_scan(keep, 0, 1, n, want_src=True)
  for path in ('fp32', 'tc'):
    prob = ops.SegsortProblem(sem.cuda(), seg.cuda(), psem.cuda(), 6.0, _lib.MODE_CLASS,
                              row_index=rows, group_off=off, num_groups=1, n_rows=n,
                              max_rows_per_group=n, proto_valid=(psem < 5).cuda(), path=path)
    got = float(ops.SegsortLossFn.apply(e.cuda(), protos.cuda(), prob))
    assert abs(got - want) <= 1e-4 * abs(want), (path, got, want)


@pytest.mark.parametrize('n,m,dim,kappa', [(128, 128, 64, 6.0), (1000, 200, 64, 12.0),
                                           (777, 65, 66, 16.0), (5000, 1500, 128, 12.0),
                                           (300, 40, 37, 10.0), (2049, 513, 130, 8.0),
                                           (4096, 64, 32, 6.0), (9000, 700, 64, 6.0)])
def test_backward_tc_matches_oracle(n, m, dim, kappa):
  e, sem, seg, protos, psem = make_problem(n, m, dim, 5, n + m + 1, kappa)
  er, pr = e.double().requires_grad_(True), protos.double().requires_grad_(True)
  O.segsort_loss(er, sem, seg, pr, psem, kappa).backward()
  for path in ('fp32', 'tc'):
    ec, pc = e.cuda().requires_grad_(True), protos.cuda().requires_grad_(True)
    prob = ops.SegsortProblem(sem.cuda(), seg.cuda(), psem.cuda(), kappa, _lib.MODE_CLASS,
                              path=path)
    (ops.SegsortLossFn.apply(ec, pc, prob) * 1.7).backward()
    de, dp = norm_err(ec.grad.cpu() / 1.7, er.grad), norm_err(pc.grad.cpu() / 1.7, pr.grad)
    assert de < 2e-4 and dp < 2e-4, (path, de, dp)


def test_backward_tc_filters_and_groups():
  """sem_ann-style row list + column mask, and img_sim-style per-image groups."""
  n, m, dim = 3000, 700, 66
  g = torch.Generator().manual_seed(21)
  e, _, seg, protos, _ = make_problem(n, m, dim, 5, 23, 6.0)
  psem = torch.randint(0, 8, (m,), generator=g)
  sem = psem[seg]
  keep_r = (sem < 5).nonzero().view(-1)
  keep_c = (psem < 5).nonzero().view(-1)
  remap = torch.full((m,), -1, dtype=torch.long)
  remap[keep_c] = torch.arange(keep_c.numel())
  er, pr = e.double().requires_grad_(True), protos.double().requires_grad_(True)
  O.segsort_loss(er[keep_r], sem[keep_r], remap[seg[keep_r]], pr[keep_c], psem[keep_c],
                 6.0).backward()
  keep = (sem < 5).long().view(1, n).cuda()
  _, rows, off = ops.valid_scan(keep, 0, 1, n, want_src=True)
  for path in ('fp32', 'tc'):
    ec, pc = e.cuda().requires_grad_(True), protos.cuda().requires_grad_(True)
    prob = ops.SegsortProblem(sem.cuda(), seg.cuda(), psem.cuda(), 6.0, _lib.MODE_CLASS,
                              row_index=rows, group_off=off, num_groups=1, n_rows=n,
                              max_rows_per_group=n, proto_valid=(psem < 5).cuda(), path=path)
    ops.SegsortLossFn.apply(ec, pc, prob).backward()
    assert norm_err(ec.grad.cpu(), er.grad) < 2e-4, path
    assert norm_err(pc.grad.cpu(), pr.grad) < 2e-4, path
  # groups: 3 "images" with their own prototype ranges, mean of per-image means
  sizes, cols = [1100, 0, 1900], [300, 0, 400]
  seg2 = torch.cat([torch.randint(0, 300, (1100,), generator=g),
                    300 + torch.randint(0, 400, (1900,), generator=g)])
  inst = torch.randint(0, 30, (m,), generator=g)
  e2 = O.l2_normalize(protos[seg2] + 0.5 * torch.randn(n, dim, generator=g))
  er, pr = e2.double().requires_grad_(True), protos.double().requires_grad_(True)
  l0 = O.segsort_loss(er[:1100], inst[seg2[:1100]], seg2[:1100], pr[:300], inst[:300], 16.0)
  l1 = O.segsort_loss(er[1100:], inst[seg2[1100:]], seg2[1100:] - 300, pr[300:], inst[300:], 16.0)
  want = (l0 + l1) / 2
  want.backward()
  goff = torch.tensor([0, 1100, 1100, 3000], dtype=torch.int32).cuda()
  coff = torch.tensor([0, 300, 300, 700], dtype=torch.int32).cuda()
  for path in ('fp32', 'tc'):
    ec, pc = e2.cuda().requires_grad_(True), protos.cuda().requires_grad_(True)
    prob = ops.SegsortProblem(inst[seg2].cuda(), seg2.cuda(), inst.cuda(), 16.0, _lib.MODE_CLASS,
                              reduction=_lib.REDUCE_GROUP_MEAN, group_off=goff, col_off=coff,
                              num_groups=3, n_rows=n, max_rows_per_group=n, path=path)
    got = ops.SegsortLossFn.apply(ec, pc, prob)
    got.backward()
    assert abs(float(got) - float(want)) <= 1e-4 * abs(float(want)), (path, float(got), float(want))
    assert norm_err(ec.grad.cpu(), er.grad) < 2e-4, path
    assert norm_err(pc.grad.cpu(), pr.grad) < 2e-4, path
