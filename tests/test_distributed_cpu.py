"""World-size-2 gloo tests of the rank plumbing (CPU, no kernels)."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spml_b200 import distributed as D


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, size, port, out):
  os.environ.update(RANK=str(rank), WORLD_SIZE=str(size), LOCAL_RANK=str(rank),
                    MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  r, s, _ = D.init('gloo')
  assert (r, s) == (rank, size)
  lo, hi = D.shard_images(7, rank, size)
  mx = D.max_over_ranks(10.0 + rank)
  total = D.sum_over_ranks(hi - lo)
  w1 = torch.nn.Parameter(torch.zeros(3, 5))
  w2 = torch.nn.Parameter(torch.zeros(11))
  w1.grad = torch.full((3, 5), float(rank + 1))
  w2.grad = torch.arange(11, dtype=torch.float32) * (rank + 1)
  buckets = D.all_reduce_gradients([w1, w2], bucket_bytes=32)
  ok = (torch.allclose(w1.grad, torch.full((3, 5), 1.5)) and
        torch.allclose(w2.grad, torch.arange(11, dtype=torch.float32) * 1.5))
  out[rank] = (lo, hi, mx, total, buckets, ok, D.batch_index_offset(4, rank))
  dist.destroy_process_group()


def test_two_rank_plumbing():
  size, port = 2, _free_port()
  with mp.Manager() as m:
    out = m.dict()
    mp.spawn(_worker, args=(size, port, out), nprocs=size, join=True)
    r0, r1 = out[0], out[1]
  assert (r0[0], r0[1], r1[0], r1[1]) == (0, 4, 4, 7)        # shards tile the batch
  assert r0[2] == r1[2] == 11.0 and r0[3] == r1[3] == 7.0
  assert r0[4] == r1[4] == 2 and r0[5] and r1[5]
  assert (r0[6], r1[6]) == (0, 4)


def test_shards_cover_any_batch():
  for n in range(0, 40):
    for w in (1, 2, 3, 4, 8):
      spans = [D.shard_images(n, r, w) for r in range(w)]
      assert spans[0][0] == 0 and spans[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
      assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


def test_single_process_is_identity():
  assert D.world() == (int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)),
                       int(os.environ.get('LOCAL_RANK', 0)))
  assert D.max_over_ranks(3.5) == 3.5 and D.sum_over_ranks(2.0) == 2.0
  assert D.all_reduce_gradients([]) == 0


def _gather_worker(rank, size, port, out):
  os.environ.update(RANK=str(rank), WORLD_SIZE=str(size), LOCAL_RANK=str(rank),
                    MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  D.init('gloo')
  g = torch.Generator().manual_seed(5)
  blocks = torch.randn(size, 3, 4, generator=g)            # every rank's prototype block
  weights = torch.randn(size, size * 3, 4, generator=g)    # every rank's loss weights on the bank
  x = blocks[rank].clone().requires_grad_(True)
  labels = torch.arange(3) + 10 * rank
  bank, bank_labels = D.all_gather_prototypes(x, labels)
  (bank * weights[rank]).sum().backward()
  want_bank = blocks.reshape(size * 3, 4)
  want_grad = weights[:, rank * 3:(rank + 1) * 3].sum(0)   # sum over ranks of the own block
  ids = D.global_segment_ids(torch.tensor([0, 2, -1]), 3)
  out[rank] = (torch.allclose(bank.detach(), want_bank), torch.allclose(x.grad, want_grad),
               bank_labels.tolist(), ids.tolist(), bool(bank_labels.requires_grad))
  dist.destroy_process_group()


def test_prototype_all_gather_and_its_backward():
  """SURVEY.md 8f-1 groundwork: the gathered bank is the concatenation of the rank blocks and
  its backward is the reduce-scatter of the bank gradients."""
  size, port = 2, _free_port()
  with mp.Manager() as m:
    out = m.dict()
    mp.spawn(_gather_worker, args=(size, port, out), nprocs=size, join=True)
    r0, r1 = out[0], out[1]
  assert r0[0] and r1[0] and r0[1] and r1[1]
  assert r0[2] == r1[2] == [0, 1, 2, 10, 11, 12]
  assert r0[3] == [0, 2, -1] and r1[3] == [3, 5, -1]
  assert not r0[4] and not r1[4]


def test_prototype_all_gather_single_process_is_identity():
  x, lab = torch.randn(3, 4), torch.arange(3)
  bank, bank_lab = D.all_gather_prototypes(x, lab)
  assert bank is x and bank_lab is lab
  assert D.global_segment_ids(lab, 3) is lab


def _exchange_worker(rank, size, port, out):
  os.environ.update(RANK=str(rank), WORLD_SIZE=str(size), LOCAL_RANK=str(rank),
                    MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  D.init('gloo')
  g = torch.Generator().manual_seed(11)
  counts = [3, 5]                                  # data-dependent block sizes
  protos = [torch.randn(c, 4, generator=g) for c in counts]
  protos_loc = [torch.randn(c, 6, generator=g) for c in counts]
  weights = [torch.randn(sum(counts), 4, generator=g) for _ in range(size)]
  x = protos[rank].clone().requires_grad_(True)
  xl = protos_loc[rank].clone().requires_grad_(True)
  m = counts[rank]
  sem, inst = torch.arange(m) + 10 * rank, torch.arange(m) + 100 * rank
  bid = torch.full((m,), rank, dtype=torch.long)
  cid = torch.tensor([0, m - 1, 1])
  bank, bank_loc, gsem, ginst, gbid, gcid = D.exchange_prototypes(x, xl, sem, inst, bid, cid)
  (bank * weights[rank]).sum().backward()
  off = sum(counts[:rank])
  want_grad = sum(w_[off:off + m] for w_ in weights)    # every rank's weights on the own block
  out[rank] = (torch.allclose(bank.detach(), torch.cat(protos)),
               torch.allclose(bank_loc.detach(), torch.cat(protos_loc)),
               torch.allclose(x.grad, want_grad), gsem.tolist(), gbid.tolist(), gcid.tolist(),
               xl.grad is None or float(xl.grad.abs().max()) == 0.0)
  dist.destroy_process_group()


def test_exchange_prototypes_variable_sizes():
  """rank blocks of different sizes, rank-major bank, gradients of every rank's
  loss reduce-scattered back to the owner."""
  size, port = 2, _free_port()
  with mp.Manager() as m:
    out = m.dict()
    mp.spawn(_exchange_worker, args=(size, port, out), nprocs=size, join=True)
    r0, r1 = out[0], out[1]
  for r in (r0, r1):
    assert r[0] and r[1] and r[2] and r[6]
    assert r[3] == [0, 1, 2, 10, 11, 12, 13, 14] and r[4] == [0, 0, 0, 1, 1, 1, 1, 1]
  assert r0[5] == [0, 2, 1] and r1[5] == [3, 7, 4]
