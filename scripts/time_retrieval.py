"""Segsort.predictions (nearest-neighbour retrieval inference, predictions/segsort.py:68-125) at
the scale of one full-resolution image: 262 144 pixels in 576 segments against a memory bank of
10 000 / 50 000 / 100 000 prototypes, k = 20.  The reference retrieves per SEGMENT prototype (a few
hundred queries per image), not per pixel.  CUDA-event timed, against the same call with the
prototypes and the top-k replaced by ATen (torch.mm + topk) as a yardstick."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spml_b200 import ops, predictions, synth, segsort_common

def timed(fn, reps=20):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(reps): fn()
  e.record(); torch.cuda.synchronize()
  return s.elapsed_time(e) / reps

g = torch.Generator().manual_seed(21)
n, segs, dim = 262144, 576, 64
centres = torch.nn.functional.normalize(torch.randn(200, dim, generator=g), dim=1)
cid = torch.randint(0, segs, (n,), generator=g)
seg_centre = torch.randint(0, 200, (segs,), generator=g)
emb = torch.nn.functional.normalize(centres[seg_centre[cid]] + 0.4 * torch.randn(n, dim, generator=g), dim=1).cuda()
cid = cid.cuda()
model = predictions.segsort(synth.make_config(synth.WORKLOADS['tiny']))
for m in (10000, 50000, 100000):
  bank_c = torch.randint(0, 200, (m,), generator=g)
  bank = torch.nn.functional.normalize(centres[bank_c] + 0.4 * torch.randn(m, dim, generator=g), dim=1).cuda()
  lab = (bank_c % 21).cuda()
  datas = {'cluster_embedding': emb, 'cluster_index': cid}
  targets = {'semantic_memory_prototype': bank, 'semantic_memory_prototype_label': lab}
  ms = timed(lambda: model.predictions(datas, targets))
  protos = segsort_common.calculate_prototypes_from_labels(emb, cid, segs)
  def aten():
    sim = torch.mm(protos, bank.t())
    return lab[sim.topk(20, dim=1).indices]
  ms_aten = timed(aten)
  qlab = torch.zeros(segs, dtype=torch.int64, device='cuda')
  ms_topk = timed(lambda: ops.topk_ranking(protos, qlab, bank, lab, 20))
  flops = 2.0 * segs * m * dim
  print('bank %6d: predictions() %.3f ms (prototypes + top-20 + majority vote + gather to %d pixels); '
        'top_k_ranking alone %.3f ms; torch.mm + topk on ready prototypes %.3f ms; similarity work %.2f GFLOP' % (m, ms, n, ms_topk, ms_aten, flops / 1e9))
