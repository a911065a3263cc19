import sys, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from test_gpu_tensorcore import make_problem
from spml_b200 import _lib, ops
for (n, m, dim, kappa) in [(128, 128, 64, 6.0), (1000, 200, 64, 12.0), (5000, 1500, 128, 12.0)]:
  e, sem, seg, protos, psem = make_problem(n, m, dim, 5, n + m + 1, kappa)
  res = {}
  for path in ('fp32', 'tc'):
    ec, pc = e.cuda().requires_grad_(True), protos.cuda().requires_grad_(True)
    prob = ops.SegsortProblem(sem.cuda(), seg.cuda(), psem.cuda(), kappa, _lib.MODE_CLASS, path=path)
    loss = ops.SegsortLossFn.apply(ec, pc, prob)
    loss.backward()
    res[path] = (float(loss), ec.grad.cpu(), pc.grad.cpu())
  l, ge, gp = res['tc']
  print(n, m, dim, 'loss', res['fp32'][0], l, 'nan rows e', ge.isnan().any(1).nonzero().view(-1)[:10].tolist(), int(ge.isnan().any(1).sum()),
        'nan rows p', gp.isnan().any(1).nonzero().view(-1)[:10].tolist(), int(gp.isnan().any(1).sum()),
        'nan cols e', ge.isnan().any(0).nonzero().view(-1)[:10].tolist())
  ok = ~ge.isnan().any(1)
  print('  err on finite rows', float((ge[ok] - res['fp32'][1][ok]).abs().max()), float(res['fp32'][1].abs().max()))
