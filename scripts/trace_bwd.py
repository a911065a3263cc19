"""Timeline of one CTA of the tcgen05 backward kernel (needs a build with
`make -C spml_b200/csrc EXTRA=-DSPML_TC_TRACE`)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spml_oracle as O
from spml_b200 import _lib, ops
g = torch.Generator().manual_seed(1)
n, m, dim = 60000, 2304, 64
protos = O.l2_normalize(torch.randn(m, dim, generator=g)); seg = torch.randint(0, m, (n,), generator=g)
psem = torch.randint(0, 21, (m,), generator=g)
e = O.l2_normalize(protos[seg] + 0.5 * torch.randn(n, dim, generator=g))
for it in range(3):
  ec, pc = e.cuda().requires_grad_(True), protos.cuda().requires_grad_(True)
  prob = ops.SegsortProblem(psem[seg].cuda(), seg.cuda(), psem.cuda(), 12.0, _lib.MODE_CLASS, path='tc')
  ops.SegsortLossFn.apply(ec, pc, prob).backward()
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (64 * 16))()
lib.spml_debug_tc_trace.argtypes = [ctypes.c_void_p]
assert lib.spml_debug_tc_trace(buf) == 0
t = torch.tensor(list(buf)).view(64, 16)
base = int(t[0, 8])
names = {0: 'epi wait t_full', 1: 'epi got t_full', 2: 'epi G computed', 3: 'epi got g_empty', 4: 'epi G stored', 5: 'epi arrived g_full',
         8: 'mma gemm1 begin', 9: 'mma got t_empty', 10: 'mma got b_full', 11: 'mma gemm1 issued', 12: 'mma wait g_full', 13: 'mma got g_full', 14: 'mma gemm2 issued'}
for j in range(0, 12):
  ev = sorted((int(t[j, k]) - base, names[k]) for k in names if t[j, k] != 0)
  print('tile', j, ' '.join('%s@%d' % (nm, ts) for ts, nm in ev))
per = [int(t[j + 1, 1] - t[j, 1]) for j in range(4, 30)]
print('cycles per step (epilogue t_full to t_full):', per)
print('epilogue compute (1->2):', [int(t[j, 2] - t[j, 1]) for j in range(4, 16)])
print('epilogue store (3->4):', [int(t[j, 4] - t[j, 3]) for j in range(4, 16)])
print('epilogue wait t_full (0->1):', [int(t[j, 1] - t[j, 0]) for j in range(4, 16)])
print('mma: g_full wait (12->13):', [int(t[j, 13] - t[j, 12]) for j in range(4, 16)])
print('mma: gemm2 issue (13->14):', [int(t[j, 14] - t[j, 13]) for j in range(4, 16)])
print('mma: gemm1 issue (10->11):', [int(t[j, 11] - t[j, 10]) for j in range(4, 16)])
