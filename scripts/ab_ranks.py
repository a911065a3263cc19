"""Two (or N) ranks on one box, each timing the drop-in step with the library named by
SPML_B200_LIB (ctypes binding), all ranks running at the same time: does a host-side change hold
up when several processes share the host?   torchrun --nproc-per-node N scripts/ab_ranks.py"""
import os, sys, time, json
os.environ['SPML_B200_BINDING'] = 'ctypes'
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(rank)
from spml_b200 import synth
from spml_b200.head import ContrastiveHead
w = synth.WORKLOADS['voc_scribble_b1']
cfg = synth.make_config(w)
head = ContrastiveHead(cfg).cuda()
batches = [{k: v.cuda() for k, v in synth.make_batch(w, step=s).items()} for s in range(4)]
def step(b):
  emb = b['embedding'].clone().requires_grad_(True)
  out = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'], b['local_feature'])
  out['loss'].backward()
for i in range(20): step(batches[i % 4])
torch.cuda.synchronize()
# crude barrier through the file system: start the timed region together
open('/tmp/ab_ready_%d' % rank, 'w').close()
world = int(os.environ.get('WORLD_SIZE', '1'))
while not all(os.path.exists('/tmp/ab_ready_%d' % r) for r in range(world)): time.sleep(0.001)
ts = []
for rep in range(10):
  torch.cuda.synchronize(); t0 = time.perf_counter()
  for i in range(50): step(batches[i % 4])
  torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) / 50 * 1e6)
ts.sort()
print('rank %d %s: median %.1f us min %.1f us' % (rank, os.path.basename(os.environ.get('SPML_B200_LIB', 'default')), ts[5], ts[0]), flush=True)
time.sleep(0.5)
try: os.remove('/tmp/ab_ready_%d' % rank)
except OSError: pass
