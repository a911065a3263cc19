"""A/B timing inside one process: the drop-in step and its clustering stage, alternating
between settings of an environment switch that the library reads per call, or between
sub-processes for switches read once.  Usage: python scripts/ab_step.py ENV_NAME A B [workload]"""
import os, subprocess, sys, json
if len(sys.argv) >= 4 and sys.argv[1] != '--child':
  name, a, b = sys.argv[1:4]
  wl = sys.argv[4] if len(sys.argv) > 4 else 'voc_scribble_b1'
  for rnd in range(3):
    for v in (a, b):
      env = dict(os.environ); env[name] = v
      out = subprocess.run([sys.executable, __file__, '--child', wl], env=env, capture_output=True, text=True)
      print('%s=%s  %s' % (name, v, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]))
  sys.exit(0)
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spml_b200 import synth
from spml_b200.head import ContrastiveHead, generate_clusters
wl = sys.argv[2]
w = synth.WORKLOADS[wl]
cfg = synth.make_config(w)
head = ContrastiveHead(cfg).cuda()
batches = [{k: v.cuda() for k, v in synth.make_batch(w, step=s).items()} for s in range(4)]
flush = torch.empty(64 << 20, dtype=torch.float32, device='cuda')
def step(b):
  emb = b['embedding'].clone().requires_grad_(True)
  out = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'], b['local_feature'])
  out['loss'].backward()
  return out
def clusters(b):
  return generate_clusters(b['embedding'], b['semantic_label'], b['instance_label'], b['local_feature'],
                           w.label_divisor, w.ignore_index, list(w.num_clusters), w.iterations)
res = {}
for nm, fn in (('step', step), ('clustering stage', clusters)):
  for i in range(10): fn(batches[i % 4])
  torch.cuda.synchronize()
  ts = []
  for rep in range(15):
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): fn(batches[i % 4])
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 20)
  ts.sort()
  res[nm] = 'median %.1f us min %.1f us' % (ts[len(ts) // 2] * 1e3, ts[0] * 1e3)
print(json.dumps(res))
