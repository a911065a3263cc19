"""cProfile of the Python side of the drop-in step (host-bound at batch 1).

    python scripts/pyprofile_step.py [workload] [steps]
"""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spml_b200 import synth  # noqa: E402
from spml_b200.head import ContrastiveHead  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'voc_scribble_b1'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
w = synth.WORKLOADS[name]
cfg = synth.make_config(w)
dev = torch.device('cuda', 0)
head = ContrastiveHead(cfg, variant=w.variant).to(dev)
res = [{k: v.to(dev) for k, v in synth.make_batch(w, step=s).items()} for s in range(4)]


def step(i):
  b = res[i % 4]
  emb = b['embedding'].detach().requires_grad_(True)
  out = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'], b['local_feature'],
             b.get('semantic_label_full'))
  out['loss'].backward()
  head.update_memory_bank(1)


for i in range(20):
  step(i)
torch.cuda.synchronize()
prof = cProfile.Profile()
prof.enable()
for i in range(steps):
  step(i)
torch.cuda.synchronize()
prof.disable()
st = pstats.Stats(prof)
st.sort_stats('tottime').print_stats(45)
