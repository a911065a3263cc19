"""Runs a few steps of the contrastive head for ncu, without CUDA graphs:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv \
        python scripts/profile_step.py --workload voc_scribble_b1 --steps 3
    ncu --set full --clock-control none --import-source on -k regex:kmeans_small -s 2 -c 1 \
        -o out python scripts/profile_step.py --steps 3

--api dropin (default): the reference operator API (ContrastiveHead: generate_clusters ->
gather_clustering_and_update_prototypes -> Segsort.forward -> backward -> memory bank), i.e. the
path bench.py times;  --api static: the fixed-capacity head.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spml_b200 import synth  # noqa: E402
from spml_b200.head import ContrastiveHead  # noqa: E402
from spml_b200.static_head import StaticContrastiveHead  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='voc_scribble_b1')
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--api', default='dropin', choices=['dropin', 'static'])
args = ap.parse_args()
w = synth.WORKLOADS[args.workload]
cfg = synth.make_config(w)
if args.api == 'static':
  head = StaticContrastiveHead(cfg, w.batch, w.height, w.width, w.loc_channels, use_graph=False)
else:
  head = ContrastiveHead(cfg, variant=w.variant).cuda()
for s in range(args.steps):
  b = {k: v.cuda() for k, v in synth.make_batch(w, step=s).items()}
  if args.api == 'static':
    out = head.step(b['embedding'], b['semantic_label'], b['instance_label'], b['semantic_tag'],
                    b['local_feature'])
  else:
    emb = b['embedding'].requires_grad_(True)
    out = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'],
               b['local_feature'], b.get('semantic_label_full'))
    out['loss'].backward()
    head.update_memory_bank(1)
torch.cuda.synchronize()
print('loss %.6f' % float(out['loss']))
