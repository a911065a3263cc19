"""Runs a few steps of the fixed-capacity head WITHOUT the CUDA graph, for ncu:

    ncu --set full --import-source on -k regex:segsort_fwd_tc -s 3 -c 1 -o out \
        python scripts/profile_step.py --workload voc_scribble_b4 --steps 3
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spml_b200 import synth  # noqa: E402
from spml_b200.static_head import StaticContrastiveHead  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='voc_scribble_b1')
ap.add_argument('--steps', type=int, default=3)
args = ap.parse_args()
w = synth.WORKLOADS[args.workload]
head = StaticContrastiveHead(synth.make_config(w), w.batch, w.height, w.width, w.loc_channels,
                             use_graph=False)
for s in range(args.steps):
  b = {k: v.cuda() for k, v in synth.make_batch(w, step=s).items()}
  out = head.step(b['embedding'], b['semantic_label'], b['instance_label'], b['semantic_tag'],
                  b['local_feature'])
torch.cuda.synchronize()
print('loss %.6f pixels %d segments %d' % (float(out['loss']), int(out['num_pixels']),
                                           int(out['num_segments'])))
