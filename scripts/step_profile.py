"""Where a step of the drop-in path spends its time: wall-clock per phase (host launch time
and, after a synchronise, GPU completion time) and a torch.profiler table of the host side.

    python scripts/step_profile.py [--workload W] [--e2e] [--trace out.json]
"""

import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from spml_b200 import model_utils, synth  # noqa: E402
from spml_b200.head import ContrastiveHead, generate_clusters  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--workload', default='voc_scribble_b1')
  ap.add_argument('--e2e', action='store_true')
  ap.add_argument('--trace', default='')
  ap.add_argument('--steps', type=int, default=30)
  args = ap.parse_args()
  w = synth.WORKLOADS[args.workload]
  cfg = synth.make_config(w)
  dev = torch.device('cuda', 0)
  head = ContrastiveHead(cfg, variant=w.variant).to(dev)
  host = [{k: v.pin_memory() for k, v in synth.make_batch(w, step=s).items()} for s in range(4)]
  res = [{k: v.to(dev) for k, v in b.items()} for b in host]

  def step(i, marks=None):
    def mark(name):
      if marks is not None:
        t_launch = time.perf_counter()
        torch.cuda.synchronize()
        marks.append((name, t_launch, time.perf_counter()))
    b = res[i % 4]
    if args.e2e:
      b = {k: v.to(dev, non_blocking=True) for k, v in host[i % 4].items()}
      mark('h2d')
    emb = b['embedding'].detach().requires_grad_(True)
    datas = generate_clusters(emb, b['semantic_label'], b['instance_label'], b['local_feature'],
                              cfg.network.label_divisor, cfg.dataset.semantic_ignore_index,
                              cfg.network.kmeans_num_clusters, cfg.network.kmeans_iterations,
                              batch_index_offset=0, densepose=w.variant == 'densepose')
    mark('generate_clusters')
    out = model_utils.gather_clustering_and_update_prototypes(
        [datas['cluster_embedding']], [datas['cluster_embedding_with_loc']],
        [datas['cluster_index']], [datas['cluster_batch_index']],
        [datas['cluster_semantic_label']], [datas['cluster_instance_label']])
    mark('gather')
    protos, protos_loc, psem, pinst, pbid, cids = out
    datas['embedding'] = emb
    targets = {'prototype': protos[0], 'prototype_with_loc': protos_loc[0],
               'prototype_semantic_label': psem[0], 'prototype_instance_label': pinst[0],
               'prototype_batch_index': pbid[0], 'semantic_label': b['semantic_label'],
               'semantic_tag': b['semantic_tag'],
               'prototype_semantic_tag': torch.index_select(b['semantic_tag'], 0, pbid[0])}
    targets.update(head.memory_banks)
    mark('targets')
    o = head.predictor(datas, targets)
    mark('predictor')
    loss = getattr(head.predictor, 'last_loss_total', None)     # as ContrastiveHead.forward does
    if loss is None:
      loss = sum(o[k] for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss') if o[k] is not None)
    head.predictor.last_loss_total = None
    mark('sum')
    loss.backward()
    mark('backward')
    head._last_targets, head._last_batch = targets, emb.shape[0]
    head.update_memory_bank(1)
    mark('bank')

  for i in range(5):
    step(i)
  torch.cuda.synchronize()
  # 1. phases with a synchronise after each (host launch time | GPU tail after the launch)
  acc = {}
  for i in range(args.steps):
    marks = []
    t0 = time.perf_counter()
    step(5 + i, marks)
    prev = t0
    for name, t_launch, t_done in marks:
      a = acc.setdefault(name, [0.0, 0.0])
      a[0] += t_launch - prev
      a[1] += t_done - t_launch
      prev = t_done
  print('phase                 host launch us   gpu tail us   (synchronised after every phase)')
  for name, (h, g) in acc.items():
    print('%-20s %12.1f %12.1f' % (name, 1e6 * h / args.steps, 1e6 * g / args.steps))
  # 2. free-running wall time
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for i in range(args.steps):
    step(100 + i)
  torch.cuda.synchronize()
  print('free-running: %.1f us / step' % (1e6 * (time.perf_counter() - t0) / args.steps))
  # 3. host-side profile
  from torch.profiler import ProfilerActivity, profile
  with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(10):
      step(200 + i)
    torch.cuda.synchronize()
  print(prof.key_averages().table(sort_by='self_cpu_time_total', row_limit=45,
                                  max_name_column_width=60))
  if args.trace:
    prof.export_chrome_trace(args.trace)


if __name__ == '__main__':
  main()
