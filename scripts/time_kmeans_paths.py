"""spml_kmeans alone on the clustering input of a workload, once per kernel (SPML_B200_KMEANS),
CUDA-event timed, labels compared across the kernels.
    python scripts/time_kmeans_paths.py [workload ...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spml_b200 import _lib, ops, synth, segsort_common
from spml_b200 import general_common as G

def clustering_input(w, step=0):
  b = synth.make_batch(w, step=step)
  e = b['embedding'].cuda()
  B, D, H, W = e.shape
  e = G.normalize_embedding(e.permute(0, 2, 3, 1).contiguous())
  loc = segsort_common.generate_location_features((H, W), 'cuda', 'float').unsqueeze(0).expand(B, H, W, 2)
  el = G.normalize_embedding(torch.cat([e, loc], -1)).reshape(B * H * W, D + 2).contiguous()
  lab0 = segsort_common.initialize_cluster_labels(list(w.num_clusters), (H, W), 'cuda')
  lab0 = lab0.reshape(1, -1).expand(B, -1).reshape(-1).to(torch.int32).contiguous()
  off = torch.arange(B + 1, dtype=torch.int32, device='cuda') * (H * W)
  return el, off, B, H * W, w.num_clusters[0] * w.num_clusters[1], lab0

import dataclasses
for name in (sys.argv[1:] or ['voc_scribble_b1', 'voc_scribble_b4', 'voc_tag_b2']):
  # NAME or NAME@BATCH (the workload at another batch size)
  base, _, nb = name.partition('@')
  w = synth.WORKLOADS[base]
  if nb:
    w = dataclasses.replace(w, batch=int(nb))
  el, off, B, n, K, lab0 = clustering_input(w)
  lib = _lib.load()
  print(name, 'default path:', lib.spml_debug_kmeans_path(B, n, el.shape[1], K))
  got = {}
  for path in ('cluster', 'small', 'fp32'):
    os.environ['SPML_B200_KMEANS'] = path
    for _ in range(5):
      out, _ = ops.kmeans(el, off, B, n, K, w.iterations, lab0, want_i64=False)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 50
    ev[0].record()
    for _ in range(reps):
      out, _ = ops.kmeans(el, off, B, n, K, w.iterations, lab0, want_i64=False)
    ev[1].record()
    torch.cuda.synchronize()
    got[path] = out.clone()
    print('  %-8s %8.1f us / call   %s' % (path, ev[0].elapsed_time(ev[1]) * 1e3 / reps,
          'labels == fp32: %s' % bool(torch.equal(out, got.get('fp32', out))) if path == 'fp32' else ''))
  for path in ('cluster', 'small'):
    print('  %s == fp32: %s (%d differ)' % (path, bool(torch.equal(got[path], got['fp32'])),
                                          int((got[path] != got['fp32']).sum())))
