"""BASELINE.json configs[4]: isolated-kernel sweep N_pix in {64k, 256k, 1M}, D = 128,
N_seg in {64, 256, 1024}, 10 k-means iterations.  Prints one JSON object per point with
the time, the algorithmic bytes / flops (SURVEY.md 8d formulas) and the fraction of the
binding roofline (MEASURED_PEAKS.json).  Run on the GPU box:

    python scripts/sweep.py > gpurun_out/sweep.jsonl
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spml_b200 import segsort_common, segsort_loss, synth  # noqa: E402

peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
    os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
HBM, TC = peaks['hbm_gbs'] * 1e9, peaks['bf16_tflops'] * 1e12
D, T = 128, 10
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, reps=3):
  fn()
  torch.cuda.synchronize()
  ms = []
  for _ in range(reps):
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    fn()
    e.record()
    torch.cuda.synchronize()
    ms.append(s.elapsed_time(e))
  return min(ms)


def report(kernel, n, m, ms, nbytes, flops, tensor):
  t_hbm, t_tc = nbytes / HBM, (flops / TC if tensor else 0.0)
  bound = 'tensor' if t_tc > t_hbm else 'hbm'
  roof_ms = 1e3 * max(t_hbm, t_tc)
  print(json.dumps({'kernel': kernel, 'n_pix': n, 'n_seg': m, 'dim': D, 'ms': round(ms, 4),
                    'bound': bound, 'roofline_ms': round(roof_ms, 4),
                    'frac': round(roof_ms / ms, 4),
                    'achieved_gbs': round(nbytes / ms / 1e6, 1),
                    'achieved_tflops': round(flops / ms / 1e9, 2)}), flush=True)


for n in (65536, 262144, 1048576):
  for m in (64, 256, 1024):
    prob = synth.sweep_problem(n, D, m)
    emb = prob['embedding'].cuda()
    seeds = prob['seed_label'].cuda()
    ms = timed(lambda: segsort_common.kmeans_with_initial_labels(emb, seeds, m, T))
    # the assignment GEMM runs on the tensor cores from K >= 256 (kmeans.cu::kmeans_use_tc), so
    # the bound is the larger of the HBM and the tensor time
    report('kmeans', n, m, ms, (T + 1) * n * D * 4 + T * n * 4 + 2 * T * m * D * 4,
           T * (2 * n * m * D + n * D), True)

    g = torch.Generator().manual_seed(n + m)
    labels = segsort_common.kmeans_with_initial_labels(emb, seeds, m, 2)
    psem = torch.randint(0, 21, (m,), generator=g).cuda()
    protos = segsort_common.calculate_prototypes_from_labels(emb, labels, m).detach()
    loss_fn = segsort_loss.SegSortLoss(12.0)

    def step():
      e = emb.detach().requires_grad_(True)
      p = protos.detach().requires_grad_(True)
      loss_fn(e, psem[labels], labels, p, psem).backward()
    ms = timed(step)
    report('segsort_fwd_bwd', n, m, ms, n * (2 * D * 4 + 4 * D + 48) + 12 * m * D,
           8 * n * m * D, True)
