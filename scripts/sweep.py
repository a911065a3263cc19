"""BASELINE.json configs[4]: isolated-kernel sweep N_pix in {64k, 256k, 1M}, D = 128,
N_seg in {64, 256, 1024}, 10 k-means iterations.  Prints one JSON object per point with
the time, the algorithmic bytes / flops (SURVEY.md 8d formulas) and the fraction of the
binding roofline (MEASURED_PEAKS.json).  Every point it times is also CHECKED against the CPU
oracle ("parity" in the line; the script fails on a mismatch): k-means ids bit-exact,
SegSort loss within 1e-4 and gradients within 1e-3 of the fp64 oracle (all rows up to 256k
pixels; at 1M pixels the first 65 536 rows against the oracle plus linearity of the
row-sum over chunks).  Run on the GPU box:

    python scripts/sweep.py > gpurun_out/sweep.jsonl         # --no-check: timing only
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import spml_oracle as O  # noqa: E402  (the checker; never what is timed)
from spml_b200 import segsort_common, segsort_loss, synth  # noqa: E402

CHECK = '--no-check' not in sys.argv

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


def norm_err(a, b):
  a, b = a.double(), b.double()
  return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_kmeans(prob, m, got):
  want = O.spherical_kmeans(prob['embedding'], prob['seed_label'], m, T)
  bad = int((got.cpu() != want).sum())
  assert bad == 0, 'k-means: %d ids differ from the oracle' % bad
  return 'ids bit-exact vs oracle (%d pixels)' % want.numel()


def check_segsort(emb, labels, protos, psem, kappa):
  """Loss + gradients of rows [0, rows) against the fp64 oracle (row-chunked), and for larger
  problems linearity of the row sum over chunks on the GPU."""
  n, rows = emb.shape[0], min(emb.shape[0], 262144)
  e_cpu, lab_cpu = emb[:rows].cpu(), labels[:rows].cpu()
  p_cpu, psem_cpu = protos.cpu(), psem.cpu()
  pr = p_cpu.double().requires_grad_(True)
  total, de = torch.zeros((), dtype=torch.float64), torch.empty(rows, D, dtype=torch.float64)
  for c0 in range(0, rows, 16384):
    ec = e_cpu[c0:c0 + 16384].double().requires_grad_(True)
    part = O.segsort_loss(ec, psem_cpu[lab_cpu[c0:c0 + 16384]], lab_cpu[c0:c0 + 16384], pr,
                          psem_cpu, kappa, reduction='sum')
    part.backward()
    total += part.detach()
    de[c0:c0 + 16384] = ec.grad
  loss_fn = segsort_loss.SegSortLoss(kappa, reduction='sum')
  eg = emb[:rows].detach().clone().requires_grad_(True)
  pg = protos.detach().clone().requires_grad_(True)
  got = loss_fn(eg, psem[labels[:rows]], labels[:rows], pg, psem)
  got.backward()
  assert abs(float(got) - float(total)) <= 1e-4 * abs(float(total)), (float(got), float(total))
  assert norm_err(eg.grad.cpu(), de) < 1e-3 and norm_err(pg.grad.cpu(), pr.grad) < 1e-3
  note = 'loss, dE, dP vs fp64 oracle on %d rows' % rows
  if n > rows:
    whole = float(loss_fn(emb, psem[labels], labels, protos, psem))
    parts = sum(float(loss_fn(emb[c:c + rows], psem[labels[c:c + rows]], labels[c:c + rows],
                              protos, psem)) for c in range(0, n, rows))
    assert abs(whole - parts) <= 1e-5 * abs(parts), (whole, parts)
    note += '; row-sum linear over %d chunks' % (n // rows)
  return note


def report(kernel, n, m, ms, nbytes, flops, tensor, parity=None):
  t_hbm, t_tc = nbytes / HBM, (flops / TC if tensor else 0.0)
  bound = 'tensor' if t_tc > t_hbm else 'hbm'
  roof_ms = 1e3 * max(t_hbm, t_tc)
  print(json.dumps({'kernel': kernel, 'n_pix': n, 'n_seg': m, 'dim': D, 'ms': round(ms, 4),
                    'bound': bound, 'roofline_ms': round(roof_ms, 4),
                    'frac': round(roof_ms / ms, 4),
                    'achieved_gbs': round(nbytes / ms / 1e6, 1),
                    'achieved_tflops': round(flops / ms / 1e9, 2), 'parity': parity}), flush=True)


for n in (65536, 262144, 1048576):
  for m in (64, 256, 1024):
    prob = synth.sweep_problem(n, D, m)
    emb = prob['embedding'].cuda()
    seeds = prob['seed_label'].cuda()
    ms = timed(lambda: segsort_common.kmeans_with_initial_labels(emb, seeds, m, T))
    # the assignment GEMM runs on the tensor cores from K >= 256 (kmeans.cu::kmeans_use_tc), so
    # the bound is the larger of the HBM and the tensor time
    parity = check_kmeans(prob, m, segsort_common.kmeans_with_initial_labels(emb, seeds, m, T)) \
        if CHECK else None
    report('kmeans', n, m, ms, (T + 1) * n * D * 4 + T * n * 4 + 2 * T * m * D * 4,
           T * (2 * n * m * D + n * D), True, parity)

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
    parity = check_segsort(emb, labels, protos, psem, 12.0) if CHECK else None
    report('segsort_fwd_bwd', n, m, ms, n * (2 * D * 4 + 4 * D + 48) + 12 * m * D,
           8 * n * m * D, True, parity)
