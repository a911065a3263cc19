"""bench.py - throughput of the pixel-to-segment contrastive hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A "step" is one pass of the hot path over one synthetic minibatch: generate_clusters
(normalise + spherical k-means + segment ids) -> segment prototypes -> sem_ann /
sem_occ / img_sim losses + top-5 accuracy -> backward to d(embedding) -> memory-bank
update (pyscripts/train/train.py:167-219,273-293 of the reference, minus the cuDNN
backbone, which is not part of the path).  The default workload is BASELINE.json
configs[1]: VOC12 scribble, batch 1 per GPU, 512x512 crop = a 128x128x64 embedding
map, 6x6 seeds, 10 k-means iterations, memory bank of 2 steps.

It prints ONE JSON line (see the keys below).  `value` is measured with the inputs
resident in HBM; `e2e` with the inputs in pinned host memory, copied in, and the
loss + d(embedding) copied back inside the timed region.  Multi-GPU: the path
shards over images, every rank runs its own minibatch (weak scaling, no data-path
collective; the head has no parameters so DDP would add none).

`--impl reference` times the CPU oracle (a port of the reference with the same ATen
ops; the Python reference itself cannot travel to the GPU box) on all host cores.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

from spml_b200 import synth  # noqa: E402

METRIC = 'contrastive-loss step images/sec (512x512 crop, 128x128 embedding map per image)'
POOL = 8           # distinct synthetic minibatches cycled through
FLUSH_BYTES = 256 << 20


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=30)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--workload', default='voc_scribble_b1', choices=sorted(synth.WORKLOADS))
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-sweep', action='store_true')
  return ap.parse_args()


def load_peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    p = json.load(open(path))
    return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
            'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']),
            'source': 'measured'}
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
          'source': 'fallback'}


# ------------------------------------------------------------------------------ clocks


class ClockSampler(threading.Thread):
  """nvidia-smi clocks / throttle reasons during the timed region."""

  QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
           'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    super().__init__(daemon=True)
    self.gpu_index = gpu_index
    self.rows = []
    self.proc = None

  def run(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.QUERY,
           '--format=csv,noheader,nounits', '-lms', '100'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      for line in self.proc.stdout:
        self.rows.append([c.strip() for c in line.split(',')])
    except Exception:       # nvidia-smi missing: report nothing rather than fail
      pass

  def stop(self):
    if self.proc is not None:
      self.proc.terminate()
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for r in self.rows:
      try:
        sm.append(float(r[1]))
        mx.append(float(r[2]))
      except (ValueError, IndexError):
        continue
      for name, val in zip(names, r[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    if not sm:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
    return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons),
            'samples': len(sm)}


# ------------------------------------------------------------------------------ reference arm


def run_reference(args):
  """The reference's CPU implementation of the path (oracle port), all host threads."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  from oracle import spml_oracle as O
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  w = synth.WORKLOADS[args.workload]
  cfg = synth.make_config(w)
  batches = [synth.make_batch(w, step=s) for s in range(min(POOL, args.steps + args.warmup))]
  bank = {}
  times = []
  for s in range(args.warmup + args.steps):
    batch = batches[s % len(batches)]
    t0 = time.perf_counter()
    out = O.contrastive_step(cfg, batch, bank)
    O.memory_bank_update(bank, {k: out[k] for k in out if k.startswith('prototype')},
                         w.memory_bank_size, w.batch)
    dt = time.perf_counter() - t0
    if s >= args.warmup:
      times.append(dt)
  total = sum(times)
  value = w.batch * len(times) / total
  sample = '%d steps of workload %s (batch %d), fp32, torch %s CPU ops' % (
      len(times), w.name, w.batch, torch.__version__)
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'images/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': workload_config(w, args.gpus),
      'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                       'sample': sample},
      'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
  }
  print(json.dumps(line))


def workload_config(w, n_gpus):
  return {'workload': w.name, 'images_per_gpu': w.batch, 'global_batch': w.batch * n_gpus,
          'embedding_map': [w.height, w.width], 'embedding_dim': w.dim,
          'kmeans_seeds': list(w.num_clusters), 'kmeans_iterations': w.iterations,
          'memory_bank_steps': w.memory_bank_size, 'regions_per_image': w.num_regions,
          'parallelism': 'images sharded over %d GPU(s), no data-path collective' % n_gpus,
          'l2': 'flushed with a %d MiB write between timed steps' % (FLUSH_BYTES >> 20)}


# ------------------------------------------------------------------------------ our arm


def algorithmic_bytes(name, shapes):
  """Algorithmic HBM bytes of one call of an entry point (SURVEY.md section 8d formulas,
  fp32 embeddings: b_e = 4)."""
  n, d, dl, m, k, t = (shapes[x] for x in ('n', 'd', 'dl', 'm_all', 'k', 't'))
  if name == 'spml_kmeans':
    return (t + 1) * n * dl * 4 + t * n * 4 + 2 * t * k * dl * 4
  if name == 'spml_normalize_pack_fwd':
    return n * d * 4 + n * (d + dl) * 4 + n * 32
  if name == 'spml_normalize_pack_bwd':
    return n * (2 * d + 2 * dl) * 4 + n * d * 4
  if name == 'spml_segsort_fwd':
    return n * (d * 4 + 24 + 12) + m * d * 4
  if name == 'spml_segsort_bwd':
    return n * (2 * d * 4 + d * 4 + 24 + 12) + 3 * m * d * 4
  if name == 'spml_segment_prototypes_fwd':
    return n * d * 4 + n * 8 + 3 * m * d * 4
  if name == 'spml_segment_prototypes_bwd':
    return n * d * 4 + n * 8 + 2 * m * d * 4
  return None


def run_b200(args):
  import torch.distributed as dist
  from spml_b200 import _lib
  from spml_b200.head import ContrastiveHead

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device (the product has no CPU path)')
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  _lib.load()

  w = synth.WORKLOADS[args.workload]
  cfg = synth.make_config(w)
  head = ContrastiveHead(cfg).to(dev)

  # every rank gets its own minibatches (different seeds): weak scaling over images
  host = [synth.make_batch(w, seed=235 + rank, step=s) for s in range(POOL)]
  host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
  resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
  flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
  h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
  grad_host = torch.empty(host[0]['embedding'].shape, dtype=torch.float32).pin_memory()
  loss_host = torch.empty(4, dtype=torch.float32).pin_memory()
  d2h_bytes = grad_host.numel() * 4 + loss_host.numel() * 4

  def step_resident(i):
    b = resident[i % POOL]
    emb = b['embedding'].detach().requires_grad_(True)
    out = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'],
               b['local_feature'])
    out['loss'].backward()
    head.update_memory_bank(world)
    return out, emb.grad

  def step_e2e(i):
    hb = host[i % POOL]
    b = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
    emb = b['embedding'].requires_grad_(True)
    out = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'],
               b['local_feature'])
    out['loss'].backward()
    head.update_memory_bank(world)
    grad_host.copy_(emb.grad, non_blocking=True)
    loss_host.copy_(torch.stack([out['sem_ann_loss'].detach(), out['sem_occ_loss'].detach(),
                                 out['img_sim_loss'].detach(), out['accuracy']]),
                    non_blocking=True)
    return out

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  def timed(step_fn, steps, warmup):
    """Per-step CUDA-event timing on the launching stream with an L2 flush (untimed)
    between steps; returns the summed milliseconds of `steps` steps."""
    head.memory_banks.clear()
    for i in range(warmup):
      step_fn(i)
    barrier()
    pairs = []
    launches0 = _lib.launch_count()
    wall0 = time.perf_counter()
    for i in range(steps):
      flush.fill_(i & 0xff)
      s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      s.record()
      step_fn(warmup + i)
      e.record()
      pairs.append((s, e))
    barrier()
    wall = time.perf_counter() - wall0
    ms = sum(s.elapsed_time(e) for s, e in pairs)
    return ms, _lib.launch_count() - launches0, wall

  sampler = ClockSampler(local_rank) if rank == 0 else None
  if sampler:
    sampler.start()
    time.sleep(0.3)
  ms_res, launches, wall_res = timed(step_resident, args.steps, max(args.warmup, 3))
  ms_e2e, _, wall_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))
  clocks = sampler.stop() if sampler else None

  def max_over_ranks(x):
    if world == 1:
      return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

  ms_res, ms_e2e = max_over_ranks(ms_res), max_over_ranks(ms_e2e)
  images = w.batch * world * args.steps

  # ---- per-entry-point profile of a few steps (separate pass; events per C-ABI call)
  roofline, breakdown = None, None
  if rank == 0:
    head.memory_banks.clear()
    for i in range(3):
      step_resident(i)
    torch.cuda.synchronize()
    _lib.PROFILE = []
    prof_steps = 5
    for i in range(prof_steps):
      flush.fill_(i)
      out, _ = step_resident(3 + i)
    torch.cuda.synchronize()
    records, _lib.PROFILE = _lib.PROFILE, None
    agg = {}
    for name, s, e, kernels in records:
      a = agg.setdefault(name, {'ms': 0.0, 'calls': 0, 'kernels': 0})
      a['ms'] += s.elapsed_time(e)
      a['calls'] += 1
      a['kernels'] += kernels
    total_ms = sum(a['ms'] for a in agg.values())
    breakdown = {k: {'ms_per_step': round(v['ms'] / prof_steps, 4),
                     'calls_per_step': v['calls'] / prof_steps,
                     'kernels_per_step': v['kernels'] / prof_steps,
                     'share': round(v['ms'] / total_ms, 4)}
                 for k, v in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])}
    top = next(iter(breakdown))
    n_rows = out['datas']['cluster_index'].shape[0]
    m_cur = out['targets']['prototype'].shape[0]
    m_all = m_cur * (1 + len(head.memory_banks.get('memory_prototype', [])))
    shapes = {'n': n_rows, 'd': w.dim, 'dl': w.dim + w.loc_channels, 'm_all': m_all,
              'k': w.num_clusters[0] * w.num_clusters[1], 't': w.iterations}
    peaks = load_peaks()
    alg = algorithmic_bytes(top, shapes)
    call_ms = agg[top]['ms'] / agg[top]['calls']
    if alg is not None:
      achieved = alg / (call_ms * 1e-3) / 1e9
      roofline = {'kernel': top, 'bound': 'hbm', 'achieved': round(achieved, 2),
                  'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                  'frac': round(achieved / peaks['hbm_gbs'], 5), 'traffic': None,
                  'peak_source': peaks['source'],
                  'algorithmic_bytes_per_call': alg, 'ms_per_call': round(call_ms, 4),
                  'launches_per_call': agg[top]['kernels'] / agg[top]['calls'],
                  'shapes': shapes}

  # ---- CPU baseline (oracle port) on this box's host cores, rank 0, N == 1 only
  cpu_baseline = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    from oracle import spml_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cpu_batches = [synth.make_batch(w, seed=235, step=s) for s in range(4)]
    bank, times = {}, []
    budget_end = time.perf_counter() + 20.0
    for s in range(2 + 12):
      t0 = time.perf_counter()
      o = O.contrastive_step(cfg, cpu_batches[s % 4], bank)
      O.memory_bank_update(bank, {k: o[k] for k in o if k.startswith('prototype')},
                           w.memory_bank_size, w.batch)
      if s >= 2:
        times.append(time.perf_counter() - t0)
      if time.perf_counter() > budget_end and len(times) >= 3:
        break
    cpu_baseline = {'value': w.batch * len(times) / sum(times), 'unit': 'images/s',
                    'cores': cores, 'kind': 'port',
                    'sample': '%d steps of %s after 2 warm-ups, fp32 torch CPU ops, %d threads'
                              % (len(times), w.name, cores),
                    'ms_per_step': 1e3 * sum(times) / len(times)}

  if rank == 0:
    value = images / (ms_res * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_res / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(w, world),
        'contrastive_step_ms': ms_res / args.steps,
        'e2e': {'value': images / (ms_e2e * 1e-3), 'unit': 'images/s',
                'ms_per_step': ms_e2e / args.steps, 'h2d_bytes_per_step': h2d_bytes,
                'd2h_bytes_per_step': d2h_bytes},
        'gpu_launches': launches,
        'gpu_launches_per_step': launches / args.steps,
        'wall_s': {'resident': round(wall_res, 4), 'e2e': round(wall_e2e, 4)},
        'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
        'breakdown': breakdown,
    }
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  args = parse_args()
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_b200(args)


if __name__ == '__main__':
  main()
