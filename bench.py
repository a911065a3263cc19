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
losses + accuracy copied back inside the timed region.  Multi-GPU: the path
shards over images, every rank runs its own minibatch (weak scaling, no data-path
collective; the head has no parameters so DDP would add none).

`value` / `e2e` are measured through the reference's operator API (the symbols
spml_b200.install() rebinds, composed as train.py does); the fixed-capacity CUDA-graph head
is reported next to them under `static_head`.

`--impl reference` times the reference's own CPU implementation on all host cores: the copy of
twke18/SPML under baseline/_ref when it is there (scripts/install_reference.py), else the
oracle port (same ATen CPU kernels in the same order).
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
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--workload', default='voc_scribble_b1', choices=sorted(synth.WORKLOADS))
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--repeats', type=int, default=0,
                  help='timed regions of --steps steps each (0: enough for ~1.2 s of load)')
  ap.add_argument('--backbone-dtype', default='bf16', choices=['bf16', 'fp32'],
                  help='train_* workloads: autocast dtype of the cuDNN backbone')
  ap.add_argument('--no-graph', action='store_true',
                  help='launch the kernels directly instead of replaying the CUDA graph (for ncu)')
  ap.add_argument('--no-train-arm', action='store_true',
                  help='contrastive-step workloads: skip the secondary training-step measurement '
                       '(ResNet-101 backbone + head + one NCCL all-reduce) reported under '
                       '"training_step"')
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
  """SM clock and throttle reasons of one GPU during the timed region, every 100 ms, through
  NVML in-process (nvidia_ml_py).  A looping `nvidia-smi` subprocess does the same job but its
  queries stall the CUDA calls of the process that drives the sampled GPU (measured: +15 ms per
  34 ms training step on rank 0 of a 2-GPU run), so it is only the fallback, at 500 ms."""

  QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
           'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
  # NVML clocks-event (throttle) reason bits
  REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
             0x4: 'sw_power_cap'}

  def __init__(self, gpu_index, period=0.1):
    super().__init__(daemon=True)
    self.gpu_index = gpu_index
    self.period = period
    self.query_ms = []
    self.rows = []            # (sm_mhz, max_mhz, reason bits or list of names)
    self.proc = None
    self._stop_flag = threading.Event()
    self.how = None

  def _nvml_handle(self):
    import pynvml
    pynvml.nvmlInit()
    uuid = str(torch.cuda.get_device_properties(self.gpu_index).uuid)
    if not uuid.startswith('GPU-'):
      uuid = 'GPU-' + uuid
    try:
      return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
    except Exception:
      return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)

  def run(self):
    try:
      nv, h = self._nvml_handle()
      mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
      get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
          nv.nvmlDeviceGetCurrentClocksThrottleReasons
      self.how = 'nvml'
      while not self._stop_flag.is_set():
        t0 = time.perf_counter()
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        bits = int(get_reasons(h))
        self.query_ms.append(1e3 * (time.perf_counter() - t0))
        self.rows.append((float(sm), float(mx), [n for b, n in self.REASONS.items() if bits & b]))
        self._stop_flag.wait(self.period)
      return
    except Exception:
      pass
    try:
      self.how = 'nvidia-smi'
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.QUERY,
           '--format=csv,noheader,nounits', '-lms', '500'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
      for line in self.proc.stdout:
        r = [c.strip() for c in line.split(',')]
        try:
          self.rows.append((float(r[1]), float(r[2]),
                            [n for n, v in zip(names, r[5:9]) if v.lower().startswith('active')]))
        except (ValueError, IndexError):
          continue
    except Exception:       # neither NVML nor nvidia-smi: report nothing rather than fail
      pass

  def stop(self):
    self._stop_flag.set()
    if self.proc is not None:
      self.proc.terminate()
    self.join(timeout=2.0)
    if not self.rows:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0, 'how': self.how}
    sm = [r[0] for r in self.rows]
    reasons = sorted({n for r in self.rows for n in r[2]})
    return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(r[1] for r in self.rows),
            'reasons': reasons, 'samples': len(sm), 'how': self.how,
            'period_s': self.period,
            'query_ms': round(statistics.median(self.query_ms), 3) if self.query_ms else None}


# ------------------------------------------------------------------------------ reference arm


def reference_runner():
  """(step function, kind): the reference's own CPU implementation of the path.  With the copy
  of twke18/SPML under baseline/_ref (scripts/install_reference.py; it travels to the GPU box
  with the snapshot) this IS the reference's code, driven as pyscripts/train/train.py:167-219
  drives it (kind 'reference'); without it, the oracle port, which issues the same ATen CPU
  kernels in the same order (kind 'port')."""
  ref_root = os.path.join(ROOT, 'baseline', '_ref')
  if os.path.isdir(os.path.join(ref_root, 'spml')):
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import reference_step
    return reference_step.make_runner(ref_root), 'reference'
  from oracle import spml_oracle as O

  def step(cfg, w, batch, bank):
    if w.variant != 'segsort' and w.name not in step.classifier:
      step.classifier[w.name] = O.make_classifier(cfg).eval()
    out = O.contrastive_step(cfg, batch, bank, variant=w.variant,
                             classifier=step.classifier.get(w.name))
    O.memory_bank_update(bank, {k: out[k] for k in out if k.startswith('prototype')},
                         w.memory_bank_size, w.batch)
    return out
  step.classifier = {}
  return step, 'port'


def time_cpu_reference(w, cfg, warm, steps, budget_s):
  """Times the reference's CPU path on a bounded sample of the workload, all host threads."""
  step_fn, kind = reference_runner()
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  batches = [synth.make_batch(w, seed=235, step=s) for s in range(4)]
  bank, times = {}, []
  budget_end = time.perf_counter() + budget_s
  for s in range(warm + steps):
    t0 = time.perf_counter()
    step_fn(cfg, w, batches[s % 4], bank)
    if s >= warm:
      times.append(time.perf_counter() - t0)
    if time.perf_counter() > budget_end and len(times) >= 3:
      break
  return {'value': w.batch * len(times) / sum(times), 'unit': 'images/s', 'cores': cores,
          'kind': kind,
          'sample': '%d steps of %s after %d warm-ups, fp32 torch %s CPU ops, %d threads'
                    % (len(times), w.name, warm, torch.__version__, cores),
          'ms_per_step': 1e3 * sum(times) / len(times)}


def run_reference(args):
  """The reference's CPU implementation of the path, all host threads.  Under torchrun rank 0
  alone runs ONE CPU process on its own minibatch; the other ranks exit without work."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  step_fn, kind = reference_runner()
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
    step_fn(cfg, w, batch, bank)
    dt = time.perf_counter() - t0
    if s >= args.warmup:
      times.append(dt)
  total = sum(times)
  value = w.batch * len(times) / total
  sample = '%d steps of workload %s (batch %d), fp32, torch %s CPU ops, ONE process' % (
      len(times), w.name, w.batch, torch.__version__)
  line = {
      'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'images/s',
      'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': workload_config(w, args.gpus),
      'note': 'one CPU process on rank 0 (batch %d) whatever --gpus says: the ratio against an '
              'N-GPU run compares N GPUs with one CPU process' % w.batch,
      'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': cores, 'kind': kind,
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


def algorithmic_work(key, sh):
  """(bytes, flops) of ONE call of a profiled entry point: SURVEY.md section 8d formulas with
  fp32 embeddings in HBM (b_e = 4).  `key` is 'entry' or 'entry:problem'."""
  name, _, prob = key.partition(':')
  d, dl, k, t = sh['d'], sh['dl'], sh['k'], sh['t']
  n = sh['n']
  if name == 'spml_kmeans':
    return ((t + 1) * n * dl * 4 + t * n * 4 + 2 * t * k * dl * 4, t * (2 * n * k * dl + n * dl))
  if name == 'spml_normalize_pack_fwd':
    return (sh['cap'] * d * 4 + n * (d + dl) * 4 + n * 32, 6 * n * dl)
  if name == 'spml_normalize_pack_bwd':
    return (n * (2 * d + 2 * dl) * 4 + sh['cap'] * d * 4, 8 * n * dl)
  if name in ('spml_segsort_fwd', 'spml_segsort_bwd'):
    rows, cols, dim = {'sem_ann': (sh['n_lab'], sh['m_lab'], d), 'sem_occ': (n, sh['m_all'], d),
                       'img_sim': (n, sh['m_cur'] / max(sh['batch'], 1), dl)}.get(
                           prob, (n, sh['m_all'], d))
    if name == 'spml_segsort_fwd':
      return (rows * (dim * 4 + 36) + cols * dim * 4, 2 * rows * cols * dim)
    # backward: S again, d(embedding), and d(prototypes) of the CURRENT step's columns only
    # (the memory-bank columns of sem_ann / sem_occ are detached: no gradient is computed)
    gcols = cols if prob == 'img_sim' else cols * sh['m_cur'] / max(sh['m_all'], 1)
    return (rows * (2 * dim * 4 + dim * 4 + 36) + (2 * cols + gcols) * dim * 4,
            (4 * cols + 2 * gcols) * rows * dim)
  if name == 'spml_segment_prototypes_fwd':
    return (n * d * 4 + n * 8 + 3 * sh['m_cur'] * d * 4, 2 * n * d)
  if name == 'spml_segment_prototypes_bwd':
    return (n * d * 4 + n * 8 + 2 * sh['m_cur'] * d * 4, 4 * n * d)
  if name == 'spml_topk_ranking':
    return (2 * sh['m_all'] * d * 4, 2 * sh['m_all'] * sh['m_all'] * d)
  return (None, None)


def roofline_of(key, sh, ms_per_call, launches_per_call, peaks):
  """Roofline of one profiled call: the bound is whichever of HBM / tensor time is larger."""
  nbytes, flops = algorithmic_work(key, sh)
  if nbytes is None:
    return None
  t_hbm = nbytes / (peaks['hbm_gbs'] * 1e9)
  tensor = key.startswith('spml_segsort')     # the only tensor-core kernels of the path
  t_tc = flops / (peaks['bf16_tflops'] * 1e12) if tensor else 0.0
  sec = ms_per_call * 1e-3
  if t_tc > t_hbm:
    achieved = flops / sec / 1e12
    out = {'bound': 'tensor', 'achieved': round(achieved, 3), 'peak': peaks['bf16_tflops'],
           'unit': 'TFLOP/s', 'frac': round(achieved / peaks['bf16_tflops'], 5)}
  else:
    achieved = nbytes / sec / 1e9
    out = {'bound': 'hbm', 'achieved': round(achieved, 2), 'peak': peaks['hbm_gbs'],
           'unit': 'GB/s', 'frac': round(achieved / peaks['hbm_gbs'], 5)}
  out.update({'kernel': key, 'traffic': None, 'peak_source': peaks['source'],
              'algorithmic_bytes_per_call': int(nbytes), 'algorithmic_flops_per_call': int(flops),
              'ms_per_call': round(ms_per_call, 4), 'launches_per_call': launches_per_call,
              'roofline_ms_per_call': round(1e3 * max(t_hbm, t_tc), 5)})
  return out


def run_b200(args):
  import torch.distributed as dist
  from spml_b200 import _lib, ops
  from spml_b200.head import ContrastiveHead
  from spml_b200.static_head import StaticContrastiveHead

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device (the product has no CPU path)')
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    # keep stdout for the one JSON line: NCCL prints its version (and any NCCL_DEBUG output)
    # on stdout when the communicator is created, so fd 1 points at stderr until then
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
      dist.init_process_group('nccl', device_id=dev)
      dist.barrier()
      torch.cuda.synchronize()
    finally:
      sys.stdout.flush()
      os.dup2(saved, 1)
      os.close(saved)
  _lib.load()

  w = synth.WORKLOADS[args.workload]
  cfg = synth.make_config(w)
  # THE PRODUCT PATH: the reference's operator API (generate_clusters ->
  # gather_clustering_and_update_prototypes -> Segsort*.forward -> backward -> memory bank),
  # data-dependent shapes, one host synchronisation per step
  head = ContrastiveHead(cfg, variant=w.variant).to(dev)
  # extra: the fixed-capacity CUDA-graph head (an API the reference does not have)
  static_ok = w.variant == 'segsort' and w.sem_occ_loss_types == 'segsort'
  static_head = StaticContrastiveHead(cfg, w.batch, w.height, w.width, w.loc_channels, device=dev,
                                      use_graph=not args.no_graph) if static_ok else None

  # every rank gets its own minibatches (different seeds): weak scaling over images
  host = [synth.make_batch(w, seed=235 + rank, step=s) for s in range(POOL)]
  host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
  resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
  flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
  keys = ('embedding', 'semantic_label', 'instance_label', 'semantic_tag', 'local_feature') + (
      ('semantic_label_full',) if 'semantic_label_full' in host[0] else ())
  h2d_bytes = sum(host[0][k].numel() * host[0][k].element_size() for k in keys)
  grad_host = torch.empty(host[0]['embedding'].shape, dtype=torch.float32).pin_memory()
  loss_host = torch.empty(4, dtype=torch.float32).pin_memory()

  def args_of(b):
    return (b['embedding'], b['semantic_label'], b['instance_label'], b['semantic_tag'],
            b['local_feature'])

  def run_head(b):
    emb = b['embedding'].detach().requires_grad_(True)
    out = head(emb, b['semantic_label'], b['instance_label'], b['semantic_tag'],
               b['local_feature'], b.get('semantic_label_full'))
    out['loss'].backward()
    head.update_memory_bank(1)                         # train.py:276-293
    return out, emb.grad

  def losses_of(out):
    zero = out['loss'].new_zeros(())
    # detached: a copy_ of a tensor with a grad_fn into the pinned buffer would chain every
    # step's autograd graph onto the buffer's history
    return torch.stack([out[k].detach() if out.get(k) is not None else zero
                        for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss', 'accuracy')])

  def step_resident(i):
    return run_head(resident[i % POOL])

  def step_e2e(i):
    # pinned host -> device inside the step; back: what train.py reads every step (the three
    # losses + accuracy, train.py:213-219); d(embedding) stays on the device for the backbone
    hb = host[i % POOL]
    out, _ = run_head({k: hb[k].to(dev, non_blocking=True) for k in keys})
    loss_host.copy_(losses_of(out), non_blocking=True)

  def step_e2e_grad(i):
    hb = host[i % POOL]
    out, grad = run_head({k: hb[k].to(dev, non_blocking=True) for k in keys})
    grad_host.copy_(grad, non_blocking=True)
    loss_host.copy_(losses_of(out), non_blocking=True)

  def step_static(i):
    return static_head.step(*args_of(resident[i % POOL]))

  def step_static_e2e(i):
    out = static_head.step(*args_of(host[i % POOL]))
    loss_host.copy_(torch.stack([out['sem_ann_loss'], out['sem_occ_loss'], out['img_sim_loss'],
                                 out['accuracy']]), non_blocking=True)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  def reset_banks():
    head.memory_banks.clear()
    if static_head is not None:
      static_head.reset_memory_bank()

  def timed(step_fn, steps, warmup, repeats):
    """`repeats` back-to-back regions of EXACTLY `steps` steps each, every step bracketed by
    CUDA events on the launching stream with an (untimed) L2 flush in front of it; returns the
    summed milliseconds of all steps * repeats steps, the library's kernel launches and the
    wall time."""
    reset_banks()
    for i in range(warmup):
      step_fn(i)
    barrier()
    pairs = []
    launches0 = _lib.launch_count()
    wall0 = time.perf_counter()
    for i in range(steps * repeats):
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

  warmup = max(args.warmup, 3)
  # how many K-step regions make ~1.2 s of load (so that the clock sampler sees it)
  probe_ms, _, _ = timed(step_resident, 5, warmup, 1)
  repeats = args.repeats or int(min(200, max(1, round(1200.0 / max(probe_ms / 5 * args.steps, 1e-3)))))

  def max_over_ranks(x):
    if world == 1:
      return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

  sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get('SPML_BENCH_NO_SAMPLER') else None
  if sampler:
    sampler.start()
    time.sleep(0.3)
  ms_res, launches, wall_res = timed(step_resident, args.steps, warmup, repeats)
  ms_e2e, _, wall_e2e = timed(step_e2e, args.steps, warmup, repeats)
  clocks = sampler.stop() if sampler else None
  short = max(1, repeats // 4)
  ms_e2e_g, _, _ = timed(step_e2e_grad, args.steps, warmup, short)
  ops.check_status(dev)
  static = None
  if static_head is not None:
    ms_s, launches_s, _ = timed(step_static, args.steps, warmup, short)
    ms_se, _, _ = timed(step_static_e2e, args.steps, warmup, short)
    static_head.check_overflow()         # a silent max_segments overflow would void the figure
    ms_s, ms_se = max_over_ranks(ms_s), max_over_ranks(ms_se)
    static = {'what': 'StaticContrastiveHead: fixed-capacity buffers, the whole step replayed as '
                      'one CUDA graph, no host synchronisation (not a reference API)',
              'ms_per_step': ms_s / (args.steps * short),
              'e2e_ms_per_step': ms_se / (args.steps * short),
              'kernels_per_step': static_head.kernels_per_step, 'cuda_graph': not args.no_graph}
  n_timed = args.steps * repeats
  ms_res, ms_e2e = max_over_ranks(ms_res), max_over_ranks(ms_e2e)
  ms_e2e_g = max_over_ranks(ms_e2e_g)
  images = w.batch * world * n_timed

  # ---- per-entry-point profile of a few steps (separate pass; events per C-ABI call of the
  # fine-grained entry points, which launch the same kernels as the stage-group calls)
  roofline, breakdown = None, None
  if rank == 0 and static_ok:
    prof_head = StaticContrastiveHead(cfg, w.batch, w.height, w.width, w.loc_channels,
                                      device=dev, use_graph=False)
    prof_head.collect_stats = True
    for i in range(3):
      prof_head.step(*args_of(resident[i % POOL]))
    torch.cuda.synchronize()
    _lib.PROFILE = []
    prof_steps = 10
    for i in range(prof_steps):
      flush.fill_(i)
      out = prof_head.step(*args_of(resident[(3 + i) % POOL]))
    torch.cuda.synchronize()
    records, _lib.PROFILE = _lib.PROFILE, None
    agg = {}
    for name, s, e, kernels in records:
      name = name.replace('spml_segsort_bwd_rows', 'spml_segsort_bwd')   # same kernels
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
    n_rows = int(out['num_pixels'])
    shapes = {'n': n_rows, 'cap': w.batch * w.height * w.width, 'batch': w.batch, 'd': w.dim,
              'dl': w.dim + w.loc_channels, 'm_cur': int(out['num_segments']),
              'm_all': int(out['num_live_prototypes']), 'm_lab': int(out['num_labelled_prototypes']),
              'n_lab': int(out['num_labelled_pixels']),
              'k': w.num_clusters[0] * w.num_clusters[1], 't': w.iterations}
    peaks = load_peaks()
    per_kernel = {}
    for key, a in agg.items():
      r = roofline_of(key, shapes, a['ms'] / a['calls'], a['kernels'] / a['calls'], peaks)
      if r is not None:
        per_kernel[key] = r
    top = next(k for k in breakdown if k in per_kernel)
    roofline = dict(per_kernel[top])
    # DRAM bytes per launch of that kernel from the committed ncu --set full capture (a profiler
    # number cannot be taken inside a timed run); null when this workload was not captured
    for tname in ('r2b_traffic.json', 'r2_traffic.json', 'r1c_traffic.json'):
      tpath = os.path.join(ROOT, 'profiles', tname)
      if os.path.exists(tpath):
        traffic = json.load(open(tpath))
        roofline['traffic'] = traffic.get(w.name, {}).get(top)
        roofline['traffic_source'] = traffic['source']
        break
    roofline['shapes'] = shapes
    roofline['all'] = {k: {'bound': v['bound'], 'frac': v['frac'], 'ms_per_call': v['ms_per_call'],
                           'roofline_ms_per_call': v['roofline_ms_per_call']}
                       for k, v in per_kernel.items()}

  # ---- CPU baseline on this box's host cores, rank 0, N == 1 only
  cpu_baseline = None
  if rank == 0 and world == 1 and not args.no_cpu_baseline:
    cpu_baseline = time_cpu_reference(w, cfg, warm=2, steps=12, budget_s=20.0)

  if rank == 0:
    value = images / (ms_res * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': warmup, 'repeats': repeats, 'timed_steps': n_timed,
        'ms_per_step': ms_res / n_timed, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(w, world),
        'api': 'the reference operator API (spml_b200.install symbols composed as train.py:167-293 '
               'does): data-dependent shapes, one host synchronisation per step',
        'contrastive_step_ms': ms_res / n_timed,
        'e2e': {'value': images / (ms_e2e * 1e-3), 'unit': 'images/s',
                'ms_per_step': ms_e2e / n_timed, 'h2d_bytes_per_step': h2d_bytes,
                'd2h_bytes_per_step': loss_host.numel() * 4,
                'd2h': 'three losses + accuracy (what train.py:213-219 reads); d(embedding) '
                       'stays on the device for the backbone',
                # secondary: the gradient copied out as well
                'with_gradient_copy': {
                    'value': w.batch * world * args.steps * short / (ms_e2e_g * 1e-3),
                    'ms_per_step': ms_e2e_g / (args.steps * short),
                    'd2h_bytes_per_step': grad_host.numel() * 4 + loss_host.numel() * 4}},
        'gpu_launches': launches,
        'gpu_launches_per_step': launches / n_timed,
        'static_head': static,
        'wall_s': {'resident': round(wall_res, 4), 'e2e': round(wall_e2e, 4)},
        'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu_baseline,
        'breakdown': breakdown,
    }
  else:
    line = None
  if world > 1:
    dist.destroy_process_group()
  return line


def training_step_arm(args):
  """BASELINE.json's metric has two clauses; the line above is the second (contrastive-loss
  step).  The first - training images/s with the ResNet-101 backbone and ONE NCCL all-reduce of
  the gradients at 1/2/4/8 GPUs - is the `train_voc_b4` workload of this script; it is run here
  as well (every rank starts the same command on its own rendezvous port, after this process has
  left its process group) so that the default invocation reports both.  Never fails the line."""
  import subprocess
  # (torchrun's agent serves the rendezvous store of THIS process group only: without its
  # variables rank 0 of the second command hosts its own store on the other port)
  env = {k: v for k, v in os.environ.items() if not k.startswith('TORCHELASTIC')}
  if 'MASTER_PORT' in env:
    env['MASTER_PORT'] = str(1024 + (int(env['MASTER_PORT']) + 4321 - 1024) % 60000)
  cmd = [sys.executable, os.path.abspath(__file__), '--workload', 'train_voc_b4', '--gpus',
         str(args.gpus), '--steps', '10', '--warmup', '3']
  try:
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=150)
    rows = [ln for ln in out.stdout.splitlines() if ln.startswith('{')]
    if not rows:
      return {'error': (out.stderr or 'no output')[-300:]}
    t = json.loads(rows[-1])
    return {'metric': t['metric'], 'value': t['value'], 'unit': t['unit'], 'n_gpus': t['n_gpus'],
            'ms_per_step': t['ms_per_step'], 'steps': t['steps'], 'warmup': t['warmup'],
            'scaling': t['scaling'], 'dtype': t['dtype'], 'config': t['config'],
            'phase_ms': t['phase_ms'], 'all_reduce': t['all_reduce'],
            'contrastive_head_ms': t['contrastive_head_ms'], 'clocks': t['clocks'],
            'command': 'bench.py --workload train_voc_b4 --gpus %d --steps 10 --warmup 3' % args.gpus}
  except Exception as e:   # noqa: BLE001 (a secondary figure must not cost the headline)
    return {'error': repr(e)[-300:]}


# ------------------------------------------------------------------------------ training step

TRAIN_METRIC = 'training images/sec (512x512 crop, ResNet-101 DeepLab backbone + contrastive head)'


def run_train(args):
  """BASELINE.json metric, first clause: the whole training iteration of
  pyscripts/train/train.py:154-293 on synthetic 512x512 images: ResNet-101 DeepLab backbone
  (PyTorch / cuDNN, channels-last, bf16 autocast) -> contrastive head (libspml_b200, fp32) ->
  backward -> ONE bucketed NCCL all-reduce of the 47.3 M backbone gradients (DDP, overlapped
  with the backbone's backward) -> SGD step -> memory-bank update.  One process per GPU, the
  minibatch sharded over the ranks (weak scaling: 4 images per GPU as shipped)."""
  import torch.distributed as dist
  from spml_b200 import _lib, ops
  from spml_b200.backbone import ResnetDeeplabEmbedding, num_parameters
  from spml_b200.head import ContrastiveHead

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  dev = torch.device('cuda', local_rank)
  if world > 1:
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
      dist.init_process_group('nccl', device_id=dev)
      dist.barrier()
      torch.cuda.synchronize()
    finally:
      sys.stdout.flush()
      os.dup2(saved, 1)
      os.close(saved)
  _lib.load()
  w = synth.WORKLOADS[args.workload]
  cfg = synth.make_config(w)
  torch.manual_seed(235)
  bf16 = args.backbone_dtype == 'bf16'
  net = ResnetDeeplabEmbedding(w.dim).to(dev).to(memory_format=torch.channels_last)
  n_params = num_parameters(net)
  size = 4 * w.height      # 512: the embedding map is a quarter of the crop
  graphed = not args.no_graph
  if graphed:
    # The backbone alone is ~1 000 cuDNN / ATen launches per step, i.e. launch-bound on the host
    # (34 ms eager vs ~15 ms of GPU work at batch 4): its forward and backward are captured
    # as two CUDA graphs (torch.cuda.make_graphed_callables); the head, the all-reduce and
    # the optimizer stay outside the graphs.
    sample = torch.randn(w.batch, 3, size, size, device=dev).contiguous(
        memory_format=torch.channels_last)
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bf16, cache_enabled=False):
      net = torch.cuda.make_graphed_callables(net, (sample,), num_warmup_iters=3)
    del sample
  model = net
  if world > 1:
    # bucket_cap_mb above the 190 MB of fp32 gradients: ONE NCCL all-reduce per step
    model = torch.nn.parallel.DistributedDataParallel(
        net, device_ids=[local_rank], gradient_as_bucket_view=True, broadcast_buffers=False,
        bucket_cap_mb=256)
  opt = torch.optim.SGD(net.parameters(), lr=3e-3, momentum=0.9, weight_decay=5e-4, fused=True)
  head = ContrastiveHead(cfg, variant=w.variant).to(dev)
  gen = torch.Generator().manual_seed(1000 + rank)
  host = []
  for s in range(4):
    b = synth.make_batch(w, seed=235 + rank, step=s)
    b['image'] = torch.randn(w.batch, 3, size, size, generator=gen)
    host.append({k: v.pin_memory() for k, v in b.items() if k != 'embedding'})
  keys = ('image', 'semantic_label', 'instance_label', 'semantic_tag', 'local_feature')
  h2d_bytes = sum(host[0][k].numel() * host[0][k].element_size() for k in keys)
  loss_host = torch.empty(4, dtype=torch.float32).pin_memory()

  phases = []     # per step: CUDA events at the phase boundaries (cheap; in every timed step)

  def step(i, mark=None):
    mark = mark if mark is not None else (lambda: None)
    mark()
    b = {k: host[i % 4][k].to(dev, non_blocking=True) for k in keys}
    images = b['image'].contiguous(memory_format=torch.channels_last)
    mark()
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=bf16, cache_enabled=False):
      emb = model(images)
    mark()
    out = head(emb.float(), b['semantic_label'], b['instance_label'], b['semantic_tag'],
               b['local_feature'])
    mark()
    opt.zero_grad(set_to_none=True)
    out['loss'].backward()          # DDP all-reduces the gradient bucket during this call
    mark()
    opt.step()
    head.update_memory_bank(world)
    zero = out['loss'].new_zeros(())
    loss_host.copy_(torch.stack([out[k].detach() if out.get(k) is not None else zero for k in
                                 ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss', 'accuracy')]),
                    non_blocking=True)
    mark()

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  warmup = max(args.warmup, 3)
  # the sampler starts BEFORE the warm-up and the barrier: anything rank 0 does alone between
  # the barrier and its first timed step is waiting time of the other ranks in the first
  # all-reduce (a 0.3 s sleep here once showed up as +15 ms per step of a 20-step run)
  sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get('SPML_BENCH_NO_SAMPLER') else None
  if sampler:
    sampler.start()
  for i in range(warmup):
    step(i)
  barrier()
  launches0 = _lib.launch_count()
  pairs = []
  for i in range(args.steps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = []

    def mark():
      ev = torch.cuda.Event(enable_timing=True)
      ev.record()
      marks.append(ev)
    s.record()
    step(warmup + i, mark)
    e.record()
    pairs.append((s, e))
    phases.append(marks)
  barrier()
  clocks = sampler.stop() if sampler else None
  launches = _lib.launch_count() - launches0
  ms = sum(s.elapsed_time(e) for s, e in pairs)
  names = ('h2d', 'backbone_forward', 'head_forward', 'backward_and_all_reduce', 'sgd_and_bank')
  phase_ms = {n: sum(m[j].elapsed_time(m[j + 1]) for m in phases) / len(phases)
              for j, n in enumerate(names)}
  ops.check_status(dev)

  # the gradient all-reduce alone (what DDP overlaps with the backward): flat fp32 buffer
  allreduce, per_rank = None, None
  if world > 1:
    flat = torch.zeros(n_params, dtype=torch.float32, device=dev)
    for _ in range(3):
      dist.all_reduce(flat)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
      dist.all_reduce(flat)
    e.record()
    torch.cuda.synchronize()
    ar_ms = s.elapsed_time(e) / 10
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {'ms_per_step': ms / args.steps, 'phase_ms': phase_ms})
    t = torch.tensor([ms, ar_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ar_ms = float(t[0]), float(t[1])
    gb = n_params * 4 / 1e9
    allreduce = {'bytes': n_params * 4, 'ms': ar_ms, 'algbw_GBps': gb / (ar_ms * 1e-3),
                 'busbw_GBps': gb / (ar_ms * 1e-3) * 2 * (world - 1) / world,
                 'what': 'one NCCL all-reduce of the flat fp32 gradient, timed alone (the step issues '
                         'the same all-reduce once, after the graphed backbone backward)'}
  # the head's share of the step (same inputs, backbone output detached)
  head_ms = None
  if rank == 0:
    b = {k: host[0][k].to(dev) for k in keys}
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=bf16):
      emb0 = net(b['image'].contiguous(memory_format=torch.channels_last)).float()
    bank = dict(head.memory_banks)
    times = []
    for _ in range(8):
      head.memory_banks = dict(bank)
      e_in = emb0.clone().requires_grad_(True)
      s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      s.record()
      o = head(e_in, b['semantic_label'], b['instance_label'], b['semantic_tag'],
               b['local_feature'])
      o['loss'].backward()
      e.record()
      torch.cuda.synchronize()
      times.append(s.elapsed_time(e))
    head_ms = sorted(times)[len(times) // 2]
  if rank == 0:
    images = w.batch * world * args.steps
    value = images / (ms * 1e-3)
    line = {
        'metric': TRAIN_METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16 backbone (autocast), f32 contrastive head' if bf16 else 'f32',
        'data': 'synthetic',
        'config': {'workload': w.name, 'model': 'ResNet-101 DeepLab (47.3 M parameters, random '
                   'init) + SPML contrastive head', 'images_per_gpu': w.batch,
                   'global_batch': w.batch * world, 'crop': [size, size],
                   'embedding_map': [w.height, w.width], 'embedding_dim': w.dim,
                   'kmeans_seeds': list(w.num_clusters), 'kmeans_iterations': w.iterations,
                   'memory_bank_steps': w.memory_bank_size,
                   'parallelism': 'dp%d: one process per GPU, ONE NCCL all-reduce of the backbone '
                                  'gradients per step (DDP, single 190 MB bucket), rank-local '
                                  'prototypes' % world,
                   'backbone_cuda_graphs': graphed,
                   'l2': 'inputs + activations of a step exceed the 126 MB L2'},
        'e2e': {'value': value, 'unit': 'images/s', 'ms_per_step': ms / args.steps,
                'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 16,
                'note': 'the timed step itself copies its images + labels from pinned host '
                        'memory and reads the losses back'},
        'gpu_launches': launches, 'gpu_launches_per_step': launches / args.steps,
        'contrastive_head_ms': head_ms, 'phase_ms': phase_ms, 'all_reduce': allreduce,
        'per_rank': per_rank if world > 1 else None,
        'parameters': n_params,
        'clocks': clocks,
    }
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def main():
  args = parse_args()
  if args.workload.startswith('train_'):
    if args.impl == 'reference':
      print(json.dumps({'impl': 'reference', 'unavailable':
                        'the reference CPU training step (ResNet-101 at 512x512) takes minutes '
                        'per step; the reference arm covers the contrastive-loss step workloads'}))
      return
    run_train(args)
  elif args.impl == 'reference':
    run_reference(args)
  else:
    line = run_b200(args)
    train = None
    if not args.no_train_arm and args.workload == 'voc_scribble_b1':
      train = training_step_arm(args)       # every rank takes part; rank 0 keeps the result
    if line is not None:
      if train is not None:
        line['training_step'] = train
      print(json.dumps(line))


if __name__ == '__main__':
  main()
