"""The contrastive head with FIXED-CAPACITY buffers: no host synchronisation anywhere in
the step, so forward + backward + memory-bank update replay as one CUDA graph.

`ContrastiveHead` (head.py) follows the reference's data-dependent shapes: the number of
kept pixels N and of segments M size the tensors, which costs one device->host read-back
per step and keeps the CPU in the launch path (~70 launches).  Here every tensor has its
capacity shape (N <= batch*H*W rows, M <= max_segments), the live counts stay on the
device (kernels take them as pointers), dead prototype columns are masked out by the
operand pre-pass, and the memory bank is a fixed ring.  The arithmetic, the column
order of the prototype bank ([current, oldest bank step, ..., newest]) and therefore the
results are those of ContrastiveHead / the reference.

    head = StaticContrastiveHead(config, batch, height, width)
    out = head.step(embedding, semantic_label, instance_label, semantic_tag, local_feature)
    embedding.backward(out['grad_embedding'])       # d loss / d embedding, already computed

`step` copies the inputs into the graph's static buffers and replays it; `out` holds
views of static output buffers (valid until the next step).
"""

from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from . import ops
from . import segsort_common


def _round_up(x, m):
  return (x + m - 1) // m * m


class StaticContrastiveHead(nn.Module):

  def __init__(self, config, batch, height, width, loc_channels=2, max_segments=None,
               device=None, use_graph=True):
    super(StaticContrastiveHead, self).__init__()
    self.config = config
    t = config.train
    for name in ('sem_ann', 'sem_occ', 'img_sim'):
      if getattr(t, name + '_loss_types') != 'segsort':
        raise NotImplementedError('StaticContrastiveHead runs the three segsort losses together')
    self.B, self.H, self.W = int(batch), int(height), int(width)
    self.D = int(config.network.embedding_dim)
    self.L = int(loc_channels)
    self.C = int(config.dataset.num_classes)
    self.div = int(config.network.label_divisor)
    self.ignore = int(config.dataset.semantic_ignore_index)
    self.num_clusters = list(config.network.kmeans_num_clusters)
    self.iterations = int(config.network.kmeans_iterations)
    self.device = torch.device(device) if device is not None else torch.device(
        'cuda', torch.cuda.current_device())
    k = self.num_clusters[0] * self.num_clusters[1]
    # capacity for the segments of one step: every (cluster, label) pair that can occur is
    # bounded by the pixels; 16 labels per cluster is generous for over-segmented scribbles
    self.m_cap = int(max_segments) if max_segments else _round_up(self.B * k * 16, 128)
    self.bank_size = int(getattr(t, 'memory_bank_size', 0))
    self.use_graph = bool(use_graph)
    self._graph = None
    self.collect_stats = False
    self.kernels_per_step = 0
    self._alloc_static()

  # ------------------------------------------------------------------------------ buffers
  def _alloc_static(self):
    B, H, W, D, L, dev = self.B, self.H, self.W, self.D, self.L, self.device
    z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=dev)
    self.in_emb = z(B, D, H, W).requires_grad_(True)
    self.in_sem = z(B, H, W, dtype=torch.int64)
    self.in_inst = z(B, H, W, dtype=torch.int64)
    self.in_tags = z(B, 256, dtype=torch.int64)
    self.in_loc = z(B, H, W, L)
    self.overflow = z(1, dtype=torch.int32)
    S, M = max(self.bank_size, 1), self.m_cap
    self.bank_p = z(S, M, D)
    self.bank_sem = torch.full((S, M), self.C, dtype=torch.int64, device=dev)
    self.bank_mask = z(S, M, dtype=torch.int64)
    self.bank_live = z(S, M, dtype=torch.uint8)
    self.out = {}
    # 0, 1, 3: the three losses (high priority: the backward pass waits for them);
    # 2: top-k, 4: memory-bank FIFO (outputs only)
    # (measured on one box: 1.490 -> 1.451 ms at batch 4 with the priorities)
    self._side_streams = [torch.cuda.Stream(device=dev, priority=0 if i in (2, 4) else -1)
                          for i in range(5)]

  # ------------------------------------------------------------------------------ one step
  def _forward(self):
    B, H, W, C, dev = self.B, self.H, self.W, self.C, self.device
    cap = B * H * W
    t = self.config.train
    sem, inst = self.in_sem, self.in_inst
    # resnet_deeplab.py:112-117 gives the dropped pixels the label `labels.max() + 1`; the value
    # never leaves segment_by_kmeans (those pixels are removed, :355-365), so any label that
    # cannot occur does: a constant saves the max-reduction in front of the clustering
    ignore = 1 << 62
    labels = torch.add(inst, sem, alpha=self.div).masked_fill_(sem == self.ignore, ignore)
    main = torch.cuda.current_stream(dev)
    streams = self._side_streams
    side = {}

    def tag_masks(bid):
      # image-tag masks per pixel (segsort.py:146-150): they only need the batch ids, so they
      # run on a side stream next to the clustering kernel
      ev = torch.cuda.Event()
      ev.record(main)
      with torch.cuda.stream(streams[0]):
        streams[0].wait_event(ev)
        side['img_masks'] = ops.pack_tags(self.in_tags[:, 1:C])
        side['pix_mask'] = torch.index_select(side['img_masks'], 0, bid)

    e, el, lab, cid, bid, img_off, count, _ = segsort_common.segment_core(
        self.in_emb, labels, self.num_clusters, None, self.in_loc, ignore, self.iterations,
        batch_index_offset=0, after_pack=tag_masks)
    n_dev = img_off[B:B + 1]
    sem_pix, inst_pix, keep, p_sem, p_inst, p_bid, p_live = ops.segment_labels(
        lab, bid, cid, n_dev, self.div, C, self.m_cap, C, self.overflow)
    # the two prototype sets (models/utils.py:113-116) are independent: one per stream
    ev = torch.cuda.Event()
    ev.record(main)
    with torch.cuda.stream(streams[1]):
      streams[1].wait_event(ev)
      protos_loc = ops.SegmentPrototypes.apply(el, cid, self.m_cap, n_dev)
    protos = ops.SegmentPrototypes.apply(e, cid, self.m_cap, n_dev)
    main.wait_stream(streams[0])
    main.wait_stream(streams[1])
    img_masks, pix_mask = side['img_masks'], side['pix_mask']
    cur_mask = torch.index_select(img_masks, 0, p_bid.clamp(min=0))
    use_bank = self.bank_size > 0
    if use_bank:                                                             # :153-182
      p_all = torch.cat([protos, self.bank_p.view(-1, self.D)], 0)
      psem_all = torch.cat([p_sem, self.bank_sem.view(-1)], 0)
      pmask_all = torch.cat([cur_mask, self.bank_mask.view(-1)], 0)
      plive_all = torch.cat([p_live, self.bank_live.view(-1)], 0)
    else:
      p_all, psem_all, pmask_all, plive_all = protos, p_sem, cur_mask, p_live
    if use_bank:
      # train.py:276-293.  The concatenations above copied the bank, so the FIFO can advance on
      # its own stream while the losses and their backward run.
      ev = torch.cuda.Event()
      ev.record(main)
      with torch.cuda.stream(streams[4]), torch.no_grad():
        streams[4].wait_event(ev)
        for buf, new in ((self.bank_p, protos), (self.bank_sem, p_sem),
                         (self.bank_mask, cur_mask), (self.bank_live, p_live)):
          if self.bank_size > 1:
            buf[:-1] = buf[1:].clone()
          buf[-1] = new.detach()

    # The three losses and the accuracy are independent: each runs on its own stream so
    # that, inside the CUDA graph, they are parallel branches (at batch 1 a single loss only
    # fills ~117 of 148 SMs for a few microseconds).  Autograd replays each backward on the
    # stream of its forward, so the backward passes overlap too.
    fork = torch.cuda.Event()
    fork.record(main)
    for st in streams[:4]:
      st.wait_event(fork)

    with torch.cuda.stream(streams[0]):
      # sem_ann (:184-201): labelled pixels x labelled live prototypes
      _, rows, off = ops.valid_scan(keep.view(1, cap), 0, 1, cap, want_src=True)
      problem = ops.SegsortProblem(
          sem_pix, cid, psem_all, t.sem_ann_concentration, _lib.MODE_CLASS, row_index=rows,
          group_off=off, num_groups=1, n_rows=cap, max_rows_per_group=cap,
          proto_valid=plive_all & (psem_all < C).to(torch.uint8), name='sem_ann',
          proto_grad_rows=self.m_cap)     # the bank rows behind are detached
      sem_ann = ops.SegsortLossFn.apply(e, p_all, problem) * t.sem_ann_loss_weight
    with torch.cuda.stream(streams[1]):
      # sem_occ: all live pixels x all live prototypes, image-tag masks
      all_rows = torch.stack([img_off[0], img_off[B]])
      problem = ops.SegsortProblem(
          pix_mask, cid, pmask_all, t.sem_occ_concentration, _lib.MODE_TAGS, group_off=all_rows,
          num_groups=1, n_rows=cap, max_rows_per_group=cap, proto_valid=plive_all, name='sem_occ',
          proto_grad_rows=self.m_cap)
      sem_occ = ops.SegsortLossFn.apply(e, p_all, problem) * t.sem_occ_loss_weight
    with torch.cuda.stream(streams[2]):
      acc, _ = ops.topk_ranking(p_all.detach(), psem_all, p_all.detach(), psem_all, 5,
                                qvalid=plive_all, pvalid=plive_all)           # :212-217
    with torch.cuda.stream(streams[3]):
      # img_sim (:220-240): rows grouped by image, each image sees its own prototypes
      per_image = torch.zeros(B + 1, dtype=torch.int32, device=dev)
      per_image.scatter_add_(0, torch.where(p_live.bool(), p_bid + 1, torch.zeros_like(p_bid)),
                             p_live.to(torch.int32))
      col_off = torch.cumsum(per_image, 0, dtype=torch.int32)
      problem = ops.SegsortProblem(
          inst_pix, cid, p_inst, t.img_sim_concentration, _lib.MODE_CLASS,
          reduction=_lib.REDUCE_GROUP_MEAN, group_off=img_off, col_off=col_off, num_groups=B,
          n_rows=cap, max_rows_per_group=H * W, name='img_sim')
      img_sim = ops.SegsortLossFn.apply(el, protos_loc, problem) * t.img_sim_loss_weight
    for st in (streams[0], streams[1], streams[3]):
      main.wait_stream(st)

    loss = sem_ann + sem_occ + img_sim                                       # train.py:213-219
    loss.backward()
    main.wait_stream(streams[2])     # the retrieval accuracy is an output only: it may run on
    main.wait_stream(streams[4])     # next to the backward pass, like the memory-bank FIFO
    extra = {}
    if self.collect_stats:     # problem sizes for bench.py's roofline (adds small reductions)
      extra = {'num_labelled_pixels': off[1], 'num_live_prototypes': plive_all.sum(),
               'num_labelled_prototypes': (plive_all & (psem_all < C).to(torch.uint8)).sum()}
    return {**extra, 'sem_ann_loss': sem_ann.detach(), 'sem_occ_loss': sem_occ.detach(),
            'img_sim_loss': img_sim.detach(), 'accuracy': acc, 'loss': loss.detach(),
            'num_pixels': n_dev, 'num_segments': count, 'overflow': self.overflow,
            'cluster_index': cid, 'cluster_batch_index': bid, 'prototype': protos.detach(),
            'prototype_semantic_label': p_sem, 'prototype_live': p_live}

  def _capture(self):
    # warm-up outside the capture: lazy module loading, cudaFuncSetAttribute, allocator growth
    bank = [b.clone() for b in (self.bank_p, self.bank_sem, self.bank_mask, self.bank_live)]
    side = torch.cuda.Stream(device=self.device)
    side.wait_stream(torch.cuda.current_stream(self.device))
    with torch.cuda.stream(side):
      for _ in range(2):
        self.in_emb.grad = None
        self._forward()
    torch.cuda.current_stream(self.device).wait_stream(side)
    for dst, src in zip((self.bank_p, self.bank_sem, self.bank_mask, self.bank_live), bank):
      dst.copy_(src)
    self.overflow.zero_()
    self.in_emb.grad = None
    graph = torch.cuda.CUDAGraph()
    launches = _lib.launch_count()
    with torch.cuda.graph(graph):
      self.out = self._forward()
    self.kernels_per_step = _lib.launch_count() - launches    # kernels of this library per replay
    self.out['grad_embedding'] = self.in_emb.grad
    for dst, src in zip((self.bank_p, self.bank_sem, self.bank_mask, self.bank_live), bank):
      dst.copy_(src)
    self.overflow.zero_()
    self._graph = graph

  def load_inputs(self, embedding, semantic_label, instance_label, semantic_tag,
                  local_feature=None):
    with torch.no_grad():
      self.in_emb.copy_(embedding, non_blocking=True)
      self.in_sem.copy_(semantic_label, non_blocking=True)
      self.in_inst.copy_(instance_label, non_blocking=True)
      self.in_tags.copy_(semantic_tag, non_blocking=True)
      if local_feature is not None:
        self.in_loc.copy_(local_feature, non_blocking=True)
      else:
        self.in_loc.copy_(segsort_common._default_location(self.H, self.W, self.device))

  def step(self, embedding, semantic_label, instance_label, semantic_tag, local_feature=None):
    """Loads one minibatch (device or pinned-host tensors) and runs forward + backward +
    memory-bank update.  Returns a dict of device tensors, including 'grad_embedding'."""
    self.load_inputs(embedding, semantic_label, instance_label, semantic_tag, local_feature)
    if not self.use_graph:
      self.in_emb.grad = None
      self.out = self._forward()
      self.out['grad_embedding'] = self.in_emb.grad
      return self.out
    if self._graph is None:
      self._capture()
    self._graph.replay()
    return self.out

  def reset_memory_bank(self):
    self.bank_p.zero_()
    self.bank_sem.fill_(self.C)
    self.bank_mask.zero_()
    self.bank_live.zero_()

  def check_overflow(self):
    """Host sync: raises if a step produced more segments than `max_segments`."""
    if int(self.overflow) != 0:
      raise RuntimeError('StaticContrastiveHead: more than %d segments in a step; raise '
                         'max_segments' % self.m_cap)
