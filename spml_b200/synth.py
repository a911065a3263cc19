"""Seeded synthetic inputs for the pixel-to-segment contrastive path.

Follows the recipe of SURVEY.md section 8(d): per image a Voronoi partition of
the H' x W' embedding map into `num_regions` ground-truth regions, each region
with one class out of a small per-image class set; the over-segmentation
("instance") label is the region id; the semantic label is the class on a
labelled fraction `rho` of pixels, `num_classes` (unlabelled but participating)
elsewhere and 255 on a border (crop padding, dropped before clustering).
Embeddings are `centre[region] + sigma * N(0, 1)`.

Everything is generated on the CPU with a private torch.Generator so that the
same seed gives the same tensors in this container, on the GPU box and in the
golden-vector script.  The default seed 235 is the reference's own seed
(pyscripts/train/train.py:34-35).
"""

from __future__ import annotations

import dataclasses
import math
from types import SimpleNamespace

import torch


@dataclasses.dataclass
class Workload:
  """Shape and label statistics of one synthetic minibatch."""
  name: str = 'voc_scribble_b1'
  batch: int = 1
  height: int = 128            # embedding-map height (512 crop / 4)
  width: int = 128
  dim: int = 64                # network.embedding_dim
  num_clusters: tuple = (6, 6)  # kmeans_num_clusters
  iterations: int = 10         # kmeans_iterations
  num_classes: int = 21
  label_divisor: int = 2048
  ignore_index: int = 255
  num_regions: int = 64
  classes_per_image: int = 4
  rho: float = 0.05            # labelled pixel fraction (scribble)
  border: float = 0.05         # fraction of the map side that is padding
  sigma: float = 1.0
  memory_bank_size: int = 2
  sem_ann_concentration: float = 6.0
  sem_occ_concentration: float = 12.0
  img_sim_concentration: float = 16.0
  sem_ann_loss_weight: float = 1.0
  sem_occ_loss_weight: float = 0.5
  img_sim_loss_weight: float = 0.1
  loc_channels: int = 2        # location (y, x); DensePose adds RGB -> 5
  # which head: 'segsort' (predictions/segsort.py), 'softmax' (segsort_softmax.py, the class
  # train.py instantiates) or 'densepose' (resnet_pspnet_densepose.py generate_clusters +
  # segsort_softmax_densepose.py)
  variant: str = 'segsort'
  sem_occ_loss_types: str = 'segsort'
  label_upsample: int = 1      # > 1: also emit the full-resolution label map (classifier CE)


WORKLOADS = {
    # BASELINE.json configs[0]/[1]: VOC12 scribble, batch 1, 512x512 crop.
    'voc_scribble_b1': Workload(),
    # bashscripts/voc12/train_spml_scribble.sh:28 per-GPU batch of 4.
    'voc_scribble_b4': Workload(name='voc_scribble_b4', batch=4),
    # BASELINE.json configs[2]: image tags (CAM), 2 img/GPU, 8x8 seeds.
    'voc_tag_b2': Workload(name='voc_tag_b2', batch=2, num_clusters=(8, 8),
                           rho=0.6, sem_occ_concentration=8.0),
    # BASELINE.json configs[3]: DensePose points, 769 crop -> 194x194, D=32, 128 seeds,
    # location + RGB features, the DensePose head (1-NN tag propagation over
    # prototype_with_loc, img_sim on the plain embeddings), no memory bank, sem_occ off
    # (bashscripts/densepose/train_spml_point.sh:14-44).
    'densepose_b1': Workload(name='densepose_b1', batch=1, height=194,
                             width=194, dim=32, num_clusters=(8, 16),
                             num_classes=15, rho=0.02, memory_bank_size=0,
                             sem_occ_concentration=8.0, loc_channels=5,
                             variant='densepose', sem_occ_loss_types='none'),
    # the same shapes through the VOC head (round-1 workload, kept for comparison)
    'densepose_shape_voc_head_b1': Workload(
        name='densepose_shape_voc_head_b1', batch=1, height=194, width=194, dim=32,
        num_clusters=(8, 16), num_classes=15, rho=0.02, memory_bank_size=0,
        sem_occ_concentration=8.0, loc_channels=5),
    # the softmax head train.py instantiates (segsort_softmax.py) at the VOC shapes
    'voc_scribble_softmax_b1': Workload(name='voc_scribble_softmax_b1', variant='softmax',
                                        label_upsample=4),
    # BASELINE.json's first metric clause: the whole training step (ResNet-101 DeepLab backbone
    # on cuDNN -> this head -> backward -> gradient all-reduce -> SGD) at the shipped per-GPU
    # batch of 4 (bashscripts/voc12/train_spml_scribble.sh:28) and at batch 1; bench.py only.
    'train_voc_b4': Workload(name='train_voc_b4', batch=4),
    'train_voc_b1': Workload(name='train_voc_b1', batch=1),
    # small cases used by the parity tests / golden vectors.
    'tiny': Workload(name='tiny', batch=2, height=24, width=20, dim=16,
                     num_clusters=(3, 3), num_regions=9, iterations=10,
                     rho=0.3, memory_bank_size=1),
    'small': Workload(name='small', batch=2, height=40, width=32, dim=24,
                      num_clusters=(4, 4), num_regions=16, rho=0.2,
                      memory_bank_size=2),
    'tiny_softmax': Workload(name='tiny_softmax', batch=2, height=24, width=20, dim=16,
                             num_clusters=(3, 3), num_regions=9, rho=0.3, memory_bank_size=1,
                             variant='softmax', label_upsample=2),
    # DensePose head with every branch on: sem_occ enabled and a memory bank
    'tiny_densepose': Workload(name='tiny_densepose', batch=2, height=24, width=20, dim=16,
                               num_clusters=(3, 3), num_regions=9, rho=0.3, memory_bank_size=1,
                               num_classes=15, loc_channels=5, variant='densepose',
                               sem_occ_concentration=8.0),
    # ... and as shipped (sem_occ off, no bank)
    'tiny_densepose_shipped': Workload(
        name='tiny_densepose_shipped', batch=2, height=24, width=20, dim=16,
        num_clusters=(3, 3), num_regions=9, rho=0.3, memory_bank_size=0, num_classes=15,
        loc_channels=5, variant='densepose', sem_occ_loss_types='none'),
}


def make_config(w: Workload) -> SimpleNamespace:
  """The attribute tree `segsort(config)` reads (segsort_softmax.py:41-71)."""
  return SimpleNamespace(
      dataset=SimpleNamespace(num_classes=w.num_classes,
                              semantic_ignore_index=w.ignore_index),
      network=SimpleNamespace(embedding_dim=w.dim,
                              label_divisor=w.label_divisor,
                              kmeans_iterations=w.iterations,
                              kmeans_num_clusters=list(w.num_clusters)),
      train=SimpleNamespace(
          sem_ann_loss_types='segsort', sem_occ_loss_types=w.sem_occ_loss_types,
          img_sim_loss_types='segsort', feat_aff_loss_types='none',
          sem_ann_concentration=w.sem_ann_concentration,
          sem_occ_concentration=w.sem_occ_concentration,
          img_sim_concentration=w.img_sim_concentration,
          feat_aff_concentration=0,
          sem_ann_loss_weight=w.sem_ann_loss_weight,
          sem_occ_loss_weight=w.sem_occ_loss_weight,
          img_sim_loss_weight=w.img_sim_loss_weight,
          feat_aff_loss_weight=0.0,
          memory_bank_size=w.memory_bank_size))


def location_features(height: int, width: int) -> torch.Tensor:
  """[H, W, 2] (y, x) in [-0.5, 0.5]: what LocationColorNetwork feeds k-means
  (local_model.py:89-92 on top of segsort/common.py:156-189)."""
  ys = torch.linspace(0, 1, height)
  xs = torch.linspace(0, 1, width)
  yy, xx = torch.meshgrid(ys, xs, indexing='ij')
  return torch.stack([yy, xx], dim=2) - 0.5


def make_batch(w: Workload, seed: int = 235, step: int = 0):
  """Returns a dict of CPU tensors for one minibatch.

  embedding       [B, D, H, W] float32 (backbone output, un-normalised)
  semantic_label  [B, H, W] int64   (0..C-1 labelled, C unlabelled, 255 pad)
  instance_label  [B, H, W] int64   (over-segmentation id < 256)
  semantic_tag    [B, 256] int64    (1 where the class occurs in the image)
  local_feature   [B, H, W, loc_channels] float32
  semantic_label_full [B, H*u, W*u] int64, only when w.label_upsample = u > 1
  """
  g = torch.Generator().manual_seed(seed * 1000003 + step)
  B, H, W, D = w.batch, w.height, w.width, w.dim
  G = min(w.num_regions, 255)
  yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32),
                          torch.arange(W, dtype=torch.float32), indexing='ij')
  emb = torch.empty(B, D, H, W)
  sem = torch.empty(B, H, W, dtype=torch.long)
  inst = torch.empty(B, H, W, dtype=torch.long)
  tags = torch.zeros(B, 256, dtype=torch.long)
  by = max(1, int(round(H * w.border / 2))) if w.border > 0 else 0
  bx = max(1, int(round(W * w.border / 2))) if w.border > 0 else 0
  for b in range(B):
    cy = torch.rand(G, generator=g) * H
    cx = torch.rand(G, generator=g) * W
    d2 = (yy[None] - cy[:, None, None]) ** 2 + (xx[None] - cx[:, None, None]) ** 2
    region = d2.argmin(0)                                   # [H, W] in [0, G)
    # a per-image class set; rotate so that images of a batch differ
    perm = torch.randperm(w.num_classes, generator=g)
    classes = perm[:w.classes_per_image]
    region_class = classes[torch.randint(0, w.classes_per_image, (G,), generator=g)]
    cls_map = region_class[region]
    labelled = torch.rand(H, W, generator=g) < w.rho
    s = torch.where(labelled, cls_map, torch.full_like(cls_map, w.num_classes))
    if by > 0:
      s[:by, :] = w.ignore_index
      s[H - by:, :] = w.ignore_index
    if bx > 0:
      s[:, :bx] = w.ignore_index
      s[:, W - bx:] = w.ignore_index
    centres = torch.randn(G, D, generator=g)
    e = centres[region] + w.sigma * torch.randn(H, W, D, generator=g)
    emb[b] = e.permute(2, 0, 1)
    sem[b] = s
    inst[b] = region
    present = torch.unique(s)
    tags[b, present] = 1                                    # list_tag_dataset.py:75-80
  loc = location_features(H, W)
  if w.loc_channels > 2:
    rgb = torch.rand(B, H, W, w.loc_channels - 2, generator=g) * 2 - 1
    local = torch.cat([loc[None].expand(B, H, W, 2), rgb], dim=-1).contiguous()
  else:
    local = loc[None].expand(B, H, W, 2).contiguous()
  out = {'embedding': emb, 'semantic_label': sem, 'instance_label': inst,
         'semantic_tag': tags, 'local_feature': local}
  if w.label_upsample > 1:
    # targets['semantic_label'] of the softmax heads is the full-resolution map; the one the
    # clustering sees is its nearest-neighbour resize (resnet_deeplab.py:163-167)
    u = w.label_upsample
    out['semantic_label_full'] = sem.repeat_interleave(u, 1).repeat_interleave(u, 2)
  return out


def sweep_problem(n_pix: int, dim: int, n_seg: int, seed: int = 235):
  """BASELINE.json configs[4]: an isolated-kernel problem with N_pix unit
  embeddings, N_seg prototypes/seeds and uniformly random class labels."""
  g = torch.Generator().manual_seed(seed + 7919 * n_seg + n_pix)
  side = int(math.isqrt(n_pix))
  assert side * side == n_pix, 'sweep sizes are squares'
  ks = int(math.isqrt(n_seg))
  assert ks * ks == n_seg
  centres = torch.randn(n_seg, dim, generator=g)
  ys = torch.arange(side) * ks // side
  cell = (ys[:, None] + ks * ys[None, :]).reshape(-1)       # column-major seeds
  emb = centres[cell] + torch.randn(n_pix, dim, generator=g)
  emb = emb / emb.norm(dim=1, keepdim=True)
  return {'embedding': emb, 'seed_label': cell, 'side': side,
          'num_clusters': (ks, ks)}
