"""ResNet-101 DeepLab embedding network for the end-to-end training benchmark.

BASELINE.json's first metric clause is "training images/sec (512x512, ResNet-101)".  The
backbone is NOT part of the hot path (north_star: "the ResNet/DeepLab backbone forward/backward
stays PyTorch/cuDNN"); this module only exists so that `bench.py --workload train_voc_b4` can
time the whole step: cuDNN backbone -> contrastive head (libspml_b200) -> backward -> gradient
all-reduce (NCCL) -> SGD.  It follows the architecture the reference trains
(spml/models/backbones/resnet.py:66-178 with spml/models/embeddings/resnet_deeplab.py:20-88,
spml/models/heads/spp.py): three 3x3 stem convolutions, bottleneck stages [3, 4, 23, 3] with
res3 at stride 2 and res4 / res5 dilated 2 / 4 (output stride 8), an ASPP head that sums four
dilated 3x3 convolutions (rates 6, 12, 18, 24) into `embedding_dim` channels, and a 2x bilinear
up-sampling: a 512x512 crop gives a 128x128 embedding map; 47.3 M parameters.  Weights are
random (there is no network access for checkpoints), which does not change the step time.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _bn(channels):
  return nn.BatchNorm2d(channels, momentum=3e-4)


class _Bottleneck(nn.Module):
  """1x1 reduce -> 3x3 (strided / dilated) -> 1x1 expand (x4), residual."""

  def __init__(self, cin, width, stride=1, dilation=1, project=False):
    super().__init__()
    cout = width * 4
    self.reduce = nn.Conv2d(cin, width, 1, bias=False)
    self.bn_a = _bn(width)
    self.conv = nn.Conv2d(width, width, 3, stride=stride, padding=dilation, dilation=dilation,
                          bias=False)
    self.bn_b = _bn(width)
    self.expand = nn.Conv2d(width, cout, 1, bias=False)
    self.bn_c = _bn(cout)
    self.shortcut = None
    if project:
      self.shortcut = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), _bn(cout))

  def forward(self, x):
    y = F.relu(self.bn_a(self.reduce(x)), inplace=True)
    y = F.relu(self.bn_b(self.conv(y)), inplace=True)
    y = self.bn_c(self.expand(y))
    return F.relu(y + (x if self.shortcut is None else self.shortcut(x)), inplace=True)


def _stage(cin, width, blocks, stride, dilation):
  first_dilation = 1 if dilation in (1, 2) else 2          # resnet.py:143-150
  layers = [_Bottleneck(cin, width, stride, first_dilation, project=True)]
  layers += [_Bottleneck(width * 4, width, 1, dilation) for _ in range(blocks - 1)]
  return nn.Sequential(*layers)


class ResnetDeeplabEmbedding(nn.Module):
  """images [B, 3, H, W] -> embedding [B, embedding_dim, H / 4, W / 4]."""

  def __init__(self, embedding_dim=64, blocks=(3, 4, 23, 3)):
    super().__init__()
    self.stem = nn.Sequential(
        nn.Conv2d(3, 64, 3, stride=2, padding=1, bias=False), _bn(64), nn.ReLU(inplace=True),
        nn.Conv2d(64, 64, 3, padding=1, bias=False), _bn(64), nn.ReLU(inplace=True),
        nn.Conv2d(64, 128, 3, padding=1, bias=False), _bn(128), nn.ReLU(inplace=True),
        nn.MaxPool2d(3, stride=2, padding=1))
    self.res2 = _stage(128, 64, blocks[0], 1, 1)
    self.res3 = _stage(256, 128, blocks[1], 2, 1)
    self.res4 = _stage(512, 256, blocks[2], 1, 2)
    self.res5 = _stage(1024, 512, blocks[3], 1, 4)
    self.aspp = nn.ModuleList(
        nn.Conv2d(2048, embedding_dim, 3, padding=r, dilation=r, bias=True)
        for r in (6, 12, 18, 24))
    for m in self.modules():
      if isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')

  def forward(self, images):
    x = self.res5(self.res4(self.res3(self.res2(self.stem(images)))))
    y = self.aspp[0](x)
    for conv in self.aspp[1:]:
      y = y + conv(x)
    return F.interpolate(y, scale_factor=2, mode='bilinear', align_corners=False)


def num_parameters(module):
  return sum(p.numel() for p in module.parameters())
