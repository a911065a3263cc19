"""StaticContrastiveHead (fixed-capacity buffers, no host sync, CUDA graph) against
ContrastiveHead (the reference's data-dependent shapes) and the oracle."""

import pytest
import torch

from helpers import cuda_step, norm_err
from oracle import spml_oracle as O
from spml_b200 import synth
from spml_b200.head import ContrastiveHead
from spml_b200.static_head import StaticContrastiveHead

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,use_graph', [('small', False), ('small', True),
                                            ('voc_scribble_b1', True), ('voc_tag_b2', True)])
def test_static_head_matches_dynamic_head(name, use_graph):
  w = synth.WORKLOADS[name]
  cfg = synth.make_config(w)
  dyn = ContrastiveHead(cfg).cuda()
  sta = StaticContrastiveHead(cfg, w.batch, w.height, w.width, w.loc_channels,
                              use_graph=use_graph)
  for step in range(4):          # the memory bank fills up and starts to rotate
    batch = synth.make_batch(w, step=step)
    ref = cuda_step(dyn, batch)
    dyn.update_memory_bank()
    b = {k: v.cuda() for k, v in batch.items()}
    out = sta.step(b['embedding'], b['semantic_label'], b['instance_label'], b['semantic_tag'],
                   b['local_feature'])
    torch.cuda.synchronize()
    sta.check_overflow()
    n, m = int(out['num_pixels']), int(out['num_segments'])
    assert n == ref['cluster_index'].numel() and m == ref['prototype'].shape[0]
    assert torch.equal(out['cluster_index'][:n].cpu(), ref['cluster_index'])
    assert torch.equal(out['prototype_semantic_label'][:m].cpu(), ref['prototype_semantic_label'])
    assert int(out['prototype_live'].sum()) == m
    for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss', 'accuracy', 'loss'):
      a, r = float(out[k]), float(ref[k])
      assert abs(a - r) <= 1e-5 * max(abs(r), 1e-3), (step, k, a, r)
    assert norm_err(out['grad_embedding'].cpu(), ref['grad_embedding']) < 1e-5, step


def test_static_head_matches_oracle_with_bank():
  w = synth.WORKLOADS['voc_scribble_b4']
  cfg = synth.make_config(w)
  sta = StaticContrastiveHead(cfg, w.batch, w.height, w.width)
  bank = {}
  for step in range(3):
    batch = synth.make_batch(w, step=step)
    ref = O.contrastive_step(cfg, batch, bank)
    O.memory_bank_update(bank, {k: ref[k] for k in ref if k.startswith('prototype')},
                         w.memory_bank_size, w.batch)
    b = {k: v.cuda() for k, v in batch.items()}
    out = sta.step(b['embedding'], b['semantic_label'], b['instance_label'], b['semantic_tag'],
                   b['local_feature'])
    n = int(out['num_pixels'])
    assert torch.equal(out['cluster_index'][:n].cpu(), ref['cluster_index'])
    for k in ('sem_ann_loss', 'sem_occ_loss', 'img_sim_loss'):
      assert abs(float(out[k]) - float(ref[k])) <= 1e-3 * abs(float(ref[k])) + 1e-7, (step, k)
    assert abs(float(out['accuracy']) - float(ref['accuracy'])) < 1e-6
    assert norm_err(out['grad_embedding'].cpu(), ref['grad_embedding']) < 1e-3


def test_overflow_is_reported():
  w = synth.WORKLOADS['small']
  cfg = synth.make_config(w)
  sta = StaticContrastiveHead(cfg, w.batch, w.height, w.width, max_segments=16, use_graph=False)
  b = {k: v.cuda() for k, v in synth.make_batch(w).items()}
  sta.step(b['embedding'], b['semantic_label'], b['instance_label'], b['semantic_tag'],
           b['local_feature'])
  with pytest.raises(RuntimeError, match='max_segments'):
    sta.check_overflow()
