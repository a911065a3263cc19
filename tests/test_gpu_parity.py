"""GPU parity tests: the CUDA path (through the C ABI) against the golden outputs of
the reference (tests/golden) and against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): segment ids and every integer output bit-exact;
losses and d(embedding) within 1e-3 relative (fp32)."""

import pytest
import torch

from conftest import load_golden
from helpers import check_step, cuda_step, norm_err, rel_err
from oracle import spml_oracle as O
from spml_b200 import synth
from spml_b200.head import ContrastiveHead

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['tiny', 'small'])
def test_steps_match_reference_golden(name):
  """Three consecutive steps (memory bank filling) against outputs of the reference."""
  w = synth.WORKLOADS[name]
  cfg = synth.make_config(w)
  head = ContrastiveHead(cfg).cuda()
  bank64 = {}
  for step in range(3):
    g = load_golden('%s_step%d.pt' % (name, step))
    ref64 = O.contrastive_step(cfg, g['inputs'], bank64, dtype=torch.float64)
    ours = cuda_step(head, g['inputs'])
    msgs = check_step(ours, g['outputs'], ref64, what='%s step %d' % (name, step))
    assert not msgs, '\n'.join(msgs)
    head.update_memory_bank()
    O.memory_bank_update(bank64, {k: ref64[k] for k in ref64 if k.startswith('prototype')},
                         w.memory_bank_size, w.batch)
  assert len(head.memory_banks['memory_prototype']) == min(3, w.memory_bank_size)


@pytest.mark.parametrize('name,seed', [('voc_scribble_b1', 235), ('voc_scribble_b1', 236),
                                       ('voc_tag_b2', 235), ('densepose_b1', 235),
                                       ('voc_scribble_b4', 235)])
def test_workloads_match_oracle(name, seed):
  """BASELINE.json configs at full size: two steps (second one with a memory bank)."""
  w = synth.WORKLOADS[name]
  if name == 'densepose_b1':
    w = synth.dataclasses.replace(w, loc_channels=5)
  cfg = synth.make_config(w)
  head = ContrastiveHead(cfg).cuda()
  bank, bank64 = {}, {}
  torch.set_num_threads(max(1, torch.get_num_threads()))
  for step in range(2):
    batch = synth.make_batch(w, seed=seed, step=step)
    ref = O.contrastive_step(cfg, batch, bank)
    ref64 = O.contrastive_step(cfg, batch, bank64, dtype=torch.float64)
    ours = cuda_step(head, batch)
    msgs = check_step(ours, ref, ref64, what='%s seed %d step %d' % (name, seed, step))
    assert not msgs, '\n'.join(msgs)
    head.update_memory_bank()
    O.memory_bank_update(bank, {k: ref[k] for k in ref if k.startswith('prototype')},
                         w.memory_bank_size, w.batch)
    O.memory_bank_update(bank64, {k: ref64[k] for k in ref64 if k.startswith('prototype')},
                         w.memory_bank_size, w.batch)


def test_step_is_bit_reproducible():
  """Fixed-point segment sums and fixed-order reductions: two runs are identical."""
  w = synth.WORKLOADS['voc_scribble_b4']
  cfg = synth.make_config(w)
  batch = synth.make_batch(w)
  a = cuda_step(ContrastiveHead(cfg).cuda(), batch)
  b = cuda_step(ContrastiveHead(cfg).cuda(), batch)
  for k in a:
    assert torch.equal(a[k], b[k]), k
