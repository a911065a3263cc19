"""GPU parity tests: the CUDA path (through the C ABI) against the golden outputs of
the reference (tests/golden) and against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): segment ids and every integer output bit-exact;
losses and d(embedding) within 1e-3 relative (fp32)."""

import pytest
import torch

from conftest import load_golden
from helpers import check_step, cuda_step, make_head, norm_err, rel_err
from oracle import spml_oracle as O
from spml_b200 import ops, synth
from spml_b200.head import ContrastiveHead

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,steps', [('tiny', 3), ('small', 3), ('tiny_softmax', 3),
                                        ('tiny_densepose', 2), ('tiny_densepose_shipped', 1)])
def test_steps_match_reference_golden(name, steps):
  """Consecutive steps (memory bank filling) against outputs of the reference, for the three
  heads: segsort.py, segsort_softmax.py (eval mode) and the DensePose variant."""
  w = synth.WORKLOADS[name]
  cfg = synth.make_config(w)
  head = make_head(cfg, w)
  cls64 = O.make_classifier(cfg, dtype=torch.float64).eval() if w.variant != 'segsort' else None
  bank64 = {}
  escapes = []
  for step in range(steps):
    g = load_golden('%s_step%d.pt' % (name, step))
    ref64 = O.contrastive_step(cfg, g['inputs'], bank64, dtype=torch.float64, variant=w.variant,
                               classifier=cls64)
    ours = cuda_step(head, g['inputs'])
    msgs = check_step(ours, g['outputs'], ref64, what='%s step %d' % (name, step),
                      escapes=escapes)
    assert not msgs, '\n'.join(msgs)
    head.update_memory_bank()
    O.memory_bank_update(bank64, {k: ref64[k] for k in ref64 if k.startswith('prototype')},
                         w.memory_bank_size, w.batch)
  assert not escapes, escapes       # nothing needed the fp64 criterion
  if w.memory_bank_size:
    assert len(head.memory_banks['memory_prototype']) == min(steps, w.memory_bank_size)
  ops.check_status()


# keys that may take check_step's fp64 criterion on a full-size workload, each with its reason;
# anything else taking it fails the test
ALLOWED_ESCAPES = ()


@pytest.mark.parametrize('name,seed', [('voc_scribble_b1', 235), ('voc_scribble_b1', 236),
                                       ('voc_tag_b2', 235), ('densepose_b1', 235),
                                       ('densepose_shape_voc_head_b1', 235),
                                       ('voc_scribble_softmax_b1', 235),
                                       ('voc_scribble_b4', 235)])
def test_workloads_match_oracle(name, seed):
  """BASELINE.json configs at full size: two steps (second one with a memory bank)."""
  w = synth.WORKLOADS[name]
  cfg = synth.make_config(w)
  head = make_head(cfg, w)
  cls = O.make_classifier(cfg).eval() if w.variant != 'segsort' else None
  cls64 = O.make_classifier(cfg, dtype=torch.float64).eval() if w.variant != 'segsort' else None
  bank, bank64 = {}, {}
  escapes = []
  torch.set_num_threads(max(1, torch.get_num_threads()))
  for step in range(2):
    batch = synth.make_batch(w, seed=seed, step=step)
    ref = O.contrastive_step(cfg, batch, bank, variant=w.variant, classifier=cls)
    ref64 = O.contrastive_step(cfg, batch, bank64, dtype=torch.float64, variant=w.variant,
                               classifier=cls64)
    ours = cuda_step(head, batch)
    msgs = check_step(ours, ref, ref64, what='%s seed %d step %d' % (name, seed, step),
                      escapes=escapes)
    assert not msgs, '\n'.join(msgs)
    head.update_memory_bank()
    O.memory_bank_update(bank, {k: ref[k] for k in ref if k.startswith('prototype')},
                         w.memory_bank_size, w.batch)
    O.memory_bank_update(bank64, {k: ref64[k] for k in ref64 if k.startswith('prototype')},
                         w.memory_bank_size, w.batch)
  unexpected = [e for e in escapes if not any(a in e for a in ALLOWED_ESCAPES)]
  assert not unexpected, unexpected
  ops.check_status()


def test_step_is_bit_reproducible():
  """Fixed-point segment sums and fixed-order reductions: two runs are identical."""
  w = synth.WORKLOADS['voc_scribble_b4']
  cfg = synth.make_config(w)
  batch = synth.make_batch(w)
  a = cuda_step(ContrastiveHead(cfg).cuda(), batch)
  b = cuda_step(ContrastiveHead(cfg).cuda(), batch)
  for k in a:
    assert torch.equal(a[k], b[k]), k


def test_both_host_bindings_give_the_same_step():
  """The ATen (C++) binding and the ctypes binding are two bookkeeping layers over the same
  library calls: a step through either gives bit-identical results."""
  import json
  import os
  import subprocess
  import sys
  from conftest import ROOT
  code = (
      'import json, torch\n'
      'from spml_b200 import ops, synth\n'
      'from spml_b200.head import ContrastiveHead\n'
      'w = synth.WORKLOADS["small"]; cfg = synth.make_config(w)\n'
      'b = {k: v.cuda() for k, v in synth.make_batch(w).items()}\n'
      'e = b["embedding"].clone().requires_grad_(True)\n'
      'o = ContrastiveHead(cfg).cuda()(e, b["semantic_label"], b["instance_label"], '
      'b["semantic_tag"], b["local_feature"])\n'
      'o["loss"].backward()\n'
      'print(json.dumps({"binding": ops.binding(), "loss": float(o["loss"]), '
      '"acc": float(o["accuracy"]), "grad": float(e.grad.double().abs().sum()), '
      '"ids": int(o["datas"]["cluster_index"].sum())}))\n')
  got = {}
  for binding in ('aten', 'ctypes'):
    env = dict(os.environ, PYTHONPATH=ROOT, SPML_B200_BINDING=binding)
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr
    got[binding] = json.loads(out.stdout.strip().splitlines()[-1])
  assert got['ctypes']['binding'] == 'ctypes'
  if got['aten']['binding'] != 'aten':
    pytest.skip('spml_b200/_C.so is not built')
  for k in ('loss', 'acc', 'grad', 'ids'):
    assert got['aten'][k] == got['ctypes'][k], (k, got)
