"""bench.py host logic that needs no GPU: the secondary training-step measurement is started as a
second command on its own rendezvous port, outside torchrun's agent store, and can never cost the
headline line."""

import importlib.util
import json
import os
import subprocess
import types

from conftest import ROOT


def load_bench():
  spec = importlib.util.spec_from_file_location('bench_under_test', os.path.join(ROOT, 'bench.py'))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


def test_training_step_arm_environment_and_result(monkeypatch):
  bench = load_bench()
  seen = {}

  def fake_run(cmd, env=None, capture_output=None, text=None, timeout=None):
    seen['cmd'], seen['env'], seen['timeout'] = cmd, env, timeout
    line = {'metric': 'm', 'value': 123.0, 'unit': 'images/s', 'n_gpus': 2, 'ms_per_step': 31.0,
            'steps': 10, 'warmup': 3, 'scaling': 'weak', 'dtype': 'bf16', 'config': {'workload': 'train_voc_b4'},
            'phase_ms': {}, 'all_reduce': {'ms': 0.4}, 'contrastive_head_ms': 1.9, 'clocks': {}}
    return types.SimpleNamespace(stdout='NCCL version 2.28\n' + json.dumps(line) + '\n', stderr='')

  monkeypatch.setattr(subprocess, 'run', fake_run)
  monkeypatch.setenv('MASTER_PORT', '29513')
  monkeypatch.setenv('TORCHELASTIC_USE_AGENT_STORE', 'True')
  monkeypatch.setenv('TORCHELASTIC_RUN_ID', 'x')
  monkeypatch.setenv('RANK', '1')
  out = bench.training_step_arm(types.SimpleNamespace(gpus=2))
  assert out['value'] == 123.0 and out['all_reduce'] == {'ms': 0.4} and out['n_gpus'] == 2
  env = seen['env']
  assert not any(k.startswith('TORCHELASTIC') for k in env)      # rank 0 of the child hosts its own store
  assert env['RANK'] == '1'
  port = int(env['MASTER_PORT'])
  assert port != 29513 and 1024 <= port < 65536
  assert seen['cmd'][-6:] == ['--gpus', '2', '--steps', '10', '--warmup', '3']
  assert '--workload' in seen['cmd'] and 'train_voc_b4' in seen['cmd']
  assert seen['timeout'] and seen['timeout'] <= 300


def test_training_step_arm_never_raises(monkeypatch):
  bench = load_bench()

  def hang(*a, **k):
    raise subprocess.TimeoutExpired(cmd='bench.py', timeout=1)

  monkeypatch.setattr(subprocess, 'run', hang)
  out = bench.training_step_arm(types.SimpleNamespace(gpus=8))
  assert 'error' in out and 'TimeoutExpired' in out['error']

  def silent(*a, **k):
    return types.SimpleNamespace(stdout='', stderr='boom')

  monkeypatch.setattr(subprocess, 'run', silent)
  assert bench.training_step_arm(types.SimpleNamespace(gpus=1)) == {'error': 'boom'}
