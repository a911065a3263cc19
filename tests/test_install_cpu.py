"""install() rebinds the reference's symbols (only where the reference is importable:
the build container; the GPU box has no /root/reference)."""

import importlib
import os
import sys

import pytest

REF = '/root/reference'


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'spml')), reason='reference not present')
def test_install_rebinds_and_uninstall_restores():
  sys.path.insert(0, REF)
  try:
    import spml_b200
    inst = importlib.import_module('spml_b200.install')
    common = importlib.import_module('spml.utils.segsort.common')
    loss = importlib.import_module('spml.utils.segsort.loss')
    orig_fn, orig_cls = common.segment_by_kmeans, loss.SegSortLoss
    done = spml_b200.install()
    assert 'spml.utils.segsort.common.segment_by_kmeans' in done
    assert common.segment_by_kmeans is spml_b200.segsort_common.segment_by_kmeans
    assert loss.SegSortLoss is spml_b200.segsort_loss.SegSortLoss
    # the reference's own call sites resolve through the module attribute
    pred = importlib.import_module('spml.models.predictions.segsort')
    assert pred.segsort_loss.SegSortLoss is spml_b200.segsort_loss.SegSortLoss
    assert pred.Segsort is spml_b200.predictions.Segsort
    # A9: the generate_clusters METHOD of the embedding models, and the DensePose head
    deeplab = importlib.import_module('spml.models.embeddings.resnet_deeplab')
    psp_dp = importlib.import_module('spml.models.embeddings.resnet_pspnet_densepose')
    dp = importlib.import_module('spml.models.predictions.segsort_softmax_densepose')
    assert 'spml.models.embeddings.resnet_deeplab.ResnetDeeplab.generate_clusters' in done
    assert deeplab.ResnetDeeplab.generate_clusters.__module__ == 'spml_b200.head'
    assert psp_dp.ResnetPspnet.generate_clusters.__module__ == 'spml_b200.head'
    assert dp.SegsortSoftmax is spml_b200.predictions.SegsortSoftmaxDensepose
    mu = importlib.import_module('spml.models.utils')
    assert (mu.gather_multiset_labels_per_batch_by_nearest_neighbor
            is spml_b200.model_utils.gather_multiset_labels_per_batch_by_nearest_neighbor)
    spml_b200.uninstall()
    assert common.segment_by_kmeans is orig_fn and loss.SegSortLoss is orig_cls
    assert deeplab.ResnetDeeplab.generate_clusters.__module__ == deeplab.__name__
    assert len(inst.BINDINGS) == 11
  finally:
    sys.path.remove(REF)
    for k in [k for k in sys.modules if k == 'spml' or k.startswith('spml.')]:
      del sys.modules[k]


def test_binding_table_names_exist_in_package():
  inst = importlib.import_module('spml_b200.install')
  for mod, attrs in inst.BINDINGS.items():
    for name, repl in attrs.items():
      assert callable(repl), (mod, name)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'spml')), reason='reference not present')
def test_run_launcher_executes_an_unchanged_script_with_rebound_symbols(tmp_path):
  """python -m spml_b200.run SCRIPT: the script sees the rebound reference modules."""
  import subprocess
  script = tmp_path / 'probe.py'
  script.write_text(
      'import sys\n'
      'import spml.utils.segsort.common as c, spml.utils.segsort.loss as l\n'
      'print(c.segment_by_kmeans.__module__, l.SegSortLoss.__module__, sys.argv[1:])\n')
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ, PYTHONPATH=os.pathsep.join([root, REF]))
  out = subprocess.run([sys.executable, '-m', 'spml_b200.run', '--spml-b200-lenient', str(script),
                        '--flag', '7'], env=env, capture_output=True, text=True, timeout=300)
  assert out.returncode == 0, out.stderr
  assert out.stdout.split()[:2] == ['spml_b200.segsort_common', 'spml_b200.segsort_loss']
  assert "['--flag', '7']" in out.stdout
