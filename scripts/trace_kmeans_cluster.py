"""Timeline of kmeans_cluster_kernel (cycles per phase and pass, every CTA).  Needs a trace build
next to the product library:
    make -C spml_b200/csrc BUILD=build_trace TARGET=../libspml_b200_trace.so EXTRA=-DSPML_KM_TRACE
    SPML_B200_LIB=spml_b200/libspml_b200_trace.so python scripts/trace_kmeans_cluster.py [workload]"""
import ctypes, os, sys
os.environ['SPML_B200_BINDING'] = 'ctypes'   # the ATen binding links the product library, not the trace build
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spml_b200 import _lib, synth
from spml_b200.head import generate_clusters
name = sys.argv[1] if len(sys.argv) > 1 else 'voc_scribble_b1'
w = synth.WORKLOADS[name]
for s in range(3):
  b = {k: v.cuda() for k, v in synth.make_batch(w, step=s).items()}
  generate_clusters(b['embedding'], b['semantic_label'], b['instance_label'], b['local_feature'],
                    w.label_divisor, w.ignore_index, list(w.num_clusters), w.iterations)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (256 * 256))()
lib.spml_debug_kc_trace.argtypes = [ctypes.c_void_p]
assert lib.spml_debug_kc_trace(buf) == 0
a = torch.tensor(list(buf)).view(256, 16, 16)
live = (a[:, 1, 0] != 0).nonzero().flatten()
print(name, '- %d CTAs; cycles, median / max over the CTAs' % live.numel())
r0 = a[live, 0]
print('pass 0 (load rows, fp16 operand, initial M-step): %d / %d' %
      (int((r0[:, 1] - r0[:, 0]).median()), int((r0[:, 1] - r0[:, 0]).max())))
names = [(2, 'barrier A'), (3, 'rebuild'), (4, 'barrier B'), (5, 'E-step'), (6, 're-check'), (7, 'M-step')]
for it in range(1, w.iterations + 1):
  r = a[live, it]
  prev = r[:, 0]
  parts = []
  for slot, nm in names:
    if it == w.iterations and slot == 7:
      continue
    d = r[:, slot] - prev
    parts.append('%s %d/%d' % (nm, int(d.median()), int(d.max())))
    prev = r[:, slot]
  total = prev - r[:, 0]
  parts.append('| pass %d/%d' % (int(total.median()), int(total.max())))
  parts.append('| all MMAs issued %d after barrier B' % int((r[:, 8] - r[:, 4]).median()))
  own = r[r[:, 9] > r[:, 2]]
  if own.numel():
    parts.append('| rebuild (warp 0 of the owners that had work): gathered +%d, normalised +%d, operand rows sent +%d'
                 % (int((own[:, 9] - own[:, 2]).median()), int((own[:, 10] - own[:, 9]).median()),
                    int((own[:, 3] - own[:, 10]).median())))
  parts.append('| ambiguous rows %d (max %d per CTA), changed rows %d' %
               (int(r[:, 14].sum()), int(r[:, 14].max()), int(r[:, 15].sum())))
  print('it %2d  %s' % (it, '  '.join(parts)))
