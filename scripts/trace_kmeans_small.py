"""Timeline of kmeans_small_kernel (cycles, CTA 0, its last tile of each pass).  Needs a trace
build next to the product library:
    make -C spml_b200/csrc BUILD=build_trace TARGET=../libspml_b200_trace.so EXTRA=-DSPML_KM_TRACE
    SPML_B200_LIB=spml_b200/libspml_b200_trace.so python scripts/trace_kmeans_small.py [workload]"""
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
tr = (ctypes.c_longlong * 256)()
lib.spml_debug_kms_trace.argtypes = [ctypes.c_void_p]
assert lib.spml_debug_kms_trace(tr) == 0
t = torch.tensor(list(tr)).view(16, 16)
names = {1: 'tile ready', 2: 'done seen', 11: 'staged', 12: 'carried', 13: 'normalised', 3: 'protos built', 4: 'mma done', 5: 'merged',
         6: 'rechecked', 10: 'sorted', 7: 'accumulated', 8: 'flushed', 9: 'counted'}
print(name, '(cycles since the CTA that owns tile 0 started its last tile of the pass)')
for it in range(w.iterations + 1):
  r = t[it]
  ev = ' '.join('%s@%d' % (names[k], int(r[k] - r[0])) for k in (1, 2, 11, 12, 13, 3, 4, 5, 6, 10, 7, 8, 9) if r[k] != 0)
  nxt = int(t[it + 1][0] - r[0]) if it < w.iterations else 0
  print('it %2d  %s   next pass @%d' % (it, ev, nxt))

# ---- every CTA: busy time of a pass = (own "counted") - (own "done seen"); the pass is as long
# as the slowest CTA.  Phase lengths: median / max over the CTAs.
allb = (ctypes.c_longlong * (160 * 256))()
lib.spml_debug_kms_trace_all.argtypes = [ctypes.c_void_p]
assert lib.spml_debug_kms_trace_all(allb) == 0
a = torch.tensor(list(allb)).view(160, 16, 16)
live = (a[:, 1, 9] != 0).nonzero().flatten()
order = (2, 11, 12, 13, 3, 4, 5, 6, 7, 8, 9)
print('\nper-CTA phase lengths over %d CTAs (cycles): median / max [CTA of the max]' % live.numel())
for it in range(1, w.iterations):
  r = a[live, it]
  parts = []
  prev = r[:, 2]
  for k in order[1:]:
    d = r[:, k] - prev
    parts.append('%s %d/%d[%d]' % (names[k], int(d.median()), int(d.max()), int(live[d.argmax()])))
    prev = r[:, k]
  amb = r[:, 14]
  parts.append('ambiguous rows %d total, max %d[%d]' % (int(amb.sum()), int(amb.max()), int(live[amb.argmax()])))
  parts.append('exact row (warp 0) %d' % int(r[:, 15].max()))
  busy = r[:, 9] - r[:, 2]
  wait = r[:, 2] - a[live, it - 1, 9] if it > 1 else None
  print('it %2d busy %d/%d[%d]%s  %s' % (
      it, int(busy.median()), int(busy.max()), int(live[busy.argmax()]),
      '' if wait is None else ' waited %d/%d min %d' % (int(wait.median()), int(wait.max()), int(wait.min())),
      ' '.join(parts)))
